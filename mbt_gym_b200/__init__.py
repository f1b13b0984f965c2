"""mbt_gym_b200 -- the vectorised limit-order-book trading environment of JJJerome/mbt_gym with its hot path
(`TradingEnvironment.step` / `reset` / rollouts) as hand-written sm_100a CUDA kernels behind a C ABI.

    from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment
    env = TradingEnvironment(num_trajectories=1 << 20)        # same keywords as the reference
    obs = env.reset(); obs, rewards, dones, infos = env.step(action)

Importing the package never touches CUDA; creating an environment loads mbt_gym_b200/libmbt_b200.so and fails loudly
if it (or a CUDA device) is missing -- there is no CPU path.
"""
__version__ = "0.1.0"
