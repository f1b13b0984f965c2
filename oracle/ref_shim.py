"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference (/root/reference) with the oracle's random numbers.

The reference draws from numpy PCG64 Generators (one per stochastic process,
mbt_gym/stochastic_processes/StochasticProcessModel.py:27).  Here each Generator is replaced by a
duck-typed object that serves, for env-step n, exactly the numbers the Philox draw contract
(include/mbt_philox.h) assigns to (trajectory, n):
    arrival_model.rng.uniform(size=(N,2))          -> u[:, 0:2]     arrival_models.py:55,122
    fill_probability_model.rng.uniform(size=(N,2)) -> u[:, 2:4]     fill_probability_models.py:33
    midprice_model.rng.normal(size=(N,1))          -> z[:, None]    midprice_models.py:64,143
    env.rng.integers(lo, hi, size=N)               -> reset-stream  TradingEnvironment.py:272
    np.random.multivariate_normal(0, corr, size=N) -> (z, rho z + sqrt(1-rho^2) z2)   midprice_models.py:357 (Heston draws
                                                      from the GLOBAL numpy generator; patched while the pair runs)
so `reference.step()` and `orc_step_core_f64` can be compared trajectory by trajectory.

Only usable where /root/reference exists (the build container); tools/make_golden.py turns its output
into committed fixtures under tests/golden/ for the GPU box.
"""
import os
import sys

import numpy as np

from mbt_gym_b200 import _abi
from oracle import oracle as O

REFERENCE_ROOT = os.environ.get("MBT_REFERENCE_ROOT", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_stub")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mbt_gym"))


def import_reference():
    """Put the gym stub and the reference on sys.path (once) and return the `mbt_gym` package."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (_STUB, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mbt_gym  # noqa: F401

    return mbt_gym


class DrawSource:
    """Holds the random numbers of the env-step about to run."""

    def __init__(self, seed, N, traj_offset=0):
        self.seed, self.N, self.traj_offset = int(seed), int(N), int(traj_offset)
        self.n_step = 0
        self.n_episode = 0
        self.u = None
        self.z = None

    def load_step(self):
        self.u, self.z = O.draws(_abi.MBT_F64, self.seed, self.traj_offset, self.N, self.n_step)
        self.z2 = O.draws2(_abi.MBT_F64, self.seed, self.traj_offset, self.N, self.n_step)
        self.n_step += 1

    def multivariate_normal(self, mean, cov, size=None):
        """Stands in for np.random.multivariate_normal in HestonMidpriceModel.update (midprice_models.py:355-357): the
        correlated pair of the draw contract, W_S = z, W_v = rho z + sqrt(1 - rho^2) z2."""
        assert int(size) == self.N and np.all(np.asarray(mean) == 0)
        rho = float(np.asarray(cov)[0, 1])
        return np.stack([self.z, rho * self.z + np.sqrt(1.0 - rho * rho) * self.z2], axis=1)


class InjectedRng:
    """Stands in for numpy.random.Generator on ONE reference object."""

    def __init__(self, source, role):
        self.source, self.role = source, role

    def uniform(self, low=0.0, high=1.0, size=None):
        assert tuple(size) == (self.source.N, 2), size
        cols = {"arrival": slice(0, 2), "fill": slice(2, 4)}[self.role]
        return self.source.u[:, cols].copy()

    def normal(self, loc=0.0, scale=1.0, size=None):
        assert self.role == "midprice" and tuple(size) == (self.source.N, 1), (self.role, size)
        return self.source.z.reshape(-1, 1).copy()

    def integers(self, low, high=None, size=None):
        assert self.role == "env"
        out = O.q0_draws(self.source.seed, self.source.traj_offset, self.source.N, self.source.n_episode, int(low), int(high))
        return out


# --------------------------------------------------------------------------- spec -> reference env / mbt_config
def build_reference_env(spec):
    """Construct the reference TradingEnvironment described by `spec` (plain dict, see tools/make_golden.py)."""
    import_reference()
    from mbt_gym.gym.TradingEnvironment import TradingEnvironment
    from mbt_gym.gym import ModelDynamics as MD
    from mbt_gym.stochastic_processes import midprice_models as MM, arrival_models as AM
    from mbt_gym.stochastic_processes import fill_probability_models as FM, price_impact_models as PM
    from mbt_gym.rewards import RewardFunctions as RF

    N, n_steps, T = spec["N"], spec["n_steps"], spec["terminal_time"]
    dt = T / n_steps
    m = spec["midprice"]
    if m["kind"] == "bm":
        mid = MM.BrownianMotionMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"],
                                             initial_price=m["initial_price"], terminal_time=T, step_size=dt,
                                             num_trajectories=N)
    elif m["kind"] == "gbm":
        mid = MM.GeometricBrownianMotionMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"],
                                                      initial_price=m["initial_price"], terminal_time=T, step_size=dt,
                                                      num_trajectories=N)
    elif m["kind"] == "ou":
        mid = MM.OuMidpriceModel(mean_reversion_level=m["level"], mean_reversion_speed=m["speed"],
                                 volatility=m["volatility"], initial_price=m["initial_price"], terminal_time=T,
                                 step_size=dt, num_trajectories=N)
    elif m["kind"] == "bm_jump":
        mid = MM.BrownianMotionJumpMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"],
                                                 jump_size=m["jump"], initial_price=m["initial_price"], terminal_time=T,
                                                 step_size=dt, num_trajectories=N)
    elif m["kind"] == "ou_jump":
        mid = MM.OuJumpMidpriceModel(mean_reversion_level=m["level"], mean_reversion_speed=m["speed"],
                                     volatility=m["volatility"], jump_size=m["jump"], initial_price=m["initial_price"],
                                     terminal_time=T, step_size=dt, num_trajectories=N)
    elif m["kind"] == "heston":
        mid = MM.HestonMidpriceModel(drift=m["drift"], volatility_mean_reversion_rate=m["speed"],
                                     volatility_mean_reversion_level=m["level"], weiner_correlation=m["corr"],
                                     volatility_of_volatility=m["volvol"], initial_price=m["initial_price"],
                                     initial_variance=m["initial_variance"], terminal_time=T, step_size=dt,
                                     num_trajectories=N)
    elif m["kind"] == "constant":
        mid = MM.ConstantMidpriceModel(initial_price=m["initial_price"], terminal_time=T, step_size=dt,
                                       num_trajectories=N)
    else:
        raise ValueError(m)
    arr = fill = imp = None
    a = spec.get("arrival")
    if a:
        if a["kind"] == "poisson":
            arr = AM.PoissonArrivalModel(intensity=np.array(a["intensity"], float), step_size=dt, num_trajectories=N)
        elif a["kind"] == "poisson_nonlinear":
            arr = AM.PoissonArrivalNonLinearModel(intensity=np.array(a["intensity"], float), step_size=dt,
                                                  num_trajectories=N)
        elif a["kind"] == "hawkes":
            arr = AM.HawkesArrivalModel(baseline_arrival_rate=np.array([a["baseline"]], float), step_size=dt,
                                        jump_size=a["jump"], mean_reversion_speed=a["speed"], terminal_time=T,
                                        num_trajectories=N)
    f = spec.get("fill")
    if f:
        if f.get("kind", "exp") == "exogenous":
            depth_models = tuple(MM.OuMidpriceModel(mean_reversion_level=d0, mean_reversion_speed=0.1, volatility=0.05,
                                                    initial_price=d0, terminal_time=T, step_size=dt, num_trajectories=N)
                                 for d0 in f["best_depths"])
            fill = FM.ExogenousMmFillProbabilityModel(depth_models, fill_exponent=f["fill_exponent"],
                                                      base_fill_probability=f["base"], step_size=dt, num_trajectories=N)
        elif f.get("kind", "exp") == "triangular":
            fill = FM.TriangularFillFunction(max_fill_depth=f["max_fill_depth"], step_size=dt, num_trajectories=N)
        elif f.get("kind", "exp") == "power":
            fill = FM.PowerFillFunction(fill_exponent=f["fill_exponent"], fill_multiplier=f["fill_multiplier"],
                                        step_size=dt, num_trajectories=N)
        else:
            fill = FM.ExponentialFillFunction(fill_exponent=f["fill_exponent"], step_size=dt, num_trajectories=N)
    p = spec.get("impact")
    if p:
        if p["kind"] == "temp_perm":
            imp = PM.TemporaryAndPermanentPriceImpact(temporary_impact_coefficient=p["temp"],
                                                      permanent_impact_coefficient=p["perm"], n_steps=n_steps,
                                                      terminal_time=T, num_trajectories=N)
        elif p["kind"] == "temp_power":
            imp = PM.TemporaryPowerPriceImpact(temporary_impact_coefficient=p["temp"],
                                               temporary_impact_exponent=p["exponent"], num_trajectories=N)
        elif p["kind"] == "temp_transient":
            imp = PM.TemporaryAndTransientPriceImpact(p["temp"], p["transient"], p["resilience"], p["initial"], p["kernel"],
                                                      n_steps=n_steps, terminal_time=T, num_trajectories=N)
        elif p["kind"] == "transient":
            imp = PM.TransientPriceImpact(p["transient"], p["resilience"], p["initial"], p["kernel"], n_steps=n_steps,
                                          terminal_time=T, num_trajectories=N)
    dyn_kind = spec["dynamics"]
    if dyn_kind == "limit":
        dyn = MD.LimitOrderModelDynamics(midprice_model=mid, arrival_model=arr, fill_probability_model=fill,
                                         num_trajectories=N)
    elif dyn_kind == "touch":
        dyn = MD.AtTheTouchModelDynamics(midprice_model=mid, arrival_model=arr, num_trajectories=N,
                                         fixed_market_half_spread=spec.get("half_spread", 0.5))
    elif dyn_kind == "limit_and_market":
        dyn = MD.LimitAndMarketOrderModelDynamics(midprice_model=mid, arrival_model=arr, fill_probability_model=fill,
                                                  num_trajectories=N,
                                                  fixed_market_half_spread=spec.get("half_spread", 0.5))
    elif dyn_kind == "speed":
        dyn = MD.TradinghWithSpeedModelDynamics(midprice_model=mid, price_impact_model=imp, num_trajectories=N)
    else:
        raise ValueError(dyn_kind)
    r = spec["reward"]
    if r["kind"] == "pnl":
        rew = RF.PnL()
    elif r["kind"] == "rip":
        rew = RF.RunningInventoryPenalty(r["phi"], r["alpha"], r.get("exponent", 2.0))
    elif r["kind"] == "cjmm":
        rew = RF.CjMmCriterion(r["phi"], r["alpha"], r.get("exponent", 2.0), T)
    elif r["kind"] == "cjoe":
        rew = RF.CjOeCriterion(r["phi"], r["alpha"], r.get("exponent", 2.0), T)
    elif r["kind"] == "exputil":
        rew = RF.ExponentialUtility(r["risk_aversion"])
    else:
        raise ValueError(r)
    q0 = spec.get("initial_inventory", 0)
    if isinstance(q0, list):
        q0 = tuple(q0)
    env = TradingEnvironment(terminal_time=T, n_steps=n_steps, reward_function=rew, model_dynamics=dyn,
                             initial_cash=spec.get("initial_cash", 0.0), initial_inventory=q0,
                             max_inventory=spec.get("max_inventory", 10_000), max_cash=spec.get("max_cash"),
                             start_time=spec.get("start_time", 0.0), seed=None, num_trajectories=N,
                             normalise_action_space=spec.get("normalise_action", False),
                             normalise_observation_space=spec.get("normalise_obs", False),
                             normalise_rewards=False)
    if spec.get("normalise_rewards"):
        # the constructor's own bootstrap (TradingEnvironment.py:90-94,329-343: a 100 000-trajectory rollout from the
        # env's generators) cannot be driven by injected draws; the reward-scaling leg of step() (:128-129) is pinned with
        # the scale set after construction -- exactly the two attributes the constructor would have set
        env.normalise_rewards_ = True
        env.reward_scaling = float(spec["reward_scaling"])
    return env


_DYN = {"limit": _abi.MBT_DYN_LIMIT, "speed": _abi.MBT_DYN_SPEED, "touch": _abi.MBT_DYN_AT_TOUCH,
        "limit_and_market": _abi.MBT_DYN_LIMIT_AND_MARKET}
_MID = {"constant": _abi.MBT_MID_CONSTANT, "bm": _abi.MBT_MID_BM, "gbm": _abi.MBT_MID_GBM, "ou": _abi.MBT_MID_OU,
        "bm_jump": _abi.MBT_MID_BM_JUMP, "ou_jump": _abi.MBT_MID_OU_JUMP, "heston": _abi.MBT_MID_HESTON}
_ARR = {"poisson": _abi.MBT_ARR_POISSON, "poisson_nonlinear": _abi.MBT_ARR_POISSON_NONLINEAR,
        "hawkes": _abi.MBT_ARR_HAWKES}
_REW = {"pnl": _abi.MBT_REW_PNL, "rip": _abi.MBT_REW_RUNNING_INVENTORY_PENALTY, "cjmm": _abi.MBT_REW_CJ_MM,
        "cjoe": _abi.MBT_REW_CJ_OE, "exputil": _abi.MBT_REW_EXP_UTILITY}


def config_from_reference_env(spec, env, precision=_abi.MBT_F64, traj_offset=0):
    """Flatten a constructed reference env into mbt_config, reading every number off the reference objects."""
    md = env.model_dynamics
    cfg = _abi.new_config(precision=precision, num_trajectories=env.num_trajectories, traj_offset=traj_offset,
                          n_steps=env.n_steps, dynamics=_DYN[spec["dynamics"]],
                          terminal_time=env.terminal_time, step_size=env.step_size,
                          start_time=float(env._get_start_time()), initial_cash=env.initial_cash,
                          max_inventory=env.max_inventory, max_cash=env.max_cash)
    q0 = env.initial_inventory
    if isinstance(q0, tuple):
        cfg.q0_mode, cfg.q0_lo, cfg.q0_hi = _abi.MBT_Q0_UNIFORM_INT, int(q0[0]), int(q0[1])
    else:
        cfg.q0_mode, cfg.q0_const = _abi.MBT_Q0_CONST, float(q0)
    mid = md.midprice_model
    cfg.midprice = _MID[spec["midprice"]["kind"]]
    cfg.mid_initial = float(mid.initial_state[0, 0])
    cfg.mid_drift = float(getattr(mid, "drift", 0.0))
    cfg.mid_vol = float(getattr(mid, "volatility", 0.0))
    cfg.mid_step = float(mid.step_size)
    cfg.ou_level = float(getattr(mid, "mean_reversion_level", 0.0))
    cfg.ou_speed = float(getattr(mid, "mean_reversion_speed", 0.0))
    cfg.mid_jump = float(getattr(mid, "jump_size", 0.0))
    if cfg.midprice == _abi.MBT_MID_HESTON:
        cfg.heston_speed = float(mid.volatility_mean_reversion_rate)
        cfg.heston_level = float(mid.volatility_mean_reversion_level)
        cfg.heston_corr = float(mid.weiner_correlation)
        cfg.heston_volvol = float(mid.volatility_of_volatility)
        cfg.heston_var0 = float(mid.initial_state[0, 1])
    arr = md.arrival_model
    if arr is not None:
        cfg.arrival = _ARR[spec["arrival"]["kind"]]
        cfg.arr_step = float(arr.step_size)
        if cfg.arrival == _abi.MBT_ARR_HAWKES:
            cfg.arr_rate[0], cfg.arr_rate[1] = (float(x) for x in np.asarray(arr.baseline_arrival_rate).reshape(-1))
            cfg.hawkes_jump, cfg.hawkes_speed = float(arr.jump_size), float(arr.mean_reversion_speed)
        else:
            cfg.arr_rate[0], cfg.arr_rate[1] = (float(x) for x in arr.intensity)
    fill = md.fill_probability_model
    if fill is not None:
        cfg.fill = {"exp": _abi.MBT_FILL_EXPONENTIAL, "triangular": _abi.MBT_FILL_TRIANGULAR,
                    "power": _abi.MBT_FILL_POWER, "exogenous": _abi.MBT_FILL_EXOGENOUS_MM}[spec["fill"].get("kind", "exp")]
        if cfg.fill == _abi.MBT_FILL_EXOGENOUS_MM:
            cfg.fill_base = float(fill.base_fill_probability)
            cfg.fill_depth0[0], cfg.fill_depth0[1] = (float(x) for x in fill.initial_state[0])
        cfg.fill_exponent = float(getattr(fill, "fill_exponent", 0.0))
        cfg.fill_max_depth = float(getattr(fill, "max_fill_depth", 0.0))
        cfg.fill_multiplier = float(getattr(fill, "fill_multiplier", 0.0))
    imp = md.price_impact_model
    if imp is not None:
        kind = spec["impact"]["kind"]
        cfg.imp_temp = float(getattr(imp, "temporary_impact_coefficient", 0.0))
        if kind == "temp_perm":
            cfg.impact = _abi.MBT_IMP_TEMP_PERM
            cfg.imp_perm, cfg.imp_step = float(imp.permanent_impact_coefficient), float(imp.step_size)
        elif kind == "temp_power":
            cfg.impact = _abi.MBT_IMP_TEMP_POWER
            cfg.imp_exponent = float(imp.temporary_impact_exponent)
        else:
            cfg.impact = _abi.MBT_IMP_TEMP_TRANSIENT if kind == "temp_transient" else _abi.MBT_IMP_TRANSIENT
            cfg.imp_transient, cfg.imp_resilience = float(imp.transient_impact_coefficient), float(imp.resilience_coefficient)
            cfg.imp_kernel, cfg.imp_initial = float(imp.linear_kernel_coefficient), float(imp.initial_transient_impact)
            cfg.imp_step = float(imp.step_size)
    cfg.half_spread = float(getattr(md, "fixed_market_half_spread", 0.0))
    rf = env.reward_function
    cfg.reward = _REW[spec["reward"]["kind"]]
    cfg.rew_phi = float(getattr(rf, "per_step_inventory_aversion", 0.0))
    cfg.rew_alpha = float(getattr(rf, "terminal_inventory_aversion", 0.0))
    cfg.rew_exponent = float(getattr(rf, "inventory_exponent", 2.0))
    cfg.rew_terminal_time = float(getattr(rf, "terminal_time", env.terminal_time))
    cfg.rew_risk_aversion = float(getattr(rf, "risk_aversion", 0.0))
    cfg.normalise_action = int(env.normalise_action_space_)
    cfg.normalise_obs = int(env.normalise_observation_space_)
    cfg.normalise_rewards = int(bool(env.normalise_rewards_))
    if env.normalise_rewards_:
        cfg.reward_scaling = float(env.reward_scaling)
    if env.normalise_action_space_:
        lo, hi = env.original_action_space.low, env.original_action_space.high
        for i in range(lo.shape[0]):
            cfg.act_low[i] = float(lo[i])
            cfg.act_grad[i] = float(((hi - lo) / 2)[i])
    if env.normalise_observation_space_:
        lo, hi = env.original_observation_space.low, env.original_observation_space.high
        for i in range(lo.shape[0]):
            cfg.obs_low[i] = float(lo[i])
            cfg.obs_grad[i] = float(((hi - lo) / 2)[i])
    return cfg


def inject(env, source):
    """Swap every Generator the hot path touches for the injected source."""
    md = env.model_dynamics
    md.midprice_model.rng = InjectedRng(source, "midprice")
    if md.arrival_model is not None:
        md.arrival_model.rng = InjectedRng(source, "arrival")
    if md.fill_probability_model is not None:
        md.fill_probability_model.rng = InjectedRng(source, "fill")
    env.rng = InjectedRng(source, "env")


def make_actions(spec, env, n_steps_run, action_seed):
    """Per-step action matrices (n_steps_run, N, A) exercising in-range, out-of-range and edge values."""
    rng = np.random.default_rng(action_seed)
    N = env.num_trajectories
    space = env.action_space
    lo = np.asarray(space.low, float) if hasattr(space, "low") else np.zeros(2)
    hi = np.asarray(space.high, float) if hasattr(space, "high") else np.ones(2)
    A = lo.shape[0]
    span = hi - lo
    acts = rng.uniform(lo - 0.15 * span, hi + 0.05 * span, size=(n_steps_run, N, A))  # slightly out of range too
    if "depth_range" in spec:
        # batch-reduced fill functions: the step's fill probability follows the DEEPEST quote of the batch, so the
        # depths are drawn below a per-step ceiling that sweeps the interesting range (and sometimes beyond it)
        d_lo, d_hi = spec["depth_range"]
        ceil_k = rng.uniform(d_lo, d_hi, size=(n_steps_run, 1, 1))
        depths = d_lo + rng.uniform(0.0, 1.0, size=(n_steps_run, N, 2)) * (ceil_k - d_lo)
        if spec.get("normalise_action"):  # given in normalised units, de-normalised by the env
            o_lo = np.asarray(env.original_action_space.low, float)[:2]
            o_hi = np.asarray(env.original_action_space.high, float)[:2]
            depths = (depths - o_lo) / ((o_hi - o_lo) / 2) - 1
        acts[:, :, :2] = depths
    if spec["dynamics"] == "touch":
        acts = rng.integers(0, 2, size=(n_steps_run, N, 2)).astype(float)
    if spec["dynamics"] == "limit_and_market":
        acts[:, :, 2:] = rng.uniform(0, 1, size=(n_steps_run, N, 2))
    return acts


def run_pair(spec, n_steps_run=None, action_seed=7, actions=None, n_episodes=1):
    """Run reference (injected draws) and oracle (f64) side by side; return dict of per-step outputs of both."""
    env = build_reference_env(spec)
    cfg = config_from_reference_env(spec, env)
    source = DrawSource(spec["seed"], spec["N"])
    inject(env, source)
    orc = O.OracleEnv(cfg)
    orc.seed(spec["seed"])
    n_run = n_steps_run or spec["n_steps"]
    if actions is None:
        actions = make_actions(spec, env, n_run * n_episodes, action_seed)
    out = dict(cfg=cfg, actions=actions, ref_obs=[], ref_rew=[], ref_done=[], orc_obs=[], orc_rew=[], orc_done=[],
               ref_reset=[], orc_reset=[])
    global_mvn = np.random.multivariate_normal
    np.random.multivariate_normal = source.multivariate_normal  # only HestonMidpriceModel.update calls it
    try:
        _run_episodes(spec, env, orc, source, actions, out, n_run, n_episodes)
    finally:
        np.random.multivariate_normal = global_mvn
    for key in ("ref_obs", "ref_rew", "orc_obs", "orc_rew", "ref_reset", "orc_reset"):
        out[key] = np.stack(out[key])
    out["ref_state"] = np.array(env.state, float)
    out["orc_state"] = orc.state
    return out


def _run_episodes(spec, env, orc, source, actions, out, n_run, n_episodes):
    import contextlib
    import io
    k = 0
    for ep in range(n_episodes):
        source.n_episode = ep
        ref_obs0 = env.reset()
        orc_obs0 = orc.reset()
        out["ref_reset"].append(np.array(ref_obs0, float))
        out["orc_reset"].append(orc_obs0)
        for _ in range(n_run):
            source.load_step()
            with contextlib.redirect_stdout(io.StringIO()):  # the reference prints whole arrays when it clips
                o, r, d, _ = env.step(actions[k])
            oo, orr, od = orc.step(actions[k])
            out["ref_obs"].append(np.array(o, float))
            out["ref_rew"].append(np.broadcast_to(np.asarray(r, float), (spec["N"],)).copy())
            out["ref_done"].append(bool(np.asarray(d).reshape(-1)[0]))
            out["orc_obs"].append(oo)
            out["orc_rew"].append(orr)
            out["orc_done"].append(od)
            k += 1
