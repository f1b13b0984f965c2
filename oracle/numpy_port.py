"""TEST / BENCH INFRASTRUCTURE: NumPy restatement of the reference's step(), used ONLY as the timed CPU baseline.

`/root/reference` does not exist on the GPU box, so `bench.py --impl reference` and the `cpu_baseline` leg time this
port instead.  It performs the same sequence of whole-array NumPy operations and PCG64 `Generator` draws as the
reference (one Generator per stochastic process, (N,2) boolean arrival/fill arrays, two state copies per step), so its
throughput is representative of the reference's; tests/test_numpy_port.py checks it reproduces the reference's
notebook goldens (Test_1 AS table) where /root/reference's PCG64 streams are matched exactly.

Citations are into /root/reference/mbt_gym/.
"""
import numpy as np

CASH, INV, TIME, PRICE = 0, 1, 2, 3  # gym/index_names.py:1-4


class NumpyPortEnv:
    """AS / Hawkes limit-order market making and OU optimal execution, the BASELINE.json configurations."""

    def __init__(self, kind="as", N=1000, n_steps=200, T=1.0, seed=None, sigma=2.0, S0=100.0, lam=(140.0, 140.0),
                 kappa=1.5, reward="pnl", phi=0.01, alpha=0.001, max_inventory=None, q0=0,
                 hawkes=(10.0, 40.0, 60.0), ou=(100.0, 1.0), impact=(0.01, 0.01)):
        self.kind, self.N, self.n_steps, self.T = kind, N, n_steps, T
        self.dt = T / n_steps
        self.sigma, self.S0, self.kappa = sigma, S0, kappa
        self.lam = np.array(lam, float)
        self.reward, self.phi, self.alpha = reward, phi, alpha
        self.q0 = q0
        self.max_inventory = max_inventory if max_inventory is not None else n_steps
        self.max_cash = n_steps * (S0 + 4 * sigma * np.sqrt(T))  # TradingEnvironment.py:229-230
        self.hawkes, self.ou, self.impact = hawkes, ou, impact
        self.D = {"as": 4, "hawkes": 6, "oe": 5}[kind]
        self.A = 1 if kind == "oe" else 2
        # if seed: process i gets seed+i+1 (midprice, arrival, fill order)   TradingEnvironment.py:70-71,345-348
        self.rng_mid = np.random.default_rng(seed + 1 if seed else None)
        self.rng_arr = np.random.default_rng(seed + 2 if seed else None)
        self.rng_fill = np.random.default_rng(seed + 3 if seed else None)
        self.fill_multiplier = np.append(-np.ones((N, 1)), np.ones((N, 1)), axis=1)  # ModelDynamics.py:71-73
        self.state = None
        self.reset()

    def reset(self):  # TradingEnvironment.py:96-101,131-140
        s = np.zeros((self.N, self.D))
        s[:, INV] = self.q0
        s[:, PRICE] = self.S0
        if self.kind == "hawkes":
            s[:, 4:6] = self.hawkes[0]
        self.state = s
        self.q_init = s[:, INV].copy()
        return s.copy()

    def step(self, action):  # TradingEnvironment.py:103-110
        N, dt = self.N, self.dt
        cur = self.state.copy()
        st = self.state
        mid = st[:, PRICE].reshape(-1, 1).copy()
        if self.kind == "oe":  # ModelDynamics.py:262-267, price_impact_models.py:88-92
            k_tmp, b_perm = self.impact
            px = mid + (k_tmp * action + st[:, 4:5])
            vol = action * dt
            st[:, CASH] -= np.squeeze(vol * px)
            st[:, INV] += np.squeeze(vol)
            arrivals = None
        else:  # ModelDynamics.py:127-131, arrival_models.py:54-56,121-123, fill_probability_models.py:28-34
            unif = self.rng_arr.uniform(size=(N, 2))
            arrivals = unif < (st[:, 4:6] * dt if self.kind == "hawkes" else self.lam * dt)
            unif = self.rng_fill.uniform(size=(N, 2))
            fills = unif < np.exp(-self.kappa * action[:, 0:2])
            keep = np.concatenate(((1 - (st[:, INV] >= self.max_inventory)).reshape(-1, 1),
                                   (1 - (st[:, INV] <= -self.max_inventory)).reshape(-1, 1)), axis=1)
            fills = keep * fills  # TradingEnvironment.py:323-327
            st[:, INV] += np.sum(arrivals * fills * -self.fill_multiplier, axis=1)  # ModelDynamics.py:108-116
            st[:, CASH] += np.sum(self.fill_multiplier * arrivals * fills * (mid + action[:, 0:2] * self.fill_multiplier), axis=1)
        st[:, INV] = np.clip(st[:, INV], -self.max_inventory, self.max_inventory)  # TradingEnvironment.py:283-297
        st[:, CASH] = np.clip(st[:, CASH], -self.max_cash, self.max_cash)
        st[:, TIME] += dt
        z = self.rng_mid.normal(size=(N, 1))
        if self.kind == "oe":  # midprice_models.py:140-143 (literal: drift not scaled by dt)
            theta, kap = self.ou
            st[:, PRICE:PRICE + 1] = mid + (-kap * (mid - theta * np.ones((N, 1))) + self.sigma * np.sqrt(dt) * z)
            st[:, 4:5] = st[:, 4:5] + self.impact[1] * action * dt
        else:  # midprice_models.py:60-65
            st[:, PRICE:PRICE + 1] = mid + 0.0 * dt * np.ones((N, 1)) + self.sigma * np.sqrt(dt) * z
        if self.kind == "hawkes":  # arrival_models.py:110-119
            lbar, eta, beta = self.hawkes
            st[:, 4:6] = (st[:, 4:6] + beta * (np.ones((N, 2)) * lbar - st[:, 4:6]) * dt * np.ones((N, 2)) + eta * arrivals)
        done = st[0, TIME] >= self.T - dt / 2  # TradingEnvironment.py:218-220
        dones = np.full((N,), done, dtype=bool)
        pnl = (st[:, CASH] + st[:, INV] * st[:, PRICE]) - (cur[:, CASH] + cur[:, INV] * cur[:, PRICE])  # RewardFunctions.py:23-33
        if self.reward == "pnl":
            rew = pnl
        else:
            d = st[:, TIME] - cur[:, TIME]
            if self.reward == "cjmm":  # RewardFunctions.py:96-109
                rew = (pnl - d * self.phi * st[:, INV] ** 2.0
                       - self.alpha * (st[:, INV] ** 2.0 - cur[:, INV] ** 2.0 + d / self.T * self.q_init ** 2.0))
            else:  # cjoe  RewardFunctions.py:55-70
                rew = (pnl - d * self.phi * st[:, INV] ** 2.0
                       - d * self.alpha * (2.0 * np.squeeze(action) * cur[:, INV] ** 1.0 + self.q_init ** 2.0 * self.T))
        return st.copy(), rew, dones, None


def as_agent_action(state, gamma, sigma, kappa, T):
    """AvellanedaStoikovAgent.get_action   agents/BaselineAgents.py:62-83"""
    q, t = state[:, INV], state[:, TIME]
    adj = q * gamma * sigma ** 2 * (T - t)
    spread = gamma * sigma ** 2 * (T - t) + 2 / gamma * np.log(1 + gamma / kappa)
    return np.append((adj + spread / 2).reshape(-1, 1), (-adj + spread / 2).reshape(-1, 1), axis=1)


def make_env(workload, N, seed=None):
    if workload == "as":
        return NumpyPortEnv("as", N=N, seed=seed)
    if workload == "cjmm":
        return NumpyPortEnv("as", N=N, seed=seed, reward="cjmm", max_inventory=100)
    if workload == "hawkes":
        return NumpyPortEnv("hawkes", N=N, seed=seed, lam=(10.0, 10.0))
    if workload == "oe":
        return NumpyPortEnv("oe", N=N, seed=seed, reward="cjoe", q0=100, max_inventory=10_000)
    raise ValueError(workload)


def fixed_action(workload, N):
    return np.full((N, 1), -1.0) if workload == "oe" else np.full((N, 2), 0.7)


def _worker(args):
    import time
    workload, n, steps, warmup = args
    env = make_env(workload, n, seed=None)
    a = fixed_action(workload, n)
    env.reset()
    for _ in range(warmup):
        env.step(a)
    t0 = time.perf_counter()
    for k in range(steps):
        _o, _r, d, _ = env.step(a)
        if d[0]:
            env.reset()
    return time.perf_counter() - t0


def time_port(workload, N, steps, warmup, procs):
    """One env per process with N/procs trajectories each (the reference's MultiprocessTradingEnv design,
    gym/MultiprocessTradingEnv.py:72-95); returns (seconds of the slowest worker, trajectories actually run)."""
    import multiprocessing as mp
    per = max(1, N // procs)
    if procs == 1:
        return _worker((workload, per, steps, warmup)), per
    with mp.get_context("fork").Pool(procs) as pool:
        times = pool.map(_worker, [(workload, per, steps, warmup)] * procs)
    return max(times), per * procs
