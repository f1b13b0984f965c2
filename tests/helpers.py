"""Shared test helpers: golden fixtures (reference outputs) and config handling."""
import ctypes as C
import json
import os

import numpy as np

from mbt_gym_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return sorted(json.load(f).keys())


def golden_specs():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return json.load(f)


class Golden:
    """One fixture: the reference's outputs for a recorded action sequence (tools/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        raw = z["cfg"].tobytes()
        assert len(raw) == C.sizeof(_abi.mbt_config), "fixture was written with a different mbt_config layout"
        self.cfg = _abi.mbt_config.from_buffer_copy(raw)
        self.seed = int(z["seed"])
        self.n_episodes = int(z["n_episodes"])
        self.actions = z["actions"]        # (steps_total, N, A)
        self.reset_obs = z["reset_obs"]    # (n_episodes, N, D)
        self.obs = z["obs"]                # (steps_total, N, D)
        self.rew = z["rew"]                # (steps_total, N)
        self.done = z["done"].astype(bool)  # (steps_total,)
        self.final_state = z["final_state"]
        self.steps_per_episode = self.obs.shape[0] // self.n_episodes

    def config(self, precision=_abi.MBT_F64, **overrides):
        cfg = _abi.mbt_config.from_buffer_copy(bytes(self.cfg))
        cfg.precision = precision
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg

    # libm pow / exp (numpy) vs mbt_math differ by ulps in these two fixtures only
    @property
    def exact(self):
        return self.name not in ("rip_cubic", "exputil")


def copy_config(cfg, **overrides):
    out = _abi.mbt_config.from_buffer_copy(bytes(cfg))
    for k, v in overrides.items():
        setattr(out, k, v)
    return out


def assert_same(a, b, exact=True, rtol=1e-11, atol=1e-11, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if exact:
        if not np.array_equal(a, b, equal_nan=True):
            bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
            i = tuple(bad[0])
            raise AssertionError(f"{what}: {len(bad)} of {a.size} values differ; first at {i}: {a[i]!r} vs {b[i]!r}")
    else:
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def build_facade_env(spec, **extra):
    """The mbt_gym_b200 environment for a fixture spec, built with the same keyword arguments the reference takes
    (mirror of oracle/ref_shim.build_reference_env)."""
    from mbt_gym_b200.gym import ModelDynamics as MD
    from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment
    from mbt_gym_b200.rewards import RewardFunctions as RF
    from mbt_gym_b200.stochastic_processes import arrival_models as AM, fill_probability_models as FM
    from mbt_gym_b200.stochastic_processes import midprice_models as MM, price_impact_models as PM

    N, n_steps, T = spec["N"], spec["n_steps"], spec["terminal_time"]
    dt = T / n_steps
    m = spec["midprice"]
    kw = dict(initial_price=m["initial_price"], terminal_time=T, step_size=dt, num_trajectories=N)
    if m["kind"] == "bm":
        mid = MM.BrownianMotionMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"], **kw)
    elif m["kind"] == "gbm":
        mid = MM.GeometricBrownianMotionMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"], **kw)
    elif m["kind"] == "ou":
        mid = MM.OuMidpriceModel(mean_reversion_level=m["level"], mean_reversion_speed=m["speed"],
                                 volatility=m["volatility"], **kw)
    elif m["kind"] == "bm_jump":
        mid = MM.BrownianMotionJumpMidpriceModel(drift=m.get("drift", 0.0), volatility=m["volatility"], jump_size=m["jump"], **kw)
    elif m["kind"] == "ou_jump":
        mid = MM.OuJumpMidpriceModel(mean_reversion_level=m["level"], mean_reversion_speed=m["speed"],
                                     volatility=m["volatility"], jump_size=m["jump"], **kw)
    elif m["kind"] == "heston":
        mid = MM.HestonMidpriceModel(drift=m["drift"], volatility_mean_reversion_rate=m["speed"],
                                     volatility_mean_reversion_level=m["level"], weiner_correlation=m["corr"],
                                     volatility_of_volatility=m["volvol"], initial_variance=m["initial_variance"], **kw)
    else:
        mid = MM.ConstantMidpriceModel(**kw)
    arr = fill = imp = None
    a = spec.get("arrival")
    if a:
        if a["kind"] == "poisson":
            arr = AM.PoissonArrivalModel(intensity=np.array(a["intensity"], float), step_size=dt, num_trajectories=N)
        elif a["kind"] == "poisson_nonlinear":
            arr = AM.PoissonArrivalNonLinearModel(intensity=np.array(a["intensity"], float), step_size=dt, num_trajectories=N)
        else:
            arr = AM.HawkesArrivalModel(baseline_arrival_rate=np.array([a["baseline"]], float), step_size=dt,
                                        jump_size=a["jump"], mean_reversion_speed=a["speed"], terminal_time=T,
                                        num_trajectories=N)
    f = spec.get("fill")
    if f and f.get("kind", "exp") == "triangular":
        fill = FM.TriangularFillFunction(max_fill_depth=f["max_fill_depth"], step_size=dt, num_trajectories=N)
    elif f and f.get("kind", "exp") == "exogenous":
        depth_models = tuple(MM.OuMidpriceModel(mean_reversion_level=d0, mean_reversion_speed=0.1, volatility=0.05,
                                                initial_price=d0, terminal_time=T, step_size=dt, num_trajectories=N)
                             for d0 in f["best_depths"])
        fill = FM.ExogenousMmFillProbabilityModel(depth_models, fill_exponent=f["fill_exponent"],
                                                  base_fill_probability=f["base"], step_size=dt, num_trajectories=N)
    elif f and f.get("kind", "exp") == "power":
        fill = FM.PowerFillFunction(fill_exponent=f["fill_exponent"], fill_multiplier=f["fill_multiplier"], step_size=dt,
                                    num_trajectories=N)
    elif f:
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
        else:
            imp = PM.TransientPriceImpact(p["transient"], p["resilience"], p["initial"], p["kernel"], n_steps=n_steps,
                                          terminal_time=T, num_trajectories=N)
    kind = spec["dynamics"]
    if kind == "limit":
        dyn = MD.LimitOrderModelDynamics(midprice_model=mid, arrival_model=arr, fill_probability_model=fill, num_trajectories=N)
    elif kind == "touch":
        dyn = MD.AtTheTouchModelDynamics(midprice_model=mid, arrival_model=arr, num_trajectories=N,
                                         fixed_market_half_spread=spec.get("half_spread", 0.5))
    elif kind == "limit_and_market":
        dyn = MD.LimitAndMarketOrderModelDynamics(midprice_model=mid, arrival_model=arr, fill_probability_model=fill,
                                                  num_trajectories=N, fixed_market_half_spread=spec.get("half_spread", 0.5))
    else:
        dyn = MD.TradinghWithSpeedModelDynamics(midprice_model=mid, price_impact_model=imp, num_trajectories=N)
    r = spec["reward"]
    rew = {"pnl": lambda: RF.PnL(),
           "rip": lambda: RF.RunningInventoryPenalty(r["phi"], r["alpha"], r.get("exponent", 2.0)),
           "cjmm": lambda: RF.CjMmCriterion(r["phi"], r["alpha"], r.get("exponent", 2.0), T),
           "cjoe": lambda: RF.CjOeCriterion(r["phi"], r["alpha"], r.get("exponent", 2.0), T),
           "exputil": lambda: RF.ExponentialUtility(r["risk_aversion"])}[r["kind"]]()
    q0 = spec.get("initial_inventory", 0)
    if isinstance(q0, list):
        q0 = tuple(q0)
    env = TradingEnvironment(terminal_time=T, n_steps=n_steps, reward_function=rew, model_dynamics=dyn,
                             initial_cash=spec.get("initial_cash", 0.0), initial_inventory=q0,
                             max_inventory=spec.get("max_inventory", 10_000), max_cash=spec.get("max_cash"),
                             start_time=spec.get("start_time", 0.0), seed=spec["seed"], num_trajectories=N,
                             normalise_action_space=spec.get("normalise_action", False),
                             normalise_observation_space=spec.get("normalise_obs", False), normalise_rewards=False,
                             **extra)
    if spec.get("normalise_rewards"):
        # like oracle/ref_shim.build_reference_env: the scale the constructor's bootstrap would have set, given by the spec
        env.normalise_rewards_ = True
        env.reward_scaling = float(spec["reward_scaling"])
    return env
