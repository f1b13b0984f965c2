#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

For every spec below the reference env is built from its own classes, its numpy Generators are swapped
for the Philox draw contract (oracle/ref_shim.py), and it is stepped with a recorded action sequence.
Each fixture stores: the flattened mbt_config bytes, the seed, the actions, and the REFERENCE's
observations / rewards / dones / final state.  Tests then require
   oracle (f64)  == fixture   bit-for-bit  (tests/test_oracle_vs_reference_golden.py, CPU)
   CUDA   (f64)  == fixture   bit-for-bit  (tests/test_gpu_parity.py, -m gpu)
The GPU box has no /root/reference; the fixtures are how the reference travels.

Run (build container only):  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402

AS = dict(midprice=dict(kind="bm", volatility=2.0, initial_price=100.0),
          arrival=dict(kind="poisson", intensity=[140.0, 140.0]), fill=dict(kind="exp", fill_exponent=1.5))

SPECS = {
    # C1/C2: Avellaneda-Stoikov, PnL (BASELINE.json configs[0], [1]); max_inventory = n_steps as in Test_1 notebook
    "as_pnl": dict(N=193, n_steps=200, terminal_time=1.0, seed=50, dynamics="limit", reward=dict(kind="pnl"),
                   max_inventory=200, **AS),
    # same market, tiny inventory / cash limits: max-inventory fill suppression and both clips fire
    "as_pnl_tight": dict(N=131, n_steps=60, terminal_time=1.0, seed=51, dynamics="limit", reward=dict(kind="pnl"),
                         max_inventory=2, max_cash=150.0, **AS),
    # normalised action + observation spaces (the reference's default ctor flags)
    "as_pnl_normalised": dict(N=97, n_steps=40, terminal_time=1.0, seed=52, dynamics="limit",
                              reward=dict(kind="pnl"), max_inventory=10_000, normalise_action=True,
                              normalise_obs=True, **AS),
    # C3: CJP-2015 market making, CjMmCriterion, random initial inventory, late start
    "cjmm": dict(N=149, n_steps=100, terminal_time=1.0, seed=410, dynamics="limit",
                 reward=dict(kind="cjmm", phi=0.01, alpha=0.001), max_inventory=100, initial_inventory=[-5, 6],
                 start_time=0.1, **AS),
    "rip": dict(N=101, n_steps=50, terminal_time=1.0, seed=411, dynamics="limit",
                reward=dict(kind="rip", phi=0.01, alpha=0.5), max_inventory=100, initial_inventory=3, **AS),
    "rip_cubic": dict(N=67, n_steps=30, terminal_time=1.0, seed=412, dynamics="limit",
                      reward=dict(kind="rip", phi=0.01, alpha=0.5, exponent=4.0), max_inventory=100, **AS),
    # C4: Hawkes arrivals (reference defaults lambda_bar=10, eta=40, beta=60)
    "hawkes_pnl": dict(N=163, n_steps=200, terminal_time=1.0, seed=1234, dynamics="limit", reward=dict(kind="pnl"),
                       max_inventory=200, midprice=AS["midprice"], fill=AS["fill"],
                       arrival=dict(kind="hawkes", baseline=[10.0, 10.0], jump=40.0, speed=60.0)),
    "hawkes_normalised": dict(N=89, n_steps=50, terminal_time=1.0, seed=1235, dynamics="limit",
                              reward=dict(kind="cjmm", phi=0.02, alpha=0.01), max_inventory=50,
                              normalise_action=True, normalise_obs=True, midprice=AS["midprice"], fill=AS["fill"],
                              arrival=dict(kind="hawkes", baseline=[20.0, 15.0], jump=30.0, speed=50.0)),
    # C5: optimal execution: speed dynamics + OU midprice + temporary & permanent impact
    "oe_ou_cjoe": dict(N=157, n_steps=200, terminal_time=1.0, seed=1236, dynamics="speed",
                       midprice=dict(kind="ou", level=100.0, speed=1.0, volatility=2.0, initial_price=100.0),
                       impact=dict(kind="temp_perm", temp=0.01, perm=0.01),
                       reward=dict(kind="cjoe", phi=0.01, alpha=0.001), initial_inventory=100, max_inventory=10_000),
    "oe_ou_pnl_normalised": dict(N=83, n_steps=50, terminal_time=1.0, seed=1237, dynamics="speed",
                                 midprice=dict(kind="ou", level=100.0, speed=0.05, volatility=1.0, initial_price=101.0),
                                 impact=dict(kind="temp_perm", temp=0.02, perm=0.005), reward=dict(kind="pnl"),
                                 initial_inventory=50, max_inventory=60, normalise_action=True, normalise_obs=True),
    "oe_bm_power": dict(N=71, n_steps=40, terminal_time=2.0, seed=1238, dynamics="speed",
                        midprice=dict(kind="bm", drift=0.5, volatility=1.5, initial_price=50.0),
                        impact=dict(kind="temp_power", temp=0.01, exponent=1.0), reward=dict(kind="pnl"),
                        initial_inventory=10, max_inventory=1000),
    # section 8f rank 3 variants
    "gbm_nonlinear": dict(N=73, n_steps=40, terminal_time=1.0, seed=1239, dynamics="limit",
                          midprice=dict(kind="gbm", drift=0.05, volatility=0.2, initial_price=100.0),
                          arrival=dict(kind="poisson_nonlinear", intensity=[120.0, 90.0]), fill=AS["fill"],
                          reward=dict(kind="pnl"), max_inventory=100),
    "constant_mid": dict(N=41, n_steps=20, terminal_time=1.0, seed=1240, dynamics="limit",
                         midprice=dict(kind="constant", initial_price=100.0), arrival=AS["arrival"], fill=AS["fill"],
                         reward=dict(kind="rip", phi=0.1, alpha=0.1), max_inventory=100),
    "touch": dict(N=79, n_steps=40, terminal_time=1.0, seed=1241, dynamics="touch", half_spread=0.5,
                  midprice=AS["midprice"], arrival=AS["arrival"], reward=dict(kind="pnl"), max_inventory=4),
    "limit_and_market": dict(N=61, n_steps=40, terminal_time=1.0, seed=1242, dynamics="limit_and_market",
                             half_spread=0.25, reward=dict(kind="rip", phi=0.01, alpha=0.01), max_inventory=5, **AS),
    "exputil": dict(N=53, n_steps=25, terminal_time=1.0, seed=1243, dynamics="limit",
                    reward=dict(kind="exputil", risk_aversion=0.01), max_inventory=100, **AS),
    "bm_jump": dict(N=57, n_steps=40, terminal_time=1.0, seed=1245, dynamics="limit",
                    midprice=dict(kind="bm_jump", drift=0.1, volatility=1.5, jump=0.25, initial_price=100.0),
                    arrival=AS["arrival"], fill=AS["fill"], reward=dict(kind="pnl"), max_inventory=3),
    "ou_jump_hawkes": dict(N=47, n_steps=40, terminal_time=1.0, seed=1246, dynamics="limit",
                           midprice=dict(kind="ou_jump", level=100.0, speed=0.1, volatility=1.0, jump=0.5, initial_price=100.5),
                           arrival=dict(kind="hawkes", baseline=[30.0, 20.0], jump=20.0, speed=40.0), fill=AS["fill"],
                           reward=dict(kind="rip", phi=0.01, alpha=0.1), max_inventory=50, normalise_obs=True),
    "oe_temp_transient": dict(N=43, n_steps=40, terminal_time=1.0, seed=1247, dynamics="speed",
                              midprice=dict(kind="bm", volatility=1.0, initial_price=100.0),
                              impact=dict(kind="temp_transient", temp=0.01, transient=0.5, resilience=2.0, initial=0.02, kernel=0.3),
                              reward=dict(kind="cjoe", phi=0.01, alpha=0.001), initial_inventory=20, max_inventory=1000),
    "oe_transient": dict(N=37, n_steps=30, terminal_time=1.0, seed=1248, dynamics="speed",
                         midprice=dict(kind="gbm", drift=0.02, volatility=0.1, initial_price=80.0),
                         impact=dict(kind="transient", transient=0.8, resilience=1.0, initial=0.0, kernel=0.5),
                         reward=dict(kind="pnl"), initial_inventory=-15, max_inventory=1000, normalise_action=True),
    # the reference's default-constructor flags (normalised actions + observations) with the inventory-averse rewards:
    # the compile-time "normalised" kernel variants
    "cjmm_normalised": dict(N=93, n_steps=40, terminal_time=1.0, seed=1252, dynamics="limit",
                            reward=dict(kind="cjmm", phi=0.01, alpha=0.001), max_inventory=100, initial_inventory=[-4, 5],
                            normalise_action=True, normalise_obs=True, **AS),
    "rip_normalised": dict(N=87, n_steps=40, terminal_time=1.0, seed=1253, dynamics="limit",
                           reward=dict(kind="rip", phi=0.02, alpha=0.3), max_inventory=7, normalise_action=True,
                           normalise_obs=True, **AS),
    # Heston stochastic-volatility midprice: two state columns (price, variance), a second normal per step
    "heston": dict(N=91, n_steps=40, terminal_time=1.0, seed=1254, dynamics="limit", reward=dict(kind="pnl"),
                   max_inventory=20, arrival=AS["arrival"], fill=AS["fill"],
                   midprice=dict(kind="heston", drift=0.05, speed=3.0, level=0.04, corr=-0.8, volvol=0.6,
                                 initial_price=100.0, initial_variance=0.04)),
    "heston_hawkes": dict(N=53, n_steps=30, terminal_time=1.0, seed=1255, dynamics="limit",
                          reward=dict(kind="rip", phi=0.01, alpha=0.1), max_inventory=20, fill=AS["fill"],
                          arrival=dict(kind="hawkes", baseline=[30.0, 20.0], jump=20.0, speed=25.0),
                          normalise_action=True,
                          midprice=dict(kind="heston", drift=0.05, speed=3.0, level=0.04, corr=0.3, volvol=1.5,
                                        initial_price=100.0, initial_variance=0.5)),
    "heston_oe": dict(N=47, n_steps=30, terminal_time=1.0, seed=1256, dynamics="speed",
                      impact=dict(kind="temp_perm", temp=0.01, perm=0.01),
                      reward=dict(kind="cjoe", phi=0.01, alpha=0.001), initial_inventory=20, max_inventory=1000,
                      midprice=dict(kind="heston", drift=0.0, speed=2.0, level=0.09, corr=1.0, volvol=0.3,
                                    initial_price=50.0, initial_variance=0.09)),
    # fill functions whose probability is a BATCH reduction in the reference (np.max(depths, 0) over trajectories,
    # fill_probability_models.py:82,113)
    "triangular_fill": dict(N=77, n_steps=40, terminal_time=1.0, seed=1249, dynamics="limit", reward=dict(kind="pnl"),
                            max_inventory=6, midprice=AS["midprice"], arrival=AS["arrival"],
                            fill=dict(kind="triangular", max_fill_depth=1.0), depth_range=[-0.2, 1.3]),
    "power_fill": dict(N=85, n_steps=40, terminal_time=1.0, seed=1250, dynamics="limit",
                       reward=dict(kind="rip", phi=0.01, alpha=0.1), max_inventory=8, midprice=AS["midprice"],
                       arrival=AS["arrival"], fill=dict(kind="power", fill_exponent=1.5, fill_multiplier=1.5),
                       depth_range=[0.0, 2.5], normalise_action=True, normalise_obs=True),
    "power_fill_limit_and_market": dict(N=45, n_steps=30, terminal_time=1.0, seed=1251, dynamics="limit_and_market",
                                        half_spread=0.25, reward=dict(kind="pnl"), max_inventory=5,
                                        midprice=AS["midprice"], arrival=AS["arrival"],
                                        fill=dict(kind="power", fill_exponent=2.0, fill_multiplier=0.8),
                                        depth_range=[0.0, 3.0]),
    # exogenous-market-maker fill model: two (constant) best-depth columns in the observation
    "exogenous_fill": dict(N=71, n_steps=40, terminal_time=1.0, seed=1257, dynamics="limit", reward=dict(kind="pnl"),
                           max_inventory=8, midprice=AS["midprice"], arrival=AS["arrival"],
                           fill=dict(kind="exogenous", fill_exponent=1.5, base=0.8, best_depths=[0.5, 0.4])),
    "exogenous_fill_hawkes_normalised": dict(N=43, n_steps=30, terminal_time=1.0, seed=1258, dynamics="limit",
                                             reward=dict(kind="rip", phi=0.01, alpha=0.1), max_inventory=6,
                                             midprice=AS["midprice"],
                                             arrival=dict(kind="hawkes", baseline=[30.0, 20.0], jump=20.0, speed=25.0),
                                             fill=dict(kind="exogenous", fill_exponent=2.0, base=1.0, best_depths=[0.3, 0.6]),
                                             normalise_action=True, normalise_obs=True),
    # reward normalisation (TradingEnvironment.py:128-129): rewards * reward_scaling, the scale set like the constructor's
    # bootstrap would (ref_shim.build_reference_env); alone, with the other two normalisations, and with a random q0
    "as_pnl_reward_scaled": dict(N=69, n_steps=40, terminal_time=1.0, seed=1259, dynamics="limit", reward=dict(kind="pnl"),
                                 max_inventory=40, normalise_rewards=True, reward_scaling=0.03713528, **AS),
    "rip_all_normalised": dict(N=75, n_steps=40, terminal_time=1.0, seed=1260, dynamics="limit",
                               reward=dict(kind="rip", phi=0.02, alpha=0.3), max_inventory=9, normalise_action=True,
                               normalise_obs=True, normalise_rewards=True, reward_scaling=1.0 / 27.31, **AS),
    "cjmm_reward_scaled": dict(N=63, n_steps=30, terminal_time=1.0, seed=1261, dynamics="limit",
                               reward=dict(kind="cjmm", phi=0.01, alpha=0.001), max_inventory=30, initial_inventory=[-4, 5],
                               normalise_obs=True, normalise_rewards=True, reward_scaling=3.25, n_episodes=2, **AS),
    # two episodes back to back: RNG stream continues, reset redraws inventories
    "two_episodes": dict(N=59, n_steps=30, terminal_time=1.0, seed=1244, dynamics="limit",
                         reward=dict(kind="cjmm", phi=0.01, alpha=0.001), max_inventory=20,
                         initial_inventory=[-3, 4], n_episodes=2, **AS),
}


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    index = {}
    only = set(sys.argv[1:])  # `python tools/make_golden.py name ...` rewrites just those fixtures (index.json always)
    for name, spec in SPECS.items():
        index[name] = spec
        if only and name not in only:
            continue
        n_ep = spec.get("n_episodes", 1)
        out = R.run_pair(spec, n_episodes=n_ep)
        same = (np.array_equal(out["ref_obs"], out["orc_obs"]) and np.array_equal(out["ref_rew"], out["orc_rew"])
                and out["ref_done"] == out["orc_done"] and np.array_equal(out["ref_reset"], out["orc_reset"]))
        err = max(np.max(np.abs(out["ref_obs"] - out["orc_obs"])), np.max(np.abs(out["ref_rew"] - out["orc_rew"])))
        print(f"{name:24s} steps={out['ref_obs'].shape[0]:4d} N={spec['N']:4d} oracle==reference bitwise: {same}  max|diff|={err:.3e}")
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            cfg=np.frombuffer(bytes(out["cfg"]), dtype=np.uint8),
            seed=np.uint64(spec["seed"]), n_episodes=np.int64(n_ep), actions=out["actions"],
            reset_obs=out["ref_reset"], obs=out["ref_obs"], rew=out["ref_rew"],
            done=np.array(out["ref_done"], np.uint8), final_state=out["ref_state"])
    with open(os.path.join(outdir, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
