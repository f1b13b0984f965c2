/*
 * mbt_b200.h -- C ABI of libmbt_b200.so: the B200 (sm_100a) implementation of the ONE hot
 * path of JJJerome/mbt_gym, `TradingEnvironment.step()` / `reset()` and the rollout loop
 * around them.  Plain C, no torch / C++ types: pointers, sizes, PODs.
 *
 * The reference has no FFI; its "plugin API" for this path is the gym.Env surface of
 * mbt_gym/gym/TradingEnvironment.py.  Each entry point below names the reference code it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer would add to the reference.
 *
 *   mbt_create      <- TradingEnvironment.__init__            TradingEnvironment.py:27-94
 *   mbt_seed        <- TradingEnvironment.seed                TradingEnvironment.py:345-348
 *   mbt_reset       <- TradingEnvironment.reset/initial_state TradingEnvironment.py:96-101,131-140
 *   mbt_step        <- TradingEnvironment.step                TradingEnvironment.py:103-110
 *                      (+ everything it calls: ModelDynamics.py:108-131,262-267;
 *                       arrival_models.py:54-56,110-123; fill_probability_models.py:28-34,57-58,82,113;
 *                       midprice_models.py:60-65,97-105,140-143,354-369; price_impact_models.py:88-92;
 *                       RewardFunctions.py:23-33,55-70,96-109,128-138)
 *   mbt_get_state   <- TradingEnvironment.state (property)    TradingEnvironment.py:142-144
 *   mbt_set_state   <- `env.model_dynamics.state = ...`        ModelDynamics.py:39 (test injection / resume)
 *   mbt_reward_eval <- RewardFunction.calculate               RewardFunctions.py:8-13
 *   mbt_rollout     <- generate_trajectory + summary table    gym/helpers/generate_trajectory.py:8-38,
 *                                                             gym/helpers/plotting.py:94-108
 *
 * Conventions
 *   - every function returns 0 (MBT_OK) or a negative MBT_E_* code and never throws;
 *     mbt_last_error() returns a thread-local message for the last failure.
 *   - one handle = one CUDA device = one stream.  A handle is NOT thread-safe (the reference
 *     is single-threaded per env); different handles may be driven from different threads.
 *   - the handle owns its device state (structure-of-arrays) and pinned staging;
 *     the caller owns every buffer it passes in.  `mem` says where a caller buffer lives.
 *   - layouts of caller buffers are the reference's: actions (N, A) row-major, observations
 *     (N, D) row-major, rewards (N,), in the handle's precision (float64 or float32).
 *   - there is NO CPU implementation behind this ABI.  If no CUDA device is usable, mbt_create
 *     fails with MBT_E_CUDA.
 */
#ifndef MBT_B200_H
#define MBT_B200_H

#include "mbt_rtc.h" /* <stddef.h>, <stdint.h> (or their NVRTC stand-ins) */

#ifdef __cplusplus
extern "C" {
#endif

#define MBT_ABI_VERSION 3

/* error codes */
#define MBT_OK 0
#define MBT_E_INVALID_ARG (-1)
#define MBT_E_CUDA (-2)
#define MBT_E_STATE (-3)
#define MBT_E_UNSUPPORTED (-4)
#define MBT_E_NOMEM (-5)
#define MBT_E_NCCL (-6)

/* where a caller buffer lives */
#define MBT_MEM_HOST 0   /* pageable or pinned host memory; copies are inside the call      */
#define MBT_MEM_DEVICE 1 /* device memory on the handle's device; call is enqueue-only      */

/* arithmetic type of the whole path */
#define MBT_F64 0 /* the reference's type (every reference array is float64)               */
#define MBT_F32 1 /* opt-in fast mode, same formulas evaluated in float                     */

#define MBT_IO_SAME 0
#define MBT_IO_F32 1

/* ModelDynamics.py */
#define MBT_DYN_LIMIT 0            /* LimitOrderModelDynamics              :87-131  A=2 */
#define MBT_DYN_SPEED 1            /* TradinghWithSpeedModelDynamics       :243-275 A=1 */
#define MBT_DYN_AT_TOUCH 2         /* AtTheTouchModelDynamics              :134-176 A=2 */
#define MBT_DYN_LIMIT_AND_MARKET 3 /* LimitAndMarketOrderModelDynamics     :179-240 A=4 */

/* midprice_models.py */
#define MBT_MID_CONSTANT 0 /* ConstantMidpriceModel                :12-33   */
#define MBT_MID_BM 1       /* BrownianMotionMidpriceModel          :36-68   */
#define MBT_MID_GBM 2      /* GeometricBrownianMotionMidpriceModel :71-111  */
#define MBT_MID_OU 3       /* OuMidpriceModel                      :114-146 */
#define MBT_MID_BM_JUMP 4  /* BrownianMotionJumpMidpriceModel      :193-230 (jumps on the agent's own fills) */
#define MBT_MID_OU_JUMP 5  /* OuJumpMidpriceModel                  :233-273 */
#define MBT_MID_HESTON 6   /* HestonMidpriceModel                  :322-372 state = (price, variance): D grows by 1 */

/* arrival_models.py */
#define MBT_ARR_NONE 0
#define MBT_ARR_POISSON 1           /* PoissonArrivalModel          :33-56  p = lambda*dt         */
#define MBT_ARR_POISSON_NONLINEAR 2 /* PoissonArrivalNonLinearModel :59-83  p = 1-exp(-lambda*dt) */
#define MBT_ARR_HAWKES 3            /* HawkesArrivalModel           :86-126 state = (lam_b, lam_a) */

/* fill_probability_models.py */
#define MBT_FILL_NONE 0
#define MBT_FILL_EXPONENTIAL 1 /* ExponentialFillFunction :42-65  p = exp(-kappa*depth), per trajectory and side */
/* The next two are implemented AS WRITTEN in the reference: `np.max(depths, 0)` there is a reduction over the
 * TRAJECTORY axis (not an elementwise maximum with 0), so the fill probability of a step is a function of the
 * deepest quote of the whole batch -- one value per side (Power), or one value for both sides (Triangular, whose
 * outer `np.max(..., 0)` reduces the two sides as well).  The step therefore has a batch reduction in front of it
 * (mbt_fill_batch_kernel).  "Batch" = the trajectories of this handle, like one worker of the reference's
 * MultiprocessTradingEnv -- or, once the handle has joined a group (mbt_group_create), the trajectories of ALL ranks: the
 * shard maxima are all-reduced (max) over NCCL between the reduction and the step. */
#define MBT_FILL_TRIANGULAR 2 /* TriangularFillFunction  :68-91  p = max_side(1 - max_traj(depth)/max_fill_depth)        */
#define MBT_FILL_POWER 3      /* PowerFillFunction       :94-123 p_side = 1/(1 + (multiplier*max_traj(depth))^exponent)  */
/* ExogenousMmFillProbabilityModel :126-170: p = base * exp(-kappa * (depth - d)) beyond the exogenous best depth d of the
 * side, 1 at or inside it.  The model owns two state columns (one per side) that come after the arrival model's; in the
 * reference they never leave their initial value -- update() advances the two depth processes but never copies their
 * state (:168-170) -- so d is a per-side constant and the two columns are constants of the observation. */
#define MBT_FILL_EXOGENOUS_MM 4

/* price_impact_models.py */
#define MBT_IMP_NONE 0
#define MBT_IMP_TEMP_PERM 1  /* TemporaryAndPermanentPriceImpact :64-96 state = (I)        */
#define MBT_IMP_TEMP_POWER 2 /* TemporaryPowerPriceImpact        :34-61 stateless          */
#define MBT_IMP_TEMP_TRANSIENT 3 /* TemporaryAndTransientPriceImpact :99-139 state = (Y)     */
#define MBT_IMP_TRANSIENT 4      /* TransientPriceImpact             :142-179 state = (Y)    */

/* RewardFunctions.py */
#define MBT_REW_PNL 0                       /* PnL                     :20-36   */
#define MBT_REW_RUNNING_INVENTORY_PENALTY 1 /* = CjCriterion           :116-146 */
#define MBT_REW_CJ_MM 2                     /* CjMmCriterion           :77-113  */
#define MBT_REW_CJ_OE 3                     /* CjOeCriterion           :39-74   */
#define MBT_REW_EXP_UTILITY 4               /* ExponentialUtility      :149-166 */

/* initial inventory (TradingEnvironment.py:270-281) */
#define MBT_Q0_CONST 0       /* int, or the value a callable returned on the host            */
#define MBT_Q0_UNIFORM_INT 1 /* tuple (lo, hi): rng.integers(lo, hi) per trajectory          */
#define MBT_Q0_PER_TRAJ 2    /* a callable that returned one value per trajectory (:275-279 assigns whatever the
                                callable returns to the inventory column); only through mbt_reset_args          */

#define MBT_MAX_ACTION_DIM 4
#define MBT_MAX_OBS_DIM 8

/*
 * Everything the constructors of the reference's model-builder classes hold, flattened.
 * Field comments give the reference attribute.  All reals are float64 on the ABI regardless
 * of `precision`; the handle converts once.
 */
typedef struct mbt_config {
    int32_t struct_size; /* = sizeof(mbt_config), checked */
    int32_t precision;   /* MBT_F64 | MBT_F32 */
    int64_t num_trajectories; /* trajectories held by THIS handle (env.num_trajectories / n_gpus) */
    int64_t traj_offset;      /* global id of local trajectory 0: RNG counters use offset+i      */
    int32_t n_steps;          /* env.n_steps */
    int32_t dynamics, midprice, arrival, fill, impact, reward; /* MBT_DYN_* ... MBT_REW_* */

    double terminal_time; /* env.terminal_time */
    double step_size;     /* env.step_size = terminal_time / n_steps */
    double start_time;    /* default t0 (already quantised, TradingEnvironment.py:266-268) */
    double initial_cash;  /* env.initial_cash */
    int32_t q0_mode;      /* MBT_Q0_* */
    int32_t _pad0;
    double q0_const;
    int64_t q0_lo, q0_hi; /* MBT_Q0_UNIFORM_INT: integers in [lo, hi) */
    double max_inventory; /* env.max_inventory */
    double max_cash;      /* env.max_cash */

    /* midprice model */
    double mid_initial; /* initial_price */
    double mid_drift;   /* drift */
    double mid_vol;     /* volatility */
    double mid_step;    /* midprice_model.step_size */
    double ou_level;    /* mean_reversion_level */
    double ou_speed;    /* mean_reversion_speed */
    double mid_jump;    /* jump_size (jump models) */
    /* HestonMidpriceModel (drift above) */
    double heston_speed;  /* volatility_mean_reversion_rate  */
    double heston_level;  /* volatility_mean_reversion_level */
    double heston_corr;   /* weiner_correlation              */
    double heston_volvol; /* volatility_of_volatility        */
    double heston_var0;   /* initial_variance                */

    /* arrival model */
    double arr_rate[2];  /* Poisson intensity, or Hawkes baseline_arrival_rate (bid, ask) */
    double arr_step;     /* arrival_model.step_size */
    double hawkes_jump;  /* jump_size */
    double hawkes_speed; /* mean_reversion_speed */

    /* fill model */
    double fill_exponent;   /* ExponentialFillFunction.fill_exponent / PowerFillFunction.fill_exponent */
    double fill_max_depth;  /* TriangularFillFunction.max_fill_depth */
    double fill_multiplier; /* PowerFillFunction.fill_multiplier */
    double fill_base;       /* ExogenousMmFillProbabilityModel.base_fill_probability */
    double fill_depth0[2];  /* ... initial state of its (bid, ask) exogenous best-depth processes */

    /* price impact model */
    double imp_temp;     /* temporary_impact_coefficient */
    double imp_perm;     /* permanent_impact_coefficient */
    double imp_exponent; /* temporary_impact_exponent (TemporaryPowerPriceImpact) */
    double imp_step;     /* price_impact_model.step_size */
    double imp_transient;  /* transient_impact_coefficient  (kappa in Neuman-Voss 2022) */
    double imp_resilience; /* resilience_coefficient        (rho)                        */
    double imp_kernel;     /* linear_kernel_coefficient     (gamma)                      */
    double imp_initial;    /* initial_transient_impact      (y)                          */

    /* AtTheTouch / LimitAndMarket */
    double half_spread; /* fixed_market_half_spread */

    /* reward function */
    double rew_phi;           /* per_step_inventory_aversion */
    double rew_alpha;         /* terminal_inventory_aversion */
    double rew_exponent;      /* inventory_exponent */
    double rew_terminal_time; /* reward_function.terminal_time */
    double rew_risk_aversion; /* ExponentialUtility.risk_aversion */

    /* normalisation (TradingEnvironment.py:112-129,180-194) */
    int32_t normalise_action, normalise_obs, normalise_rewards, _pad1;
    double act_low[MBT_MAX_ACTION_DIM];  /* original_action_space.low                    */
    double act_grad[MBT_MAX_ACTION_DIM]; /* (high - low) / 2                            */
    double obs_low[MBT_MAX_OBS_DIM];     /* original_observation_space.low              */
    double obs_grad[MBT_MAX_OBS_DIM];    /* (high - low) / 2                            */
    double reward_scaling;               /* env.reward_scaling                          */

    /* ReduceStateSizeWrapper fused into the observation store (gym/wrappers.py:10-43): bit d set = column d of the
     * (normalised) observation is emitted; 0 = all D columns.  Emitted rows are (N, popcount) row-major. */
    uint32_t obs_select;
    /* element type of the CALLER's action / observation / reward buffers of mbt_reset and mbt_step:
     * MBT_IO_SAME = the handle's precision; MBT_IO_F32 = float32 buffers over float64 arithmetic (values converted
     * with round-to-nearest at the boundary; SB3 keeps float32 buffers anyway and the host path moves half the bytes).
     * State, rollout outputs, get/set_state and checkpoints stay in the handle's precision. */
    uint32_t io_precision;
} mbt_config;

/* Per-reset overrides (start_time / initial_inventory may be callables on the host). */
typedef struct mbt_reset_args {
    double start_time; /* already quantised */
    int32_t q0_mode;
    int32_t _pad;
    double q0_const;
    int64_t q0_lo, q0_hi;
    const double *q0_values; /* MBT_Q0_PER_TRAJ: HOST pointer to num_trajectories float64 values (else ignored) */
} mbt_reset_args;

/* On-device policies for the fused rollout (mbt_gym/agents/BaselineAgents.py).  The host facade computes
 * every policy constant with the reference agent's own Python expressions and passes the results. */
#define MBT_POL_FIXED 0              /* FixedActionAgent / FixedSpreadAgent  :25-42   */
#define MBT_POL_AVELLANEDA_STOIKOV 1 /* AvellanedaStoikovAgent               :52-83   */
#define MBT_POL_CJ_MM_TABLE 2        /* CarteaJaimungalMmAgent               :86-137: (delta_bid, delta_ask) per
                                        (decision time, inventory index) precomputed on the host with expm */
#define MBT_POL_SCHEDULE 3           /* any policy that depends on time only, e.g. CarteaJaimungalOeAgent
                                        :173-210: one action row per decision time */

typedef struct mbt_policy {
    int32_t kind;
    int32_t table_rows; /* CJ_MM_TABLE / SCHEDULE: number of decision times (>= steps to run)      */
    int32_t table_cols; /* CJ_MM_TABLE: 2*Q+1                                                      */
    int32_t inv_offset; /* CJ_MM_TABLE: Q, index = clip(Q + inventory, 0, 2Q)  BaselineAgents.py:121 */
    double fixed[MBT_MAX_ACTION_DIM]; /* FIXED: the action row, as the agent would return it        */
    double as_gamma;         /* AVELLANEDA_STOIKOV: risk_aversion                                  */
    double as_sigma_sq;      /*   volatility**2                                 BaselineAgents.py:71 */
    double as_fill_comp;     /*   2/gamma*log(1+gamma/kappa), or 2/kappa if gamma == 0       :73-79 */
    double as_terminal_time; /*   env.terminal_time                                                */
    const double *table;     /* HOST pointer.  CJ_MM_TABLE: rows*cols*2 float64 (bid, ask);
                                SCHEDULE: rows*A float64                                           */
} mbt_policy;

/* Episode summary = the reference's results table (plotting.py:96-108) as raw moments. */
typedef struct mbt_summary {
    int64_t count;        /* trajectories summed                       */
    int64_t steps;        /* env-steps per trajectory in this rollout  */
    double sum_return;    /* sum_i R_i,  R_i = sum_t reward_{i,t}      */
    double sum_return_sq; /* sum_i R_i^2                               */
    double sum_q;         /* sum_i q_T                                 */
    double sum_q_sq;      /* sum_i q_T^2                               */
    double sum_action;    /* sum_{i,t,a} action  (mean spread = 2*mean)*/
    double sum_reward_sq; /* sum_{i,t} reward^2 (per-step dispersion)  */
    int64_t clipped;      /* inventory/cash clip events (TradingEnvironment.py:283-297 prints instead) */
} mbt_summary;

typedef struct mbt_env mbt_env;

int mbt_abi_version(void);
const char *mbt_last_error(void);

/* dims implied by a config (no device needed): action dim A, observation dim D, state columns S */
int mbt_config_dims(const mbt_config *cfg, int32_t *action_dim, int32_t *obs_dim, int32_t *state_cols);
/* width of the observation rows reset/step/rollout_record emit (= obs_dim unless cfg->obs_select picks columns) */
int mbt_config_obs_out_dim(const mbt_config *cfg, int32_t *obs_out_dim);

int mbt_create(const mbt_config *cfg, int device, mbt_env **out);
int mbt_destroy(mbt_env *env);

/* Use an external stream for all subsequent work.  The value is a cudaStream_t taken literally, so NULL is the
 * CUDA legacy default stream (what torch's default stream is); MBT_OWN_STREAM switches back to the handle's own
 * non-blocking stream (the initial state). */
#define MBT_OWN_STREAM ((void *)(intptr_t)-1)
int mbt_set_stream(mbt_env *env, void *cuda_stream);
int mbt_sync(mbt_env *env);

/* Re-key the RNG and zero the step / episode counters.  (Reference: seed=0 means "unseeded";
 * that policy lives in the Python facade, here 0 is just a key.) */
int mbt_seed(mbt_env *env, uint64_t seed);

/* The key and the counters of the Philox draw contract (include/mbt_philox.h): env-steps and resets since mbt_seed.
 * mbt_set_counters lets a caller that re-creates a handle (a model attribute was edited between episodes) continue the
 * random streams where the old handle stopped, like the reference's generators do. */
int mbt_get_seed(mbt_env *env, uint64_t *seed);
int mbt_set_counters(mbt_env *env, int64_t steps_since_seed, int64_t episodes_since_seed);

/* Replace the handle's configuration IN PLACE, keeping device state, clock and counters: the reference reads its Python
 * attributes at every step, so `env.max_inventory = 5` or `env.reward_function.per_step_inventory_aversion = 0.1` take
 * effect at the next step (TradingEnvironment.py:283-297, RewardFunctions.py:60).  Allowed: everything that keeps
 * num_trajectories, traj_offset, precision, io_precision, the action / observation widths and the set of state columns;
 * anything else returns MBT_E_STATE (destroy and create a new handle between episodes instead). */
int mbt_reconfigure(mbt_env *env, const mbt_config *cfg);

/* Start an episode.  args may be NULL (config defaults).  obs_out may be NULL. */
int mbt_reset(mbt_env *env, const mbt_reset_args *args, void *obs_out, int mem);

/* One env-step for all trajectories.  actions (N,A); obs_out (N,D); rew_out (N,);
 * done_out: ONE byte, the uniform `dones[0]` (TradingEnvironment.py:218-220).  obs_out / rew_out /
 * done_out may be NULL.  MBT_E_STATE if called before mbt_reset. */
int mbt_step(mbt_env *env, const void *actions, void *obs_out, void *rew_out, uint8_t *done_out, int mem);

/* Raw (un-normalised) state, reference layout (N, D) row-major in the handle's precision. */
int mbt_get_state(mbt_env *env, void *state_out, int mem);
int mbt_set_state(mbt_env *env, const void *state_in, int mem);

/* Clock / counters of the handle (uniform over trajectories, TradingEnvironment.py:216-220).  After graph replays the
 * clock is the one of the last EAGER call; the counters include the device-resident base (reading it synchronises). */
int mbt_get_clock(mbt_env *env, double *time, int64_t *steps_this_episode, int64_t *steps_since_seed,
                  int64_t *episodes_since_seed);

/* Checkpoint / resume.  With a counter-based RNG the whole environment is the structure-of-arrays state block plus a
 * few counters (seed, clock, step / episode counters, what reset captured): save into / load from a HOST buffer of
 * mbt_checkpoint_size bytes.  A checkpoint loads only into a handle created with the same num_trajectories, precision
 * and model layout.  (The reference has no env checkpointing) */
int mbt_checkpoint_size(mbt_env *env, size_t *bytes);
int mbt_checkpoint_save(mbt_env *env, void *host_buf, size_t capacity);
int mbt_checkpoint_load(mbt_env *env, const void *host_buf, size_t bytes);

/*
 * CUDA-graph replay of whole episodes (policy network kernels + mbt_reset / mbt_step launches captured from the caller's
 * stream, e.g. with torch.cuda.graph).  The launches bake the handle's HOST counters (env-steps and resets since
 * mbt_seed -- the Philox draw indices) into their arguments, so a replayed graph would redraw the same random numbers.
 * Launches made while the handle's stream is capturing therefore also add a DEVICE-resident base to the baked counters,
 * and mbt_fold_counters -- called last inside the captured region -- enqueues a one-thread kernel that moves the host
 * counters into that base (base += host; host = 0).  Effective counter = host + base at every launch, captured or not:
 * R replays of a captured episode are bit-identical to R episodes stepped eagerly from the same seed, and eager calls
 * after the replays continue the same streams.  Requirements inside the captured region: MBT_MEM_DEVICE buffers only
 * (host-buffer calls synchronise), the handle already on the capturing stream (mbt_set_stream before capture), a
 * constant start time / initial-inventory mode, and the region must span whole episodes (the clock values are baked).
 * If the handle was used before the capture (warm-up resets / steps), call mbt_prepare_capture first, OUTSIDE the capture:
 * it moves the counters consumed so far into the device base eagerly, so the launches baked into the graph count from
 * zero; from then on eager calls also advance the device base (one extra one-thread kernel per call), which keeps eager
 * steps and graph replays interleavable in any order without reusing a draw index.  A first captured launch on a handle
 * with non-zero host counters and no mbt_prepare_capture fails with MBT_E_STATE instead of silently skipping indices.
 * (No reference counterpart: the reference has no device path.)
 */
int mbt_prepare_capture(mbt_env *env);
int mbt_fold_counters(mbt_env *env);

/* Trajectories whose inventory or cash was clipped since mbt_create (the reference prints the arrays
 * instead, TradingEnvironment.py:283-297). */
int mbt_get_clip_count(mbt_env *env, int64_t *count);

/* reward_function.calculate(current_state, action, next_state, is_terminal) on (n, D)/(n, A) buffers,
 * using the handle's reward parameters and the UNIFORM initial inventory / episode length captured at the last reset.
 * (Per-trajectory initial inventories -- MBT_Q0_UNIFORM_INT, MBT_Q0_PER_TRAJ -- belong to the handle's own trajectories, not
 * to caller rows: the Python facade refuses calculate() on explicit matrices when the captured inventories differ.) */
int mbt_reward_eval(mbt_env *env, int64_t n, const void *current_state, const void *action,
                    const void *next_state, int is_terminal, void *rew_out, int mem);

/* Fused rollout: from the CURRENT state run until the episode ends with an on-device policy, state
 * in registers, no per-step HBM traffic; writes the summary (host pointer) and, if non-NULL,
 * per-trajectory returns (N,) and terminal inventories (N,) to `mem` buffers. */
int mbt_rollout(mbt_env *env, const mbt_policy *policy, mbt_summary *summary_out, void *returns_out,
                void *terminal_q_out, int mem);

/* mbt_rollout that also records the whole trajectory, i.e. what generate_trajectory returns
 * (gym/helpers/generate_trajectory.py:13-15,19,27-29).  Buffers are TIME-MAJOR (every store coalesced):
 *   obs (steps+1, N, D), actions (steps, N, A) as the policy returned them, rewards (steps, N);
 * the reference's (N, D, T+1) / (N, A, T) / (N, 1, T) are transposed views of these.  Any of the three may be NULL.
 * `steps_capacity` is the number of steps the buffers can hold; MBT_E_INVALID_ARG if the episode needs more. */
typedef struct mbt_record {
    void *obs, *actions, *rewards;
    int64_t steps_capacity;
} mbt_record;
int mbt_rollout_record(mbt_env *env, const mbt_policy *policy, mbt_summary *summary_out, const mbt_record *record, int mem);

/*
 * Multi-GPU: one process per GPU, every process owns a contiguous shard of the trajectories
 * (mbt_config.traj_offset = global id of its first one).  A GROUP joins the handles of all ranks through NCCL
 * (libnccl.so.2 is loaded with dlopen at mbt_group_create; no NCCL = MBT_E_UNSUPPORTED, single-GPU use never needs it).
 * Replaces mbt_gym/gym/MultiprocessTradingEnv.py:72-116 (process fan-out over pipes, results concatenated by
 * flatten_multi :112-116).  All collectives run on device buffers, on the handle's stream (summary, batch-reduced fills)
 * or on the group's own stream behind an event (the optional gather), never through host memory.
 *
 *   mbt_group_unique_id   rank 0 creates the 128-byte NCCL id; the caller ships it to the other ranks (any transport)
 *   mbt_group_create      ncclCommInitRank; exchanges the shard sizes (all-gather of one int64 per rank)
 *   mbt_group_rollout     mbt_rollout on the local shard, then ncclAllReduce(sum) of the device-resident summary on the
 *                         handle's stream: summary_out is the summary of ALL trajectories of the group.  returns_local
 *                         (N_local,) and returns_all (sum of all shard sizes, global-id order) are DEVICE buffers or NULL;
 *                         the all-gather of returns is enqueued on the group's stream after the rollout, so it overlaps
 *                         whatever the caller enqueues next on the handle's stream (the next episode); mbt_group_wait
 *                         makes the handle's stream (and the host, if `host_sync`) wait for it.
 *   mbt_group_summary     all-reduce of a summary the caller already holds (e.g. from mbt_rollout_record)
 * With a group attached, mbt_step of the batch-reduced fill models (MBT_FILL_TRIANGULAR / MBT_FILL_POWER) all-reduces
 * (max) the deepest quotes of the shards before the step, so `np.max(depths, 0)` runs over the WHOLE batch and results
 * do not depend on the shard layout.
 */
#define MBT_GROUP_ID_BYTES 128
int mbt_group_unique_id(void *id_out);
int mbt_group_create(mbt_env *env, const void *id, int rank, int world);
int mbt_group_destroy(mbt_env *env);
int mbt_group_info(mbt_env *env, int32_t *rank, int32_t *world, int64_t *total_trajectories);
int mbt_group_rollout(mbt_env *env, const mbt_policy *policy, mbt_summary *summary_out, void *returns_local_out,
                      void *returns_all_out);
int mbt_group_summary(mbt_env *env, const mbt_summary *local, mbt_summary *global_out);
int mbt_group_wait(mbt_env *env, int host_sync);

/*
 * Run-time specialisation (csrc/mbt_jit.h).  mbt_create / mbt_reconfigure compile -- once per configuration KIND, cached
 * per process and on disk -- a step kernel with every model kind, reward kind, normalisation flag and column mask of the
 * handle's configuration fixed at compile time (NVRTC; the same kernel bodies and floating-point flags as the
 * ahead-of-time build, so results are bit-identical), unless one of the fully specialised ahead-of-time variants already
 * serves the configuration.  Without libnvrtc the ahead-of-time kernels run (`message` says why); MBT_JIT=0 switches the
 * specialiser off, MBT_JIT=require turns its failure into an error of mbt_create.
 */
typedef struct mbt_kernel_info {
    int32_t aot_variant;          /* index into the ahead-of-time table (0 = generic kernel with run-time switches) */
    int32_t jit_mode;             /* 0 off, 1 on (fallback allowed), 2 required */
    int32_t step_is_jit;          /* the step kernel this handle launches was specialised at run time */
    int32_t step_registers;       /* registers per thread / local-memory bytes of that kernel (0 when not JIT) */
    int32_t step_local_bytes;
    int32_t jit_from_disk_cache;  /* the cubin came from the on-disk cache (no compile in this process) */
    double jit_compile_ms;        /* time spent compiling or reading it */
    uint64_t jit_hash;            /* cache key: <library dir>/_jit_cache/<hash>.cubin */
    int32_t rollout_is_jit[5];    /* per policy kind (MBT_POL_*), [4] = the recording kernel: specialised kernels loaded so far */
    int32_t _pad;
    char message[256];            /* why the specialiser is not in use, if it should be */
} mbt_kernel_info;
int mbt_get_kernel_info(mbt_env *env, mbt_kernel_info *out);
/* Compile the specialised kernel of `cfg` into the on-disk cache WITHOUT a CUDA device (build machines): kind 0 = step,
 * 1 = rollout with `policy_kind` compiled in (or the recording kernel when `record`). */
int mbt_jit_precompile(const mbt_config *cfg, int32_t kind, int32_t policy_kind, int32_t record);

/* Histogram of the CURRENT inventory column (after a rollout or a step loop: the terminal inventories) over the integer
 * bins lo .. hi (at most 4096): what the reference's results plot bins on the host (gym/helpers/plotting.py:94-110, the
 * "inventory distribution" of the parity statistics).  counts_out (HOST, hi - lo + 3 int64): [0] = below lo,
 * [1 + (q - lo)] = round(q) == q, [hi - lo + 2] = above hi or NaN.  group_sum != 0: the counts of ALL ranks of the handle's
 * group (NCCL all-reduce on the handle's stream). */
int mbt_inventory_histogram(mbt_env *env, int64_t lo, int64_t hi, int64_t *counts_out, int group_sum);

/* Per-call statistics for bench.py: number of kernel launches issued by this handle so far, and the
 * device time (ms, CUDA events on the handle's stream) of the most recent step kernel when enabled. */
int mbt_get_launch_count(mbt_env *env, int64_t *launches);
/* enable = 1: bracket every subsequent hot-path kernel (step / rollout) with two CUDA events on the handle's stream;
 * enable = 2: ONE event before every kernel, a launch's duration being the interval to the next launch's event
 * (kernel + gap: an upper bound that perturbs a back-to-back loop half as much); 0: off.  Up to an internal ring
 * capacity.  mbt_get_kernel_times synchronises and returns the durations (call it once after the measured loop). */
int mbt_enable_timing(mbt_env *env, int enable);
int mbt_get_kernel_times(mbt_env *env, float *ms_out, int64_t capacity, int64_t *count);

/* Pinned (page-locked) host memory for caller buffers: MBT_MEM_HOST buffers allocated here are DMA'd
 * directly; any other host pointer is staged through the handle's own pinned buffers. */
int mbt_host_alloc(size_t bytes, void **out);
/* same, placed on the NUMA node of `device` (pages on the far socket are DMA-read at less than half the rate) */
int mbt_host_alloc_near(size_t bytes, int device, void **out);
int mbt_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* MBT_B200_H */
