/*
 * mbt_capi.cu -- libmbt_b200.so: handle management and the C ABI of include/mbt_b200.h.
 *
 * There is no CPU implementation in this library: every entry point that computes launches one of the
 * kernels of mbt_kernels.cuh on the handle's stream.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h> /* header-only; ranges cost nothing unless a tool (ncu --nvtx, nsys) is attached */
#define MBT_HAVE_NVTX 1
#endif

#include "mbt_b200.h"
#include "mbt_host_params.h"
#include "mbt_variants.h"
#include "mbt_kernels.cuh"
#include "mbt_jit.h"

/* ------------------------------------------------------------------ errors */
static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            char _b[512];                                                                             \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                                       \
            return fail(MBT_E_CUDA, _b);                                                              \
        }                                                                                             \
    } while (0)

/* NVTX range around an ABI call: `ncu --nvtx --nvtx-include "mbt_step/"` or an nsys timeline attribute the launches to it */
struct NvtxRange {
#ifdef MBT_HAVE_NVTX
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char *) {}
#endif
};

/* ------------------------------------------------------------------ handle */
constexpr int MBT_TIMING_RING = 8192;
constexpr int MBT_STATE_COLUMNS = 7; /* columns of the structure-of-arrays state block (DevState) */
constexpr int MBT_PIPE_CHUNKS = 16;        /* capacity */
constexpr int MBT_PIPE_CHUNKS_DEFAULT = 4;  /* measured: 2 / 4 / 8 / 16 chunks -> 0.965 / 0.931 / 0.950 / 1.026 ms per step */

struct mbt_env {
    mbt_config cfg;
    int device = 0;
    int A = 0, D = 0, S = 0;
    int Dout = 0; /* emitted observation width (D unless cfg.obs_select picks columns) */
    int sm_count = 1;
    size_t esz = 8;    /* element size of the arithmetic / state type */
    size_t io_esz = 8; /* element size of the caller's action / observation / reward buffers */
    long long N = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;

    /* device state, structure-of-arrays (one allocation, columns of N elements) */
    void *state_block = nullptr;
    void *col[MBT_STATE_COLUMNS] = {}; /* cash, inv, mid, x0, x1, q0, var */
    unsigned long long *d_clipped = nullptr;

    /* device + pinned staging for MBT_MEM_HOST calls */
    void *d_actions = nullptr, *d_obs = nullptr, *d_rew = nullptr;
    void *h_actions = nullptr, *h_obs = nullptr, *h_rew = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr; /* H2D / D2H copy engines for the pipelined host path */
    cudaEvent_t ev_in[MBT_PIPE_CHUNKS] = {}, ev_k[MBT_PIPE_CHUNKS] = {};

    /* batch reduction in front of the step (Triangular / Power fill functions): running maxima (keys), ticket, thresholds */
    void *d_fill_partial = nullptr;
    unsigned int *d_fill_ticket = nullptr;
    int fill_blocks = 1;

    /* CUDA-graph replay: device-resident base of the (step, episode) counters; effective counter = host + base.
     * `device_counters` turns on the first time a call is made while the stream is capturing, or by mbt_fold_counters */
    unsigned long long *d_counter_base = nullptr;
    bool device_counters = false;

    /* rollout scratch */
    void *d_clocks = nullptr; /* RolloutClock<T>[steps] */
    size_t clocks_cap = 0;    /* bytes */
    void *d_table = nullptr;
    size_t table_cap = 0;
    double *d_block_sums = nullptr;
    int block_sums_cap = 0;
    double *d_summary = nullptr, *h_summary = nullptr; /* MBT_SUMMARY_DOUBLES each: device-resident fold, pinned mirror */
    unsigned int *d_roll_ticket = nullptr;

    /* group of handles (one per rank / GPU) joined through NCCL: mbt_group_* */
    void *comm = nullptr; /* ncclComm_t */
    int g_rank = 0, g_world = 1;
    std::vector<long long> g_counts; /* shard sizes of all ranks */
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_rollout = nullptr, ev_gather = nullptr;
    bool gather_pending = false;
    const void *gather_src = nullptr; /* returns buffer the pending gather reads */

    /* clock (uniform over trajectories) */
    double t = 0, t0 = 0;
    int64_t k = 0, n_step = 0, n_episode = 0;
    bool started = false;
    uint64_t seed = 0;
    int q0_per_traj = 0;
    double q0_uniform = 0;

    /* run-time specialised kernels of this configuration (mbt_jit.h); NULL = the ahead-of-time table is used */
    const mbt_jit::Module *jit_step = nullptr;
    const mbt_jit::Module *jit_roll[5] = {}; /* MBT_POL_FIXED .. MBT_POL_SCHEDULE, [4] = recording */
    int jit_roll_state[5] = {};              /* 0 untried, 1 ready, -1 failed */
    std::string jit_error;                   /* why the specialiser is not in use (empty = in use or switched off) */

    bool has_l2_window = false; /* MBT_L2_PERSIST=1: this handle raised the persisting-L2 set-aside */

    /* statistics */
    int64_t launches = 0;
    int timing = 0; /* 0 off, 1 two events around every kernel, 2 one event before every kernel (interval timing) */
    std::vector<cudaEvent_t> ev0, ev1;
    int64_t timed = 0;
    bool interval_closed = false;
};

static bool stream_is_capturing_fwd(mbt_env *e);

template <typename T>
static DevState<T> dev_state(mbt_env *e) {
    DevState<T> st;
    st.cash = (T *)e->col[0];
    st.inv = (T *)e->col[1];
    st.mid = (T *)e->col[2];
    st.x0 = (T *)e->col[3];
    st.x1 = (T *)e->col[4];
    st.q0 = (T *)e->col[5];
    st.var = (T *)e->col[6];
    return st;
}

static inline unsigned grid_for(long long n) { return (unsigned)((n + MBT_BLOCK - 1) / MBT_BLOCK); }

/* blocks per SM of the batch reduction (env MBT_FILL_BLOCKS_PER_SM overrides for tuning).  Measured at N = 2^20, f64, cold cache (ncu): 1 / 2 / 4 / 8 blocks per SM -> 12.3 / 10.9 / 10.3 / 12.9 us */
static int fill_blocks_per_sm() {
    const char *v = getenv("MBT_FILL_BLOCKS_PER_SM");
    int k = v ? atoi(v) : 4;
    return k < 1 ? 1 : (k > 8 ? 8 : k);
}

/* does a step of this config start with the batch reduction of the quoted depths? */
static inline bool needs_fill_batch(const mbt_config &c) {
    return fill_is_batch(c.fill) && (c.dynamics == MBT_DYN_LIMIT || c.dynamics == MBT_DYN_LIMIT_AND_MARKET);
}

/*
 * L2 residency of the state.  The structure-of-arrays state columns are read and rewritten by EVERY step (25 MB for the
 * BASELINE market at 2^20 float64 trajectories, 42 MB with Hawkes intensities) while the caller's action / observation /
 * reward buffers only stream through; an access-policy window on the handle's stream marks the columns in use as
 * persisting in the 126 MB L2 and everything else the kernels touch as streaming.  Measured per step at N = 2^20, f64
 * (profiles/r2_session_notes.md): AS 14.7 -> 14.3 us, CjMm 16.6 -> 15.7, OE 18.3 -> 16.3, Hawkes 25.2 -> 19.1.
 * The persisting set-aside is a device-wide setting, so it is handled like any other resource the handle owns: only when
 * the columns fit (<= 40 MiB and the device's limits), only ever raised while windowed handles exist, and given back (with
 * the persisting lines) when the last of them is destroyed -- while it is raised, other L2-hungry work on the device runs
 * a few per cent slower (a 2^24-trajectory handle next to a live windowed 2^20 one: 6 %, tools/l2_window_probe.py).
 * A caller's stream the handle leaves loses the window.  MBT_L2_PERSIST=0 turns all of it off.
 */
static bool l2_persist_enabled() {
    static const bool on = [] {
        const char *v = getenv("MBT_L2_PERSIST");
        return !(v && v[0] == '0');
    }();
    return on;
}
static int g_l2_window_handles = 0; /* handles that currently hold a window (per process; handles are created / destroyed
                                       by their owning threads -- the counter is only a hint for giving the set-aside back) */

/* state columns (of the 7 in the block: cash, inventory, midprice, x0, x1, q0, variance) a step of this config touches */
static int state_columns_in_use(const mbt_config &c) {
    int n = 3;
    if (imp_has_state(c.impact)) n = 4;
    if (c.arrival == MBT_ARR_HAWKES) n = 5;
    if ((c.reward == MBT_REW_CJ_MM || c.reward == MBT_REW_CJ_OE) && c.q0_mode != MBT_Q0_CONST) n = 6;
    if (c.midprice == MBT_MID_HESTON) n = 7;
    return n;
}

static void apply_l2_window(mbt_env *e, cudaStream_t stream) {
    if (!l2_persist_enabled() || !stream) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
    const size_t col_bytes = (((size_t)e->N * e->esz) + 255) & ~(size_t)255;
    const size_t bytes = col_bytes * (size_t)state_columns_in_use(e->cfg);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    if (max_persist > 0 && bytes <= (size_t)max_persist && bytes <= (size_t)max_window && bytes <= ((size_t)40 << 20)) {
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        if (cur < bytes) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
        v.accessPolicyWindow.base_ptr = e->state_block;
        v.accessPolicyWindow.num_bytes = bytes;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (!e->has_l2_window) {
            e->has_l2_window = true;
            g_l2_window_handles += 1;
        }
    } /* else: num_bytes = 0 clears a window this handle may have set before on this stream */
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
}

/* a caller's stream the handle leaves (mbt_set_stream, mbt_destroy) must not keep a window over the handle's memory */
static void clear_l2_window(mbt_env *e, cudaStream_t stream) {
    if (!l2_persist_enabled() || !stream || stream == e->own_stream) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
}

static int timing_begin(mbt_env *e) {
    if (!e->timing || e->timed >= MBT_TIMING_RING) return MBT_OK;
    if ((int64_t)e->ev0.size() <= e->timed) {
        cudaEvent_t a, b;
        CU(cudaEventCreate(&a));
        CU(cudaEventCreate(&b));
        e->ev0.push_back(a);
        e->ev1.push_back(b);
    }
    CU(cudaEventRecord(e->ev0[e->timed], e->stream));
    return MBT_OK;
}
static int timing_end(mbt_env *e) {
    if (!e->timing || e->timed >= MBT_TIMING_RING) return MBT_OK;
    if (e->timing == 1) CU(cudaEventRecord(e->ev1[e->timed], e->stream));
    e->timed += 1;
    return MBT_OK;
}

/*
 * NUMA placement of pinned host buffers.  cudaHostAlloc places pages on the NUMA node of the calling thread; on a
 * two-socket host a buffer on the far socket is read by the GPU at ~16-25 GB/s instead of ~45-55 GB/s (measured:
 * profiles/r1_pcie_numa.md).  While it allocates, the library therefore narrows the calling thread's affinity to the
 * CPUs of the GPU's own node (intersected with what the process is allowed to use) and restores it afterwards.
 */
static int gpu_numa_node(int device) {
    char bdf[64] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static bool node_cpus(int node, cpu_set_t *set) {
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a, b;
    bool any = false;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &b) != 1) break;
            c = fgetc(f);
        }
        for (int i = a; i <= b && i < CPU_SETSIZE; ++i) {
            CPU_SET(i, set);
            any = true;
        }
        if (c != ',') break;
    }
    fclose(f);
    return any;
}

struct ScopedNumaAffinity {
    cpu_set_t old_set;
    bool active = false;
    explicit ScopedNumaAffinity(int device) {
        if (getenv("MBT_NO_NUMA_BIND")) return;
        const int node = gpu_numa_node(device);
        cpu_set_t want, allowed;
        if (node < 0 || !node_cpus(node, &want)) return;
        if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
        old_set = allowed;
        cpu_set_t both;
        CPU_AND(&both, &want, &allowed);
        if (CPU_COUNT(&both) == 0) return;
        if (sched_setaffinity(0, sizeof both, &both) == 0) active = true;
    }
    ~ScopedNumaAffinity() {
        if (active) sched_setaffinity(0, sizeof old_set, &old_set);
    }
};

/* is this host pointer page-locked (DMA-able without staging)? */
static bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

/*
 * Parallel memcpy between pageable caller memory and pinned staging.  A single thread copies at ~10 GB/s, far below the
 * PCIe rate, and creating threads per call cost more than the copy (round 1: up to 8 std::threads per mbt_step -- 0.9 ms
 * of a 2.0 ms step for an ordinary NumPy action array).  One process-wide pool of persistent workers (created on first
 * use, never joined: they sleep on a condition variable) splits a copy into equal parts; the calling thread takes one part
 * itself.  Calls from different handles' threads serialise on the pool's mutex (the copies are memory-bound anyway).
 * MBT_COPY_THREADS overrides the worker count (default: a quarter of the hardware threads, 2..8).
 */
class CopyPool {
public:
    static CopyPool &get() {
        static CopyPool *pool = new CopyPool(); /* leaked on purpose: workers may outlive static destruction */
        return *pool;
    }
    void copy(void *dst, const void *src, size_t bytes) {
        const size_t min_part = 256u << 10;
        size_t parts = std::min<size_t>(workers_.size() + 1, (bytes + min_part - 1) / min_part);
        if (parts <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t per = ((bytes / parts) + 63) & ~(size_t)63;
        std::unique_lock<std::mutex> call(call_mu_); /* one parallel copy at a time */
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; per_ = per;
            next_part_ = 1; /* part 0 is the caller's */
            parts_ = parts;
            pending_ = parts - 1;
            generation_ += 1;
        }
        cv_work_.notify_all();
        memcpy(dst, src, std::min(per, bytes));
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = (int)std::min(8u, std::max(2u, hw / 4));
        if (const char *v = getenv("MBT_COPY_THREADS")) n = std::max(1, std::min(64, atoi(v)));
        for (int i = 0; i + 1 < n; ++i) {
            workers_.emplace_back([this] { run(); });
            workers_.back().detach();
        }
    }
    void run() {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            while (next_part_ < parts_) {
                const size_t part = next_part_++;
                const size_t off = part * per_;
                char *d = dst_;
                const char *s = src_;
                const size_t len = off < bytes_ ? std::min(per_, bytes_ - off) : 0;
                lk.unlock();
                if (len) memcpy(d + off, s + off, len);
                lk.lock();
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_work_, cv_done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0, next_part_ = 0, parts_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
};

static void par_memcpy(void *dst, const void *src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }

/* ------------------------------------------------------------------ run-time specialisation (mbt_jit.h) */
/* configurations whose ahead-of-time variant still reads model kinds or normalisation flags at run time */
static bool wants_jit(const mbt_config &c) {
    const int v = variant_of(c);
    return v == 0 || v == 4 || v == 6 || v == 9;
}

/* (re)select the step kernel of the handle's configuration: called by mbt_create and mbt_reconfigure, never inside a
 * stream capture.  Returns an error only when MBT_JIT=require. */
static int jit_select_step(mbt_env *e) {
    e->jit_step = nullptr;
    for (int i = 0; i < 5; ++i) { e->jit_roll[i] = nullptr; e->jit_roll_state[i] = 0; }
    e->jit_error.clear();
    if (mbt_jit::mode() == 0 || !wants_jit(e->cfg)) return MBT_OK;
    const mbt_jit::Key key = mbt_jit::key_of(e->cfg, e->io_esz == 8, mbt_jit::STEP, 0, 0);
    std::string err;
    int rc = mbt_jit::module_of(key, &e->jit_step, err);
    if (rc) {
        e->jit_step = nullptr;
        e->jit_error = err;
        if (mbt_jit::mode() == 2) return fail(rc, "MBT_JIT=require: " + err);
    }
    return MBT_OK;
}

/* the specialised rollout kernel for policy slot `slot` (lazily; not while capturing); *out = NULL -> ahead-of-time table */
static int jit_rollout_module(mbt_env *e, int slot, const mbt_jit::Module **out) {
    *out = nullptr;
    if (slot < 0 || slot > 4 || mbt_jit::mode() == 0 || !wants_jit(e->cfg)) return MBT_OK;
    if (e->jit_roll_state[slot] == 0 && !stream_is_capturing_fwd(e)) {
        const mbt_jit::Key key = mbt_jit::key_of(e->cfg, e->esz == 8, mbt_jit::ROLLOUT, slot == 4 ? -1 : slot, slot == 4);
        std::string err;
        int rc = mbt_jit::module_of(key, &e->jit_roll[slot], err);
        e->jit_roll_state[slot] = rc ? -1 : 1;
        if (rc) {
            e->jit_roll[slot] = nullptr;
            e->jit_error = err;
            if (mbt_jit::mode() == 2) return fail(rc, "MBT_JIT=require: " + err);
        }
    }
    if (e->jit_roll_state[slot] == 1) *out = e->jit_roll[slot];
    return MBT_OK;
}

/* ------------------------------------------------------------------ launchers */
/* is the handle's stream being captured into a CUDA graph right now? */
static bool stream_is_capturing(mbt_env *e) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(e->stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st == cudaStreamCaptureStatusActive;
}

static bool stream_is_capturing_fwd(mbt_env *e) { return stream_is_capturing(e); }

/*
 * Counters of the draw contract.  Effective counter of a launch = host counter baked into its arguments + device-resident
 * base (read by the kernel).  Plain eager use: base unused (NULL), host counters advance.  Once the handle has been used
 * under stream capture (or mbt_prepare_capture / mbt_fold_counters was called) it is in DEVICE-COUNTER mode: captured
 * launches bake host counters that count from the start of the captured region; mbt_fold_counters (captured last) adds the
 * region's totals to the base at every replay; and every EAGER call folds its own advance into the base right after its
 * launch (fold_eager), keeping the host counters at zero outside a capture -- so graph replays and eager calls can be
 * interleaved in any order without reusing or skipping a draw index.
 */
static int counter_base_for_launch(mbt_env *e, const unsigned long long **out) {
    if (!e->device_counters && stream_is_capturing(e)) {
        if (e->n_step != 0 || e->n_episode != 0)
            return fail(MBT_E_STATE,
                        "the handle was used before this stream capture began: call mbt_prepare_capture() (env.prepare_capture()) "
                        "before capturing, so the graph's baked draw counters start from zero");
        e->device_counters = true;
    }
    *out = e->device_counters ? e->d_counter_base : nullptr;
    return MBT_OK;
}

/* device-counter mode, eager call: move what the call just consumed into the device base (one-thread kernel) */
static int fold_eager(mbt_env *e) {
    if (!e->device_counters || stream_is_capturing(e)) return MBT_OK;
    if (e->n_step == 0 && e->n_episode == 0) return MBT_OK;
    mbt_fold_counters_kernel<<<1, 1, 0, e->stream>>>(e->d_counter_base, (unsigned long long)e->n_step, (unsigned long long)e->n_episode);
    CU(cudaGetLastError());
    e->launches += 1;
    e->n_step = 0;
    e->n_episode = 0;
    return MBT_OK;
}

/* MBT_PDL=0 disables programmatic dependent launch of consecutive step kernels (default: on) */
static bool use_pdl() {
    static const bool on = [] {
        const char *v = getenv("MBT_PDL");
        return !(v && v[0] == '0');
    }();
    return on;
}

template <typename T, typename E, class V, bool VEC>
static void launch_step_k(mbt_env *e, const StepArgs<T, E> &g, bool allow_pdl) {
    if (!allow_pdl || !use_pdl()) {
        mbt_step_kernel<T, E, V, VEC><<<grid_for(g.n), MBT_BLOCK, 0, e->stream>>>(g);
        return;
    }
    /* programmatic stream serialization: this kernel may begin (up to its griddepcontrol.wait) while the previous
     * kernel in the stream drains -- the next step's launch latency and Philox prologue hide behind this step's tail */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_for(g.n));
    cfg.blockDim = dim3(MBT_BLOCK);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mbt_step_kernel<T, E, V, VEC>, g);
}

/* the run-time specialised step kernel of the handle's configuration (same argument block, same launch attributes) */
template <typename T, typename E>
static void launch_step_jit(mbt_env *e, const StepArgs<T, E> &g, bool vec, bool allow_pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_for(g.n));
    cfg.blockDim = dim3(MBT_BLOCK);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const bool pdl = allow_pdl && use_pdl();
    cfg.attrs = pdl ? attr : nullptr;
    cfg.numAttrs = pdl ? 1 : 0;
    void *args[1] = {const_cast<StepArgs<T, E> *>(&g)};
    cudaLaunchKernelExC(&cfg, (const void *)(vec ? e->jit_step->k0 : e->jit_step->k1), args);
}

template <typename T, typename E, class V>
static void launch_step_v(mbt_env *e, const StepArgs<T, E> &g, bool vec, bool allow_pdl) {
    if (vec)
        launch_step_k<T, E, V, true>(e, g, allow_pdl);
    else
        launch_step_k<T, E, V, false>(e, g, allow_pdl);
}

/* may the step kernel use whole-row vector accesses on the caller's buffers?  (see load_row / store_row:
 * actions A=2 -> 2-element, A=4 -> 4-element vectors; observations D=4 -> 4-element, D=6 -> 2-element) */
template <typename E>
static bool rows_vector_aligned(const mbt_env *e, const void *actions, const void *obs) {
    const size_t need_a = (size_t)(e->A == 2 ? 2 : e->A == 4 ? 4 : 1) * sizeof(E);
    const size_t need_o = (size_t)(e->Dout == 4 ? 4 : (e->Dout == 6 || e->Dout == 2) ? 2 : 1) * sizeof(E);
    return ((uintptr_t)actions % need_a) == 0 && (!obs || ((uintptr_t)obs % need_o) == 0);
}

/* launch the step kernel on rows [r0, r0+n) of the batch (base pointers address row 0) */
template <typename T, typename E>
static int launch_step_rows(mbt_env *e, const StepParams<T> &p, const StepClock<T> &ck, const void *actions, void *obs,
                            void *rew, long long r0, long long n, bool allow_pdl) {
    const mbt_config &c = e->cfg;
    StepArgs<T, E> g;
    g.p = p;
    g.ck = ck;
    g.st = dev_state<T>(e);
    g.st.cash += r0; g.st.inv += r0; g.st.mid += r0; g.st.x0 += r0; g.st.x1 += r0; g.st.q0 += r0; g.st.var += r0;
    g.actions = (const E *)actions + r0 * e->A;
    g.obs = obs ? (E *)obs + r0 * e->Dout : nullptr;
    g.rew = rew ? (E *)rew + r0 : nullptr;
    g.n = n;
    g.keys = mbt_philox_expand(e->seed);
    g.traj_offset = (unsigned long long)c.traj_offset + (unsigned long long)r0;
    g.n_step = (unsigned long long)e->n_step;
    g.clipped = e->d_clipped;
    g.fill_cells = (unsigned long long *)e->d_fill_partial;
    g.fill_ticket = e->d_fill_ticket;
    {
        int rcb = counter_base_for_launch(e, &g.counter_base);
        if (rcb) return rcb;
    }
    if (g.counter_base) allow_pdl = false; /* the base is written by a kernel: order behind it */
    const bool vec = rows_vector_aligned<E>(e, g.actions, g.obs);
    if (e->jit_step) {
        launch_step_jit<T, E>(e, g, vec, allow_pdl);
    } else {
        switch (variant_of(c)) {
#define X(id, ...) case id: launch_step_v<T, E, __VA_ARGS__>(e, g, vec, allow_pdl); break;
            MBT_FOR_EACH_VARIANT(X)
#undef X
        }
    }
    CU(cudaGetLastError());
    e->launches += 1;
    return MBT_OK;
}

static int group_allreduce_max_u64(mbt_env *e, unsigned long long *dev, int count);

/* the batch reduction of the quoted depths (mbt_fill_batch_kernel) over ALL rows of the action matrix; the maxima stay
 * in e->d_fill_partial (keys) for the step kernel that follows */
template <typename T, typename E>
static int launch_fill_batch(mbt_env *e, const StepParams<T> &p, const void *actions, bool allow_pdl) {
    FillBatchArgs<T, E> g;
    g.p = p;
    g.actions = (const E *)actions;
    g.n = e->N;
    g.cells = (unsigned long long *)e->d_fill_partial;
    const unsigned blocks = std::min<unsigned>(grid_for(e->N), (unsigned)e->fill_blocks);
    /* rows of 2 or 4 elements whose base is aligned to a 2-element vector: one vector load per row */
    const bool vec = (e->A == 2 || e->A == 4) && ((uintptr_t)actions % (2 * sizeof(E))) == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(MBT_BLOCK);
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const bool pdl = allow_pdl && use_pdl();
    cfg.attrs = pdl ? attr : nullptr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (vec)
        cudaLaunchKernelEx(&cfg, mbt_fill_batch_kernel<T, E, true>, g);
    else
        cudaLaunchKernelEx(&cfg, mbt_fill_batch_kernel<T, E, false>, g);
    CU(cudaGetLastError());
    e->launches += 1;
    if (e->comm) {
        /* group of handles: np.max(depths, 0) runs over the trajectories of ALL ranks -- NCCL all-reduce (max) of the two
         * keys on the handle's stream, between the reduction and the step */
        int rc = group_allreduce_max_u64(e, g.cells, 2);
        if (rc) return rc;
    }
    return MBT_OK;
}

static void advance_clock(mbt_env *e, double t_next) {
    e->t = t_next;
    e->k += 1;
    e->n_step += 1;
}

template <typename T, typename E>
static int do_step_device(mbt_env *e, const void *actions, void *obs, void *rew, uint8_t *done_out) {
    const mbt_config &c = e->cfg;
    const double t_next = e->t + c.step_size; /* state[:, TIME] += step_size   TradingEnvironment.py:216 */
    const StepParams<T> p = mbt_make_params<T>(c, e->t0, e->q0_per_traj, e->q0_uniform);
    const StepClock<T> ck = mbt_make_clock<T>(c, e->t, t_next, e->t0);
    int rc = timing_begin(e);
    if (rc) return rc;
    const bool batch = needs_fill_batch(c);
    if (batch) {
        /* (device-counter mode: the previous launch may be the one-thread fold kernel, which is not PDL-aware) */
        rc = launch_fill_batch<T, E>(e, p, actions, /*allow_pdl=*/!e->device_counters);
        if (rc) return rc;
    }
    /* behind an NCCL all-reduce (group of handles) the step is launched plainly */
    rc = launch_step_rows<T, E>(e, p, ck, actions, obs, rew, 0, e->N, /*allow_pdl=*/!(batch && e->comm));
    if (rc) return rc;
    rc = timing_end(e);
    if (rc) return rc;
    advance_clock(e, t_next);
    if (done_out) *done_out = (uint8_t)ck.done;
    return fold_eager(e);
}

/* host-buffer path: MBT_HOST_PATH=zerocopy lets the kernel access pinned host memory directly; default = DMA pipeline */
static bool host_path_zero_copy() {
    static const bool zc = [] {
        const char *v = getenv("MBT_HOST_PATH");
        return v && strcmp(v, "zerocopy") == 0;
    }();
    return zc;
}

/* MBT_PIPE_CHUNKS set = equal chunks; unset (or MBT_PIPE_SCHEDULE=geo) = the geometric schedule of do_step_host_pipelined */
static bool pipe_schedule_geometric() {
    static const bool geo = [] {
        const char *sch = getenv("MBT_PIPE_SCHEDULE");
        if (sch) return strcmp(sch, "geo") == 0;
        return getenv("MBT_PIPE_CHUNKS") == nullptr;
    }();
    return geo;
}

/* number of pipeline chunks of the host-buffer path (env MBT_PIPE_CHUNKS overrides for tuning; 1..16, default 4) */
static int pipe_chunks() {
    static const int n = [] {
        const char *v = getenv("MBT_PIPE_CHUNKS");
        int k = v ? atoi(v) : MBT_PIPE_CHUNKS_DEFAULT;
        return k < 1 ? 1 : (k > MBT_PIPE_CHUNKS ? MBT_PIPE_CHUNKS : k);
    }();
    return n;
}

/*
 * Host-buffer step: the batch is cut into row chunks and each chunk flows H2D(actions) -> kernel -> D2H(obs, rewards)
 * on three streams, so the two copy engines (PCIe is full duplex) and the SMs overlap; the call returns when the last
 * chunk's results are in the caller's buffers.  `act_src`, `obs_dst`, `rew_dst` are pinned (caller's own pinned
 * buffers, or the handle's staging).  A pageable action array (an ordinary NumPy array: what `agent.get_action(obs)` returns)
 * takes part in the pipeline chunk by chunk: `act_pageable` is copied into the pinned `act_src` by the persistent copy pool
 * right before each chunk's upload, overlapped with the transfers of the earlier chunks (round 1 staged the whole array in
 * front of the pipeline with threads created per call: 2.0 ms per step, now 1.4-1.5; pinned actions: 1.0).  Pageable OUTPUT
 * arrays (`copy_outputs=True`) are filled after the pipeline: writing 42 MB of never-touched pages is page-fault-bound
 * (3-5 ms) however it is chunked -- the default pooled pinned outputs avoid it.
 */
template <typename T, typename E>
static int do_step_host_pipelined(mbt_env *e, const void *act_src, void *obs_dst, void *rew_dst, uint8_t *done_out,
                                  const void *act_pageable) {
    const mbt_config &c = e->cfg;
    const double t_next = e->t + c.step_size;
    const StepParams<T> p = mbt_make_params<T>(c, e->t0, e->q0_per_traj, e->q0_uniform);
    const StepClock<T> ck = mbt_make_clock<T>(c, e->t, t_next, e->t0);
    const long long N = e->N;
    const bool batch = needs_fill_batch(c); /* the reduction needs every action row on the device first */
    /* chunk boundaries: equal chunks (MBT_PIPE_CHUNKS=k), or by default a GEOMETRIC schedule 1/16, 1/16, 1/8, 1/4, 1/2 of the
     * rows -- the device-to-host copies (the longer direction: D+1 against A elements per row) start after 1/16 of the
     * action upload instead of 1/4, and every later chunk's upload + kernel finish before the D2H engine drains the one
     * before, so the D2H engine stays busy from ~25 us after the call to its end */
    long long bounds[MBT_PIPE_CHUNKS + 1];
    int chunks = 1;
    bounds[0] = 0;
    if (N >= (1 << 17) && !batch && act_pageable) {
        /* pageable action array: the host copy into pinned staging (~20 GB/s on the measured hosts) is the slowest stage, so
         * equal chunks -- chunk k+1 is staged while chunk k is on the wire, and the call ends one chunk after the last host
         * copy.  Measured per step at N = 2^20: 16 chunks 1.39-1.55 ms, 8 chunks 1.61-1.64 ms (pinned actions: 1.02 ms;
         * round 1's staging in front of the pipeline: 1.95-2.05 ms). */
        chunks = MBT_PIPE_CHUNKS;
        const long long rows = ((N + chunks - 1) / chunks + 255) & ~255ll;
        for (int i = 1; i <= chunks; ++i) bounds[i] = std::min(N, rows * i);
    } else if (N >= (1 << 17) && !batch) {
        if (pipe_schedule_geometric()) {
            static const int geo3 = getenv("MBT_PIPE_GEO3") != nullptr;
            const int shifts5[5] = {4, 3, 2, 1, 0}, shifts3[3] = {3, 1, 0}; /* cumulative end of chunk i = N >> shift */
            const int *shifts = geo3 ? shifts3 : shifts5;
            chunks = geo3 ? 3 : 5;
            for (int i = 0; i < chunks; ++i) bounds[i + 1] = i == chunks - 1 ? N : ((N >> shifts[i]) + 255) & ~255ll;
        } else {
            chunks = pipe_chunks();
            const long long rows = ((N + chunks - 1) / chunks + 255) & ~255ll;
            for (int i = 1; i <= chunks; ++i) bounds[i] = std::min(N, rows * i);
        }
    } else {
        bounds[1] = N;
    }
    const size_t arow = (size_t)e->A * sizeof(E), orow = (size_t)e->Dout * sizeof(E);
    /* MBT_PIPE_TRACE=1: timestamps of every chunk's H2D / kernel / D2H completion on stderr (debugging the overlap) */
    static const bool trace = getenv("MBT_PIPE_TRACE") != nullptr;
    cudaEvent_t tr[3 * MBT_PIPE_CHUNKS + 1] = {};
    if (trace) {
        for (auto &ev : tr) cudaEventCreate(&ev);
        cudaEventRecord(tr[3 * MBT_PIPE_CHUNKS], e->copy_in);
    }
    for (int k = 0; k < chunks; ++k) {
        const long long r0 = bounds[k];
        if (r0 >= N) break;
        const long long n = std::min(bounds[k + 1], N) - r0;
        if (n <= 0) continue;
        if (act_pageable) par_memcpy((char *)const_cast<void *>(act_src) + r0 * arow, (const char *)act_pageable + r0 * arow, (size_t)n * arow);
        CU(cudaMemcpyAsync((char *)e->d_actions + r0 * arow, (const char *)act_src + r0 * arow, n * arow,
                           cudaMemcpyHostToDevice, e->copy_in));
        CU(cudaEventRecord(e->ev_in[k], e->copy_in));
        if (trace) cudaEventRecord(tr[3 * k], e->copy_in);
        CU(cudaStreamWaitEvent(e->stream, e->ev_in[k], 0));
        if (batch) {
            int rcb = launch_fill_batch<T, E>(e, p, e->d_actions, /*allow_pdl=*/false);
            if (rcb) return rcb;
        }
        int rc = launch_step_rows<T, E>(e, p, ck, e->d_actions, obs_dst ? e->d_obs : nullptr, rew_dst ? e->d_rew : nullptr, r0, n,
                                     /*allow_pdl=*/false); /* ordered by stream events, not by the previous kernel */
        if (rc) return rc;
        CU(cudaEventRecord(e->ev_k[k], e->stream));
        if (trace) cudaEventRecord(tr[3 * k + 1], e->stream);
        CU(cudaStreamWaitEvent(e->copy_out, e->ev_k[k], 0));
        if (obs_dst)
            CU(cudaMemcpyAsync((char *)obs_dst + r0 * orow, (const char *)e->d_obs + r0 * orow, n * orow,
                               cudaMemcpyDeviceToHost, e->copy_out));
        /* rewards: one copy per call, behind the last chunk (every D2H copy costs ~15 us of engine time on top of its
         * bytes; the (N,) reward vector is a fifth of the output) -- unless MBT_PIPE_REW_PER_CHUNK=1 */
        static const bool rew_per_chunk = getenv("MBT_PIPE_REW_PER_CHUNK") != nullptr;
        const bool last = bounds[k + 1] >= N;
        if (rew_dst && rew_per_chunk)
            CU(cudaMemcpyAsync((char *)rew_dst + r0 * sizeof(E), (const char *)e->d_rew + r0 * sizeof(E), n * sizeof(E),
                               cudaMemcpyDeviceToHost, e->copy_out));
        else if (rew_dst && last)
            CU(cudaMemcpyAsync(rew_dst, e->d_rew, (size_t)N * sizeof(E), cudaMemcpyDeviceToHost, e->copy_out));
        if (trace) cudaEventRecord(tr[3 * k + 2], e->copy_out);
    }
    if (trace) {
        cudaStreamSynchronize(e->copy_out);
        cudaStreamSynchronize(e->stream);
        fprintf(stderr, "[mbt pipe]");
        for (int k = 0; k < chunks; ++k) {
            float a = 0, b = 0, c2 = 0;
            cudaEventElapsedTime(&a, tr[3 * MBT_PIPE_CHUNKS], tr[3 * k]);
            cudaEventElapsedTime(&b, tr[3 * MBT_PIPE_CHUNKS], tr[3 * k + 1]);
            cudaEventElapsedTime(&c2, tr[3 * MBT_PIPE_CHUNKS], tr[3 * k + 2]);
            fprintf(stderr, "  chunk %d (%lld rows): h2d %.0f  kernel %.0f  d2h %.0f us |", k, bounds[k + 1] - bounds[k], 1e3 * a, 1e3 * b, 1e3 * c2);
        }
        fprintf(stderr, "\n");
        for (auto &ev : tr) cudaEventDestroy(ev);
    }
    CU(cudaStreamSynchronize(e->copy_out));
    CU(cudaStreamSynchronize(e->stream));
    advance_clock(e, t_next);
    if (done_out) *done_out = (uint8_t)ck.done;
    return fold_eager(e);
}

template <typename T, typename E>
static int do_reset_device(mbt_env *e, const mbt_reset_args *args, void *obs) {
    const mbt_config &c = e->cfg;
    const double t0 = args ? args->start_time : c.start_time;
    const int q0_mode = args ? args->q0_mode : c.q0_mode;
    const double q0_const = args ? args->q0_const : c.q0_const;
    const int64_t lo = args ? args->q0_lo : c.q0_lo, hi = args ? args->q0_hi : c.q0_hi;
    if (q0_mode == MBT_Q0_UNIFORM_INT && !(hi > lo)) return fail(MBT_E_INVALID_ARG, "initial inventory range needs hi > lo");
    if (q0_mode != MBT_Q0_CONST && q0_mode != MBT_Q0_UNIFORM_INT && q0_mode != MBT_Q0_PER_TRAJ)
        return fail(MBT_E_INVALID_ARG, "unknown q0_mode");
    if (!(t0 >= 0.0) || !(t0 < c.terminal_time))
        return fail(MBT_E_INVALID_ARG, "Start time is not within (0, env.terminal_time)."); /* TradingEnvironment.py:267 */
    if (q0_mode == MBT_Q0_PER_TRAJ) {
        /* one initial inventory per trajectory (a callable returned an array): converted once to the state type and
         * uploaded into the q0 column, where the rewards read it (RewardFunctions.py:72,111) and reset takes it from */
        if (!args || !args->q0_values) return fail(MBT_E_INVALID_ARG, "MBT_Q0_PER_TRAJ needs mbt_reset_args.q0_values");
        if (stream_is_capturing(e)) return fail(MBT_E_STATE, "per-trajectory initial inventories are uploaded from the host: not capturable");
        std::vector<T> q0((size_t)e->N);
        for (long long i = 0; i < e->N; ++i) q0[(size_t)i] = (T)args->q0_values[i];
        CU(cudaMemcpyAsync(e->col[5], q0.data(), (size_t)e->N * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream)); /* `q0` is a pageable temporary */
    }
    ResetArgs<T, E> g;
    g.p = mbt_make_params<T>(c, t0, q0_mode != MBT_Q0_CONST, q0_const);
    g.st = dev_state<T>(e);
    g.obs = (E *)obs;
    g.n = e->N;
    g.seed = e->seed;
    g.traj_offset = (unsigned long long)c.traj_offset;
    g.n_episode = (unsigned long long)e->n_episode;
    {
        int rcb = counter_base_for_launch(e, &g.counter_base);
        if (rcb) return rcb;
    }
    g.cash0 = (T)c.initial_cash;
    g.t0_obs = mbt_time_obs<T>(c, t0);
    g.mid0 = (T)c.mid_initial;
    g.lam0[0] = (T)c.arr_rate[0];
    g.lam0[1] = (T)c.arr_rate[1];
    g.imp0 = c.impact == MBT_IMP_TEMP_PERM ? (T)0 : (T)c.imp_initial;
    g.var0 = (T)c.heston_var0;
    g.q0_mode = q0_mode;
    g.q0_const = (T)q0_const;
    g.q0_lo = lo;
    g.q0_span = (unsigned long long)(hi - lo);
    mbt_reset_kernel<T, E><<<grid_for(g.n), MBT_BLOCK, 0, e->stream>>>(g);
    CU(cudaGetLastError());
    e->launches += 1;
    e->t = t0;
    e->t0 = t0;
    e->k = 0;
    e->n_episode += 1;
    e->started = true;
    e->q0_per_traj = (q0_mode != MBT_Q0_CONST);
    e->q0_uniform = q0_const;
    return fold_eager(e);
}

/* (arithmetic type, caller-buffer element type) dispatch: (double,double), (double,float) or (float,float) */
#define MBT_CALL_TE(e, fn, ...)                                                                          \
    ((e)->cfg.precision == MBT_F64 ? ((e)->io_esz == 4 ? fn<double, float>(__VA_ARGS__) : fn<double, double>(__VA_ARGS__)) \
                                   : fn<float, float>(__VA_ARGS__))

/* the device-resident counter base, read back (synchronises the stream; not capturable) */
static int read_counter_base(mbt_env *e, unsigned long long base[2]) {
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(base, e->d_counter_base, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MBT_OK;
}

/* ------------------------------------------------------------------ ABI */
extern "C" {

int mbt_abi_version(void) { return MBT_ABI_VERSION; }
const char *mbt_last_error(void) { return g_err.c_str(); }

int mbt_config_dims(const mbt_config *cfg, int32_t *action_dim, int32_t *obs_dim, int32_t *state_cols) {
    if (!cfg) return fail(MBT_E_INVALID_ARG, "config is NULL");
    int rc = mbt_dims(cfg, action_dim, obs_dim, state_cols);
    if (rc) return fail(rc, "unknown dynamics kind");
    return MBT_OK;
}

int mbt_config_obs_out_dim(const mbt_config *cfg, int32_t *obs_out_dim) {
    if (!cfg || !obs_out_dim) return fail(MBT_E_INVALID_ARG, "NULL argument");
    int32_t D = 0;
    int rc = mbt_dims(cfg, nullptr, &D, nullptr);
    if (rc) return fail(rc, "unknown dynamics kind");
    *obs_out_dim = mbt_obs_out_dim(cfg, D);
    return MBT_OK;
}

int mbt_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(MBT_E_INVALID_ARG, "out is NULL");
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        device = 0;
    }
    return mbt_host_alloc_near(bytes, device, out);
}

int mbt_host_alloc_near(size_t bytes, int device, void **out) {
    if (!out) return fail(MBT_E_INVALID_ARG, "out is NULL");
    ScopedNumaAffinity near_gpu(device);
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    memset(*out, 0, bytes ? bytes : 1); /* first touch on the GPU-local node */
    return MBT_OK;
}
int mbt_host_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return MBT_OK;
}

static void group_teardown(mbt_env *e);

int mbt_destroy(mbt_env *e) {
    if (!e) return MBT_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->stream) clear_l2_window(e, e->stream);
    if (e->has_l2_window && --g_l2_window_handles <= 0) { /* last windowed handle: give the set-aside back to everybody */
        g_l2_window_handles = 0;
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    cudaFree(e->state_block);
    cudaFree(e->d_clipped);
    cudaFree(e->d_counter_base);
    cudaFree(e->d_fill_partial);
    cudaFree(e->d_fill_ticket);
    cudaFree(e->d_actions);
    cudaFree(e->d_obs);
    cudaFree(e->d_rew);
    cudaFreeHost(e->h_actions);
    cudaFreeHost(e->h_obs);
    cudaFreeHost(e->h_rew);
    cudaFree(e->d_clocks);
    cudaFree(e->d_table);
    cudaFree(e->d_block_sums);
    cudaFree(e->d_summary);
    cudaFreeHost(e->h_summary);
    cudaFree(e->d_roll_ticket);
    group_teardown(e);
    for (auto ev : e->ev0) cudaEventDestroy(ev);
    for (auto ev : e->ev1) cudaEventDestroy(ev);
    for (int i = 0; i < MBT_PIPE_CHUNKS; ++i) {
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_k[i]) cudaEventDestroy(e->ev_k[i]);
    }
    if (e->copy_in) cudaStreamDestroy(e->copy_in);
    if (e->copy_out) cudaStreamDestroy(e->copy_out);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    cudaGetLastError();
    delete e;
    return MBT_OK;
}

int mbt_create(const mbt_config *cfg, int device, mbt_env **out) {
    if (!out) return fail(MBT_E_INVALID_ARG, "out is NULL");
    *out = nullptr;
    std::string err;
    int rc = mbt_validate_config(cfg, err);
    if (rc) return fail(rc, err);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MBT_E_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(ce) +
                                    "); libmbt_b200 has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(MBT_E_INVALID_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    mbt_env *e = new (std::nothrow) mbt_env();
    if (!e) return fail(MBT_E_NOMEM, "out of host memory");
    e->cfg = *cfg;
    e->device = device;
    e->N = cfg->num_trajectories;
    mbt_dims(cfg, &e->A, &e->D, &e->S);
    e->Dout = mbt_obs_out_dim(cfg, e->D);
    e->esz = cfg->precision == MBT_F64 ? 8 : 4;
    e->io_esz = (cfg->precision == MBT_F64 && cfg->io_precision == MBT_IO_SAME) ? 8 : 4;
    auto bail = [&](int code) {
        std::string keep = g_err;
        mbt_destroy(e);
        g_err = keep;
        return code;
    };
#define CUB(expr)                                                                                         \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            fail(_e == cudaErrorMemoryAllocation ? MBT_E_NOMEM : MBT_E_CUDA,                              \
                 std::string(#expr " failed: ") + cudaGetErrorString(_e));                                \
            return bail(_e == cudaErrorMemoryAllocation ? MBT_E_NOMEM : MBT_E_CUDA);                      \
        }                                                                                                 \
    } while (0)
    CUB(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUB(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    CUB(cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < MBT_PIPE_CHUNKS; ++i) {
        CUB(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
        CUB(cudaEventCreateWithFlags(&e->ev_k[i], cudaEventDisableTiming));
    }
    /* columns padded to 256 B so every column base is aligned for any vector width */
    const size_t col_bytes = (((size_t)e->N * e->esz) + 255) & ~(size_t)255;
    CUB(cudaMalloc(&e->state_block, col_bytes * MBT_STATE_COLUMNS));
    CUB(cudaMemsetAsync(e->state_block, 0, col_bytes * MBT_STATE_COLUMNS, e->stream));
    for (int i = 0; i < MBT_STATE_COLUMNS; ++i) e->col[i] = (char *)e->state_block + col_bytes * i;
    CUB(cudaMalloc(&e->d_clipped, sizeof(unsigned long long)));
    CUB(cudaMemsetAsync(e->d_clipped, 0, sizeof(unsigned long long), e->stream));
    CUB(cudaMalloc((void **)&e->d_counter_base, 2 * sizeof(unsigned long long)));
    CUB(cudaMemsetAsync(e->d_counter_base, 0, 2 * sizeof(unsigned long long), e->stream));
    /* batch-reduction cells (a few bytes; allocated for every handle: mbt_reconfigure may switch the fill function) */
    e->fill_blocks = std::max(1, e->sm_count * fill_blocks_per_sm()); /* every block ends in 3 same-address atomics */
    CUB(cudaMalloc(&e->d_fill_partial, 2 * sizeof(unsigned long long)));
    CUB(cudaMemsetAsync(e->d_fill_partial, 0, 2 * sizeof(unsigned long long), e->stream));
    CUB(cudaMalloc((void **)&e->d_fill_ticket, sizeof(unsigned int)));
    CUB(cudaMemsetAsync(e->d_fill_ticket, 0, sizeof(unsigned int), e->stream));
    /* episode summary: folded on the device by the rollout kernel, mirrored to pinned host memory on demand */
    CUB(cudaMalloc((void **)&e->d_summary, MBT_SUMMARY_DOUBLES * sizeof(double)));
    CUB(cudaMemsetAsync(e->d_summary, 0, MBT_SUMMARY_DOUBLES * sizeof(double), e->stream));
    CUB(cudaHostAlloc((void **)&e->h_summary, MBT_SUMMARY_DOUBLES * sizeof(double), cudaHostAllocDefault));
    CUB(cudaMalloc((void **)&e->d_roll_ticket, sizeof(unsigned int)));
    CUB(cudaMemsetAsync(e->d_roll_ticket, 0, sizeof(unsigned int), e->stream));
    CUB(cudaStreamSynchronize(e->stream));
#undef CUB
    if ((rc = jit_select_step(e)) != MBT_OK) return bail(rc);
    apply_l2_window(e, e->own_stream);
    *out = e;
    return MBT_OK;
}

int mbt_set_stream(mbt_env *e, void *cuda_stream) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    cudaStream_t next = cuda_stream == MBT_OWN_STREAM ? e->own_stream : (cudaStream_t)cuda_stream;
    if (next == e->stream) return MBT_OK; /* unchanged: nothing to order */
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream)); /* work queued on the old stream must finish before the new one is used */
    clear_l2_window(e, e->stream);
    e->stream = next;
    apply_l2_window(e, next);
    return MBT_OK;
}

int mbt_sync(mbt_env *e) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return MBT_OK;
}

int mbt_seed(mbt_env *e, uint64_t seed) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    e->seed = seed;
    e->n_step = 0;
    e->n_episode = 0;
    if (e->device_counters) { /* the device-resident base restarts too */
        CU(cudaSetDevice(e->device));
        CU(cudaMemsetAsync(e->d_counter_base, 0, 2 * sizeof(unsigned long long), e->stream));
    }
    return MBT_OK;
}

int mbt_get_seed(mbt_env *e, uint64_t *seed) {
    if (!e || !seed) return fail(MBT_E_INVALID_ARG, "NULL argument");
    *seed = e->seed;
    return MBT_OK;
}

int mbt_set_counters(mbt_env *e, int64_t steps_since_seed, int64_t episodes_since_seed) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    if (steps_since_seed < 0 || episodes_since_seed < 0) return fail(MBT_E_INVALID_ARG, "counters must be >= 0");
    if (stream_is_capturing(e)) return fail(MBT_E_STATE, "mbt_set_counters is not capturable");
    CU(cudaSetDevice(e->device));
    if (e->device_counters) { /* the device base is the truth in this mode */
        const unsigned long long base[2] = {(unsigned long long)steps_since_seed, (unsigned long long)episodes_since_seed};
        CU(cudaMemcpyAsync(e->d_counter_base, base, sizeof base, cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->n_step = 0;
        e->n_episode = 0;
    } else {
        e->n_step = steps_since_seed;
        e->n_episode = episodes_since_seed;
    }
    return MBT_OK;
}

/* which state columns a configuration keeps alive (the part of the layout mbt_reconfigure must not change) */
static unsigned state_layout_key(const mbt_config &c) {
    return (c.midprice == MBT_MID_HESTON ? 1u : 0u) | (c.arrival == MBT_ARR_HAWKES ? 2u : 0u) | (imp_has_state(c.impact) ? 4u : 0u) |
           ((c.midprice == MBT_MID_CONSTANT ? 1u : 0u) << 3);
}

int mbt_reconfigure(mbt_env *e, const mbt_config *cfg) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    std::string err;
    int rc = mbt_validate_config(cfg, err);
    if (rc) return fail(rc, err);
    const mbt_config &o = e->cfg;
    int32_t A = 0, D = 0, S = 0;
    mbt_dims(cfg, &A, &D, &S);
    if (cfg->num_trajectories != o.num_trajectories || cfg->traj_offset != o.traj_offset || cfg->precision != o.precision ||
        cfg->io_precision != o.io_precision || A != e->A || D != e->D || state_layout_key(*cfg) != state_layout_key(o))
        return fail(MBT_E_STATE, "mbt_reconfigure: the new configuration changes the shape of the device state (num_trajectories, "
                                 "precision, action / observation widths or the set of model state columns)");
    if (stream_is_capturing(e)) return fail(MBT_E_STATE, "mbt_reconfigure is not capturable");
    e->cfg = *cfg;
    e->Dout = mbt_obs_out_dim(cfg, e->D);
    CU(cudaSetDevice(e->device));
    apply_l2_window(e, e->stream);
    return jit_select_step(e);
}

/* lazily allocate the device + pinned staging used by MBT_MEM_HOST calls */
static int ensure_staging(mbt_env *e) {
    if (e->d_actions) return MBT_OK;
    const size_t ab = (size_t)e->N * e->A * e->esz, ob = (size_t)e->N * e->D * e->esz, rb = (size_t)e->N * e->esz;
    CU(cudaMalloc(&e->d_actions, ab));
    CU(cudaMalloc(&e->d_obs, ob));
    CU(cudaMalloc(&e->d_rew, rb));
    ScopedNumaAffinity near_gpu(e->device);
    CU(cudaHostAlloc(&e->h_actions, ab, cudaHostAllocDefault));
    CU(cudaHostAlloc(&e->h_obs, ob, cudaHostAllocDefault));
    CU(cudaHostAlloc(&e->h_rew, rb, cudaHostAllocDefault));
    return MBT_OK;
}

/* device -> caller host buffer: direct DMA when the buffer is pinned, else through pinned staging */
static int d2h(mbt_env *e, void *host_dst, const void *dev_src, void *pinned_stage, size_t bytes, bool *needs_unstage) {
    *needs_unstage = false;
    if (host_ptr_is_pinned(host_dst)) {
        CU(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, e->stream));
    } else {
        CU(cudaMemcpyAsync(pinned_stage, dev_src, bytes, cudaMemcpyDeviceToHost, e->stream));
        *needs_unstage = true;
    }
    return MBT_OK;
}

int mbt_reset(mbt_env *e, const mbt_reset_args *args, void *obs_out, int mem) {
    NvtxRange nvtx("mbt_reset");
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    if (mem != MBT_MEM_HOST && mem != MBT_MEM_DEVICE) return fail(MBT_E_INVALID_ARG, "mem must be MBT_MEM_HOST or MBT_MEM_DEVICE");
    CU(cudaSetDevice(e->device));
    void *dev_obs = obs_out;
    if (obs_out && mem == MBT_MEM_HOST) {
        int rc = ensure_staging(e);
        if (rc) return rc;
        dev_obs = e->d_obs;
    }
    int rc = MBT_CALL_TE(e, do_reset_device, e, args, dev_obs);
    if (rc) return rc;
    if (obs_out && mem == MBT_MEM_HOST) {
        const size_t ob = (size_t)e->N * e->Dout * e->io_esz;
        bool unstage = false;
        rc = d2h(e, obs_out, e->d_obs, e->h_obs, ob, &unstage);
        if (rc) return rc;
        CU(cudaStreamSynchronize(e->stream));
        if (unstage) par_memcpy(obs_out, e->h_obs, ob);
    }
    return MBT_OK;
}

int mbt_step(mbt_env *e, const void *actions, void *obs_out, void *rew_out, uint8_t *done_out, int mem) {
    NvtxRange nvtx("mbt_step");
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    if (!actions) return fail(MBT_E_INVALID_ARG, "actions is NULL");
    if (mem != MBT_MEM_HOST && mem != MBT_MEM_DEVICE) return fail(MBT_E_INVALID_ARG, "mem must be MBT_MEM_HOST or MBT_MEM_DEVICE");
    if (!e->started) return fail(MBT_E_STATE, "mbt_step called before mbt_reset");
    CU(cudaSetDevice(e->device));
    if (mem == MBT_MEM_DEVICE) return MBT_CALL_TE(e, do_step_device, e, actions, obs_out, rew_out, done_out);

    /* host buffers: H2D actions -> kernel -> D2H observations + rewards, pipelined, all inside this call */
    int rc = ensure_staging(e);
    if (rc) return rc;
    const size_t ab = (size_t)e->N * e->A * e->io_esz, ob = (size_t)e->N * e->Dout * e->io_esz, rb = (size_t)e->N * e->io_esz;
    const void *src = actions;
    const bool stage_act = !host_ptr_is_pinned(actions); /* pageable caller memory: through the handle's pinned buffer */
    const bool pipelined = !host_path_zero_copy();
    if (stage_act) {
        if (!pipelined) par_memcpy(e->h_actions, actions, ab); /* (the pipeline stages chunk by chunk) */
        src = e->h_actions;
    }
    const bool un_obs = obs_out && !host_ptr_is_pinned(obs_out), un_rew = rew_out && !host_ptr_is_pinned(rew_out);
    void *obs_dst = obs_out ? (un_obs ? e->h_obs : obs_out) : nullptr;
    void *rew_dst = rew_out ? (un_rew ? e->h_rew : rew_out) : nullptr;
    if (host_path_zero_copy()) {
        /* the step kernel reads the action rows and writes observation rows / rewards DIRECTLY in pinned host memory
         * (UVA mapping): one launch, PCIe reads and posted writes in flight together, no DMA descriptors */
        void *da = nullptr, *dobs = nullptr, *drew = nullptr;
        CU(cudaHostGetDevicePointer(&da, const_cast<void *>(src), 0));
        if (obs_dst) CU(cudaHostGetDevicePointer(&dobs, obs_dst, 0));
        if (rew_dst) CU(cudaHostGetDevicePointer(&drew, rew_dst, 0));
        rc = MBT_CALL_TE(e, do_step_device, e, da, dobs, drew, done_out);
        if (rc) return rc;
        CU(cudaStreamSynchronize(e->stream));
    } else {
        rc = MBT_CALL_TE(e, do_step_host_pipelined, e, src, obs_dst, rew_dst, done_out, stage_act ? actions : nullptr);
        if (rc) return rc;
    }
    if (un_obs) par_memcpy(obs_out, e->h_obs, ob);
    if (un_rew) par_memcpy(rew_out, e->h_rew, rb);
    return MBT_OK;
}

int mbt_get_state(mbt_env *e, void *state_out, int mem) {
    if (!e || !state_out) return fail(MBT_E_INVALID_ARG, "NULL argument");
    CU(cudaSetDevice(e->device));
    void *dev = state_out;
    if (mem == MBT_MEM_HOST) {
        int rc = ensure_staging(e);
        if (rc) return rc;
        dev = e->d_obs;
    }
    if (e->cfg.precision == MBT_F64) {
        auto p = mbt_make_params<double>(e->cfg, e->t0, e->q0_per_traj, e->q0_uniform);
        mbt_gather_state_kernel<double><<<grid_for(e->N), MBT_BLOCK, 0, e->stream>>>(p, dev_state<double>(e), e->t, (double *)dev, e->N);
    } else {
        auto p = mbt_make_params<float>(e->cfg, e->t0, e->q0_per_traj, e->q0_uniform);
        mbt_gather_state_kernel<float><<<grid_for(e->N), MBT_BLOCK, 0, e->stream>>>(p, dev_state<float>(e), (float)e->t, (float *)dev, e->N);
    }
    CU(cudaGetLastError());
    e->launches += 1;
    if (mem == MBT_MEM_HOST) {
        CU(cudaMemcpyAsync(state_out, dev, (size_t)e->N * e->D * e->esz, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return MBT_OK;
}

int mbt_set_state(mbt_env *e, const void *state_in, int mem) {
    if (!e || !state_in) return fail(MBT_E_INVALID_ARG, "NULL argument");
    CU(cudaSetDevice(e->device));
    const void *dev = state_in;
    const size_t ob = (size_t)e->N * e->D * e->esz;
    if (mem == MBT_MEM_HOST) {
        int rc = ensure_staging(e);
        if (rc) return rc;
        CU(cudaMemcpyAsync(e->d_obs, state_in, ob, cudaMemcpyHostToDevice, e->stream));
        dev = e->d_obs;
        /* the clock is the TIME column of row 0 (uniform)   TradingEnvironment.py:219 */
        e->t = e->cfg.precision == MBT_F64 ? ((const double *)state_in)[2] : (double)((const float *)state_in)[2];
    } else {
        char buf[8];
        CU(cudaMemcpyAsync(buf, (const char *)state_in + 2 * e->esz, e->esz, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->t = e->cfg.precision == MBT_F64 ? *(double *)buf : (double)*(float *)buf;
    }
    if (e->cfg.precision == MBT_F64) {
        auto p = mbt_make_params<double>(e->cfg, e->t0, e->q0_per_traj, e->q0_uniform);
        mbt_scatter_state_kernel<double><<<grid_for(e->N), MBT_BLOCK, 0, e->stream>>>(p, dev_state<double>(e), (const double *)dev, e->N);
    } else {
        auto p = mbt_make_params<float>(e->cfg, e->t0, e->q0_per_traj, e->q0_uniform);
        mbt_scatter_state_kernel<float><<<grid_for(e->N), MBT_BLOCK, 0, e->stream>>>(p, dev_state<float>(e), (const float *)dev, e->N);
    }
    CU(cudaGetLastError());
    e->launches += 1;
    e->started = true;
    if (mem == MBT_MEM_HOST) CU(cudaStreamSynchronize(e->stream));
    return MBT_OK;
}

int mbt_get_clock(mbt_env *e, double *time, int64_t *steps_this_episode, int64_t *steps_since_seed, int64_t *episodes_since_seed) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    unsigned long long base[2] = {0, 0};
    if (e->device_counters && (steps_since_seed || episodes_since_seed)) {
        int rc = read_counter_base(e, base);
        if (rc) return rc;
    }
    if (time) *time = e->t;
    if (steps_this_episode) *steps_this_episode = e->k;
    if (steps_since_seed) *steps_since_seed = e->n_step + (int64_t)base[0];
    if (episodes_since_seed) *episodes_since_seed = e->n_episode + (int64_t)base[1];
    return MBT_OK;
}

/* ---- checkpoint / resume: header + the raw structure-of-arrays block (counter-based RNG: no generator state) */
struct mbt_ckpt_header {
    uint64_t magic;
    int64_t num_trajectories;
    int32_t precision, obs_dim;
    uint64_t seed;
    double t, t0, q0_uniform;
    int64_t k, n_step, n_episode;
    int32_t q0_per_traj, started;
    uint64_t state_bytes;
};
static const uint64_t MBT_CKPT_MAGIC = 0x4D42543230304231ull; /* "MBT200B1" */

static size_t state_block_bytes(const mbt_env *e) {
    return ((((size_t)e->N * e->esz) + 255) & ~(size_t)255) * MBT_STATE_COLUMNS;
}

int mbt_checkpoint_size(mbt_env *e, size_t *bytes) {
    if (!e || !bytes) return fail(MBT_E_INVALID_ARG, "NULL argument");
    *bytes = sizeof(mbt_ckpt_header) + state_block_bytes(e);
    return MBT_OK;
}

int mbt_checkpoint_save(mbt_env *e, void *host_buf, size_t capacity) {
    if (!e || !host_buf) return fail(MBT_E_INVALID_ARG, "NULL argument");
    const size_t sb = state_block_bytes(e);
    if (capacity < sizeof(mbt_ckpt_header) + sb) return fail(MBT_E_INVALID_ARG, "checkpoint buffer too small");
    CU(cudaSetDevice(e->device));
    mbt_ckpt_header h;
    memset(&h, 0, sizeof h);
    h.magic = MBT_CKPT_MAGIC;
    h.num_trajectories = e->N;
    h.precision = e->cfg.precision;
    h.obs_dim = e->D;
    h.seed = e->seed;
    h.t = e->t; h.t0 = e->t0; h.q0_uniform = e->q0_uniform;
    unsigned long long base[2] = {0, 0};
    if (e->device_counters) {
        int rc = read_counter_base(e, base);
        if (rc) return rc;
    }
    h.k = e->k; h.n_step = e->n_step + (int64_t)base[0]; h.n_episode = e->n_episode + (int64_t)base[1];
    h.q0_per_traj = e->q0_per_traj; h.started = e->started ? 1 : 0;
    h.state_bytes = sb;
    memcpy(host_buf, &h, sizeof h);
    CU(cudaMemcpyAsync((char *)host_buf + sizeof h, e->state_block, sb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MBT_OK;
}

int mbt_checkpoint_load(mbt_env *e, const void *host_buf, size_t bytes) {
    if (!e || !host_buf) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (bytes < sizeof(mbt_ckpt_header)) return fail(MBT_E_INVALID_ARG, "checkpoint truncated");
    mbt_ckpt_header h;
    memcpy(&h, host_buf, sizeof h);
    if (h.magic != MBT_CKPT_MAGIC) return fail(MBT_E_INVALID_ARG, "not an mbt_b200 checkpoint");
    if (h.num_trajectories != e->N || h.precision != e->cfg.precision || h.obs_dim != e->D ||
        h.state_bytes != state_block_bytes(e) || bytes < sizeof h + h.state_bytes)
        return fail(MBT_E_INVALID_ARG, "checkpoint does not match this handle (num_trajectories / precision / model layout)");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->state_block, (const char *)host_buf + sizeof h, h.state_bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(e->d_counter_base, 0, 2 * sizeof(unsigned long long), e->stream)); /* counters come back on the host side */
    CU(cudaStreamSynchronize(e->stream));
    e->seed = h.seed;
    e->t = h.t; e->t0 = h.t0; e->q0_uniform = h.q0_uniform;
    e->k = h.k; e->n_step = h.n_step; e->n_episode = h.n_episode;
    e->q0_per_traj = h.q0_per_traj; e->started = h.started != 0;
    return MBT_OK;
}

int mbt_get_clip_count(mbt_env *e, int64_t *count) {
    if (!e || !count) return fail(MBT_E_INVALID_ARG, "NULL argument");
    CU(cudaSetDevice(e->device));
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, e->d_clipped, sizeof v, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *count = (int64_t)v;
    return MBT_OK;
}

int mbt_reward_eval(mbt_env *e, int64_t n, const void *current_state, const void *action, const void *next_state, int is_terminal,
                    void *rew_out, int mem) {
    if (!e || !current_state || !action || !next_state || !rew_out) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (n <= 0) return fail(MBT_E_INVALID_ARG, "n must be > 0");
    CU(cudaSetDevice(e->device));
    const size_t sb = (size_t)n * e->D * e->esz, ab = (size_t)n * e->A * e->esz, rb = (size_t)n * e->esz;
    const void *dc = current_state, *da = action, *dn = next_state;
    void *dr = rew_out;
    void *tmp = nullptr;
    if (mem == MBT_MEM_HOST) {
        CU(cudaMalloc(&tmp, 2 * sb + ab + rb));
        char *b = (char *)tmp;
        CU(cudaMemcpyAsync(b, current_state, sb, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(b + sb, next_state, sb, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(b + 2 * sb, action, ab, cudaMemcpyHostToDevice, e->stream));
        dc = b; dn = b + sb; da = b + 2 * sb; dr = b + 2 * sb + ab;
    }
    if (e->cfg.precision == MBT_F64) {
        auto p = mbt_make_params<double>(e->cfg, e->t0, 0, e->q0_uniform);
        mbt_reward_kernel<double><<<grid_for(n), MBT_BLOCK, 0, e->stream>>>(p, is_terminal, (const double *)dc, (const double *)da, (const double *)dn, (double *)dr, n);
    } else {
        auto p = mbt_make_params<float>(e->cfg, e->t0, 0, e->q0_uniform);
        mbt_reward_kernel<float><<<grid_for(n), MBT_BLOCK, 0, e->stream>>>(p, is_terminal, (const float *)dc, (const float *)da, (const float *)dn, (float *)dr, n);
    }
    cudaError_t le = cudaGetLastError();
    e->launches += 1;
    if (mem == MBT_MEM_HOST) {
        if (le == cudaSuccess) le = cudaMemcpyAsync(rew_out, dr, rb, cudaMemcpyDeviceToHost, e->stream);
        cudaError_t se = cudaStreamSynchronize(e->stream);
        cudaFree(tmp);
        if (le == cudaSuccess) le = se;
    }
    if (le != cudaSuccess) return fail(MBT_E_CUDA, std::string("mbt_reward_eval: ") + cudaGetErrorString(le));
    return MBT_OK;
}

} /* extern "C" */

/* enqueue the fused rollout of the rest of the episode on the handle's stream; the summary stays on the device */
template <typename T>
static int enqueue_rollout(mbt_env *e, const mbt_policy *pol, void *returns, void *term_q, const mbt_record *rec, int *steps_out) {
    const mbt_config &c = e->cfg;
    /* the clock exactly as repeated `state[:, TIME] += step_size` produces it   TradingEnvironment.py:216 */
    std::vector<double> times;
    times.push_back(e->t);
    int steps = 0;
    {
        double t = e->t;
        const int cap = 1 << 24;
        while (steps < cap) {
            t = t + c.step_size;
            times.push_back(t);
            steps += 1;
            if (t >= c.terminal_time - c.step_size / 2) break;
        }
    }
    /* every step's uniform clock, formed exactly like the step kernel's (mbt_make_clock): the fused rollout and a loop of
     * mbt_step calls see bit-identical StepClock values */
    std::vector<RolloutClock<T>> clocks((size_t)steps);
    for (int k = 0; k < steps; ++k) {
        clocks[k].ck = mbt_make_clock<T>(c, times[k], times[k + 1], e->t0);
        clocks[k].t_cur = (T)times[k];
        clocks[k].t_cur_obs = mbt_time_obs<T>(c, times[k]);
    }
    const size_t clock_bytes = clocks.size() * sizeof(RolloutClock<T>);
    if (clock_bytes > e->clocks_cap) {
        cudaFree(e->d_clocks);
        e->d_clocks = nullptr;
        e->clocks_cap = 0;
        CU(cudaMalloc(&e->d_clocks, clock_bytes));
        e->clocks_cap = clock_bytes;
    }
    /* pageable source: the copy is staged before cudaMemcpyAsync returns, so `clocks` may go out of scope */
    CU(cudaMemcpyAsync(e->d_clocks, clocks.data(), clock_bytes, cudaMemcpyHostToDevice, e->stream));

    RolloutArgs<T> g;
    memset(&g, 0, sizeof g);
    g.p = mbt_make_params<T>(c, e->t0, e->q0_per_traj, e->q0_uniform);
    g.st = dev_state<T>(e);
    g.n = e->N;
    g.keys = mbt_philox_expand(e->seed);
    g.traj_offset = (unsigned long long)c.traj_offset;
    g.n_step0 = (unsigned long long)e->n_step;
    {
        int rcb = counter_base_for_launch(e, &g.counter_base);
        if (rcb) return rcb;
    }
    g.steps = steps;
    g.clocks = (const RolloutClock<T> *)e->d_clocks;
    g.pol_kind = pol->kind;
    g.table_rows = pol->table_rows;
    g.table_cols = pol->table_cols;
    g.inv_offset = pol->inv_offset;
    for (int j = 0; j < MBT_MAX_ACTION_DIM; ++j) g.fixed[j] = (T)pol->fixed[j];
    g.as_gamma = (T)pol->as_gamma;
    g.as_sigma_sq = (T)pol->as_sigma_sq;
    g.as_fill_comp = (T)pol->as_fill_comp;
    g.as_terminal_time = (T)pol->as_terminal_time;
    std::vector<T> host_table;
    if (pol->kind == MBT_POL_CJ_MM_TABLE || pol->kind == MBT_POL_SCHEDULE) {
        if (!pol->table) return fail(MBT_E_INVALID_ARG, "policy table is NULL");
        if (pol->table_rows < steps) return fail(MBT_E_INVALID_ARG, "policy table has fewer rows than the steps left in the episode");
        size_t count = pol->kind == MBT_POL_CJ_MM_TABLE ? (size_t)pol->table_rows * pol->table_cols * 2 : (size_t)pol->table_rows * e->A;
        if (pol->kind == MBT_POL_CJ_MM_TABLE && (pol->table_cols != 2 * pol->inv_offset + 1 || pol->inv_offset < 0))
            return fail(MBT_E_INVALID_ARG, "CJ_MM_TABLE needs table_cols == 2*inv_offset+1");
        host_table.resize(count);
        for (size_t i = 0; i < count; ++i) host_table[i] = (T)pol->table[i];
        if (count * sizeof(T) > e->table_cap) {
            cudaFree(e->d_table);
            e->d_table = nullptr;
            CU(cudaMalloc(&e->d_table, count * sizeof(T)));
            e->table_cap = count * sizeof(T);
        }
        CU(cudaMemcpyAsync(e->d_table, host_table.data(), count * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        g.table = (const T *)e->d_table;
    } else if (pol->kind != MBT_POL_FIXED && pol->kind != MBT_POL_AVELLANEDA_STOIKOV) {
        return fail(MBT_E_UNSUPPORTED, "unknown policy kind");
    }
    if (pol->kind == MBT_POL_AVELLANEDA_STOIKOV && e->A != 2) return fail(MBT_E_UNSUPPORTED, "Avellaneda-Stoikov policy needs a 2-d action");
    if (needs_fill_batch(c) && (pol->kind == MBT_POL_AVELLANEDA_STOIKOV || pol->kind == MBT_POL_CJ_MM_TABLE))
        return fail(MBT_E_UNSUPPORTED,
                    "triangular / power fill functions couple the trajectories of a step (np.max(depths, 0) in the reference): the "
                    "fused rollout supports them only with policies that are uniform over the batch (fixed, schedule); use step()");
    g.returns = (T *)returns;
    g.term_q = (T *)term_q;
    if (rec) {
        if (rec->steps_capacity < steps) return fail(MBT_E_INVALID_ARG, "record buffers hold fewer steps than the episode has left");
        g.rec_obs = (T *)rec->obs;
        g.rec_act = (T *)rec->actions;
        g.rec_rew = (T *)rec->rewards;
    }
    const int blocks = (int)grid_for(g.n);
    if (blocks > e->block_sums_cap) {
        cudaFree(e->d_block_sums);
        e->d_block_sums = nullptr;
        CU(cudaMalloc(&e->d_block_sums, (size_t)blocks * MBT_SUMMARY_DOUBLES * sizeof(double)));
        e->block_sums_cap = blocks;
    }
    g.block_sums = e->d_block_sums;
    g.summary = e->d_summary;
    g.ticket = e->d_roll_ticket;
    g.clipped = e->d_clipped;
    int rc = timing_begin(e);
    if (rc) return rc;
    /* the fast path (no recording) has the policy kind compiled in; the recording kernels (store-bound) switch at run time */
    const mbt_jit::Module *jm = nullptr;
    rc = jit_rollout_module(e, rec ? 4 : pol->kind, &jm);
    if (rc) return rc;
    if (jm) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)blocks);
        lc.blockDim = dim3(MBT_BLOCK);
        lc.stream = e->stream;
        void *args[1] = {&g};
        cudaLaunchKernelExC(&lc, (const void *)jm->k0, args);
    } else
    switch (variant_of(c)) {
#define X(id, ...)                                                                                                        \
    case id:                                                                                                              \
        if (rec) mbt_rollout_kernel<T, __VA_ARGS__, true, -1><<<blocks, MBT_BLOCK, 0, e->stream>>>(g);                    \
        else if (pol->kind == MBT_POL_FIXED)                                                                              \
            mbt_rollout_kernel<T, __VA_ARGS__, false, MBT_POL_FIXED><<<blocks, MBT_BLOCK, 0, e->stream>>>(g);             \
        else if (pol->kind == MBT_POL_AVELLANEDA_STOIKOV)                                                                 \
            mbt_rollout_kernel<T, __VA_ARGS__, false, MBT_POL_AVELLANEDA_STOIKOV><<<blocks, MBT_BLOCK, 0, e->stream>>>(g); \
        else if (pol->kind == MBT_POL_CJ_MM_TABLE)                                                                        \
            mbt_rollout_kernel<T, __VA_ARGS__, false, MBT_POL_CJ_MM_TABLE><<<blocks, MBT_BLOCK, 0, e->stream>>>(g);       \
        else                                                                                                              \
            mbt_rollout_kernel<T, __VA_ARGS__, false, MBT_POL_SCHEDULE><<<blocks, MBT_BLOCK, 0, e->stream>>>(g);          \
        break;
        MBT_FOR_EACH_VARIANT(X)
#undef X
    }
    CU(cudaGetLastError());
    rc = timing_end(e);
    if (rc) return rc;
    e->launches += 1;
    e->t = times.back();
    e->k += steps;
    e->n_step += steps;
    *steps_out = steps;
    return fold_eager(e);
}

/* the device-resident summary of the last rollout (local, or all-reduced over the group) -> host struct.  One small D2H
 * into pinned memory on the handle's stream + one synchronize: the only host round trip of an episode. */
static int fetch_summary(mbt_env *e, int steps, long long count, mbt_summary *summary) {
    CU(cudaMemcpyAsync(e->h_summary, e->d_summary, MBT_SUMMARY_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (summary) {
        summary->count = count;
        summary->steps = steps;
        summary->sum_return = e->h_summary[0];
        summary->sum_return_sq = e->h_summary[1];
        summary->sum_q = e->h_summary[2];
        summary->sum_q_sq = e->h_summary[3];
        summary->sum_action = e->h_summary[4];
        summary->sum_reward_sq = e->h_summary[5];
        summary->clipped = (int64_t)e->h_summary[6];
    }
    return MBT_OK;
}

template <typename T>
static int do_rollout(mbt_env *e, const mbt_policy *pol, mbt_summary *summary, void *returns, void *term_q,
                      const mbt_record *rec = nullptr) {
    int steps = 0;
    int rc = enqueue_rollout<T>(e, pol, returns, term_q, rec, &steps);
    if (rc) return rc;
    return fetch_summary(e, steps, e->N, summary);
}

extern "C" {

int mbt_rollout(mbt_env *e, const mbt_policy *policy, mbt_summary *summary_out, void *returns_out, void *terminal_q_out, int mem) {
    NvtxRange nvtx("mbt_rollout");
    if (!e || !policy) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (!e->started) return fail(MBT_E_STATE, "mbt_rollout called before mbt_reset");
    if (e->t >= e->cfg.terminal_time - e->cfg.step_size / 2) return fail(MBT_E_STATE, "episode already finished; call mbt_reset");
    CU(cudaSetDevice(e->device));
    void *d_ret = returns_out, *d_q = terminal_q_out;
    const size_t rb = (size_t)e->N * e->esz;
    if (mem == MBT_MEM_HOST && (returns_out || terminal_q_out)) {
        int rc = ensure_staging(e);
        if (rc) return rc;
        d_ret = returns_out ? e->d_rew : nullptr;
        d_q = terminal_q_out ? e->d_obs : nullptr; /* d_obs holds >= N elements */
    }
    int rc = e->cfg.precision == MBT_F64 ? do_rollout<double>(e, policy, summary_out, d_ret, d_q)
                                         : do_rollout<float>(e, policy, summary_out, d_ret, d_q);
    if (rc) return rc;
    if (mem == MBT_MEM_HOST) {
        if (returns_out) CU(cudaMemcpyAsync(returns_out, d_ret, rb, cudaMemcpyDeviceToHost, e->stream));
        if (terminal_q_out) CU(cudaMemcpyAsync(terminal_q_out, d_q, rb, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return MBT_OK;
}

int mbt_rollout_record(mbt_env *e, const mbt_policy *policy, mbt_summary *summary_out, const mbt_record *record, int mem) {
    NvtxRange nvtx("mbt_rollout_record");
    if (!e || !policy || !record) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (!e->started) return fail(MBT_E_STATE, "mbt_rollout_record called before mbt_reset");
    if (e->t >= e->cfg.terminal_time - e->cfg.step_size / 2) return fail(MBT_E_STATE, "episode already finished; call mbt_reset");
    if (record->steps_capacity <= 0) return fail(MBT_E_INVALID_ARG, "steps_capacity must be > 0");
    CU(cudaSetDevice(e->device));
    const size_t cap = (size_t)record->steps_capacity;
    const size_t ob = (cap + 1) * (size_t)e->N * e->Dout * e->esz, ab = cap * (size_t)e->N * e->A * e->esz, rb = cap * (size_t)e->N * e->esz;
    mbt_record dev = *record;
    void *tmp = nullptr;
    if (mem == MBT_MEM_HOST) { /* stage through one temporary device block */
        const size_t need = (record->obs ? ob : 0) + (record->actions ? ab : 0) + (record->rewards ? rb : 0);
        cudaError_t me = cudaMalloc(&tmp, need ? need : 1);
        if (me != cudaSuccess) {
            cudaGetLastError();
            return fail(MBT_E_NOMEM, "not enough device memory to record the trajectory; record fewer trajectories");
        }
        char *b = (char *)tmp;
        dev.obs = record->obs ? b : nullptr;      b += record->obs ? ob : 0;
        dev.actions = record->actions ? b : nullptr; b += record->actions ? ab : 0;
        dev.rewards = record->rewards ? b : nullptr;
    }
    int rc = e->cfg.precision == MBT_F64 ? do_rollout<double>(e, policy, summary_out, nullptr, nullptr, &dev)
                                         : do_rollout<float>(e, policy, summary_out, nullptr, nullptr, &dev);
    if (rc == MBT_OK && mem == MBT_MEM_HOST) {
        cudaError_t ce = cudaSuccess;
        if (record->obs) ce = cudaMemcpyAsync(record->obs, dev.obs, ob, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess && record->actions) ce = cudaMemcpyAsync(record->actions, dev.actions, ab, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess && record->rewards) ce = cudaMemcpyAsync(record->rewards, dev.rewards, rb, cudaMemcpyDeviceToHost, e->stream);
        cudaError_t se = cudaStreamSynchronize(e->stream);
        if (ce == cudaSuccess) ce = se;
        if (ce != cudaSuccess) rc = fail(MBT_E_CUDA, std::string("mbt_rollout_record: ") + cudaGetErrorString(ce));
    }
    if (tmp) cudaFree(tmp);
    return rc;
}

int mbt_prepare_capture(mbt_env *e) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    CU(cudaSetDevice(e->device));
    if (stream_is_capturing(e)) return fail(MBT_E_STATE, "mbt_prepare_capture must be called before the capture begins");
    e->device_counters = true;
    return fold_eager(e); /* base += counters consumed so far; host counters = 0 */
}

int mbt_fold_counters(mbt_env *e) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    CU(cudaSetDevice(e->device));
    e->device_counters = true;
    mbt_fold_counters_kernel<<<1, 1, 0, e->stream>>>(e->d_counter_base, (unsigned long long)e->n_step, (unsigned long long)e->n_episode);
    CU(cudaGetLastError());
    e->launches += 1;
    e->n_step = 0;
    e->n_episode = 0;
    return MBT_OK;
}

int mbt_get_launch_count(mbt_env *e, int64_t *launches) {
    if (!e || !launches) return fail(MBT_E_INVALID_ARG, "NULL argument");
    *launches = e->launches;
    return MBT_OK;
}

int mbt_enable_timing(mbt_env *e, int enable) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    e->timing = enable == 2 ? 2 : (enable != 0 ? 1 : 0);
    e->timed = 0;
    e->interval_closed = false;
    return MBT_OK;
}

int mbt_get_kernel_times(mbt_env *e, float *ms_out, int64_t capacity, int64_t *count) {
    if (!e || !count) return fail(MBT_E_INVALID_ARG, "NULL argument");
    CU(cudaSetDevice(e->device));
    if (e->timing == 2 && e->timed > 0 && !e->interval_closed) {
        /* interval mode: launch i lasted from its own event to the next launch's event; close the last interval */
        CU(cudaEventRecord(e->ev1[e->timed - 1], e->stream));
        e->interval_closed = true;
    }
    CU(cudaStreamSynchronize(e->stream));
    int64_t n = std::min<int64_t>(e->timed, capacity);
    for (int64_t i = 0; i < n && ms_out; ++i) {
        if (e->timing == 2 && i + 1 < e->timed)
            CU(cudaEventElapsedTime(&ms_out[i], e->ev0[i], e->ev0[i + 1]));
        else
            CU(cudaEventElapsedTime(&ms_out[i], e->ev0[i], e->ev1[i]));
    }
    *count = e->timed;
    return MBT_OK;
}

} /* extern "C" */

/* ------------------------------------------------------------------ group of handles over NCCL */
/* NCCL is loaded at run time (dlopen): single-GPU users need no NCCL, and inside a torch process the already loaded
 * libnccl.so.2 (torch's own) is the one that answers.  Only the handful of entry points below are used. */
namespace {
typedef struct { char internal[MBT_GROUP_ID_BYTES]; } nccl_unique_id;
typedef void *nccl_comm_t;
enum { NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 }; /* ncclDataType_t (nccl.h) */
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                                           /* ncclRedOp_t                */
struct NcclApi {
    void *so = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};
NcclApi *nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!a.so) {
            a.error = std::string("NCCL not found (dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?") + ")";
            return a;
        }
#define SYM(field, name)                                                      \
    *(void **)(&a.field) = dlsym(a.so, name);                                 \
    if (!a.field && a.error.empty()) a.error = std::string("libnccl lacks ") + name;
        SYM(GetUniqueId, "ncclGetUniqueId")
        SYM(CommInitRank, "ncclCommInitRank")
        SYM(CommDestroy, "ncclCommDestroy")
        SYM(AllReduce, "ncclAllReduce")
        SYM(AllGather, "ncclAllGather")
        SYM(Broadcast, "ncclBroadcast")
        SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd")
        SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return a;
    }();
    return &api;
}
} /* namespace */

#define NC(expr)                                                                                                   \
    do {                                                                                                           \
        int _r = (expr);                                                                                           \
        if (_r != 0) {                                                                                             \
            char _b[512];                                                                                          \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(_r), __FILE__, __LINE__); \
            return fail(MBT_E_NCCL, _b);                                                                           \
        }                                                                                                          \
    } while (0)

static void group_teardown(mbt_env *e) {
    if (e->comm && nccl_api()->CommDestroy) {
        if (e->comm_stream) cudaStreamSynchronize(e->comm_stream);
        nccl_api()->CommDestroy((nccl_comm_t)e->comm);
    }
    e->comm = nullptr;
    if (e->ev_rollout) cudaEventDestroy(e->ev_rollout);
    if (e->ev_gather) cudaEventDestroy(e->ev_gather);
    if (e->comm_stream) cudaStreamDestroy(e->comm_stream);
    e->ev_rollout = e->ev_gather = nullptr;
    e->comm_stream = nullptr;
    e->g_rank = 0;
    e->g_world = 1;
    e->g_counts.clear();
    e->gather_pending = false;
}

static int group_allreduce_max_u64(mbt_env *e, unsigned long long *dev, int count) {
    NC(nccl_api()->AllReduce(dev, dev, (size_t)count, NCCL_UINT64, NCCL_MAX, (nccl_comm_t)e->comm, e->stream));
    return MBT_OK;
}

extern "C" {

int mbt_group_unique_id(void *id_out) {
    if (!id_out) return fail(MBT_E_INVALID_ARG, "id_out is NULL");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(MBT_E_UNSUPPORTED, api->error);
    nccl_unique_id id;
    NC(api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return MBT_OK;
}

int mbt_group_create(mbt_env *e, const void *id, int rank, int world) {
    if (!e || !id) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(MBT_E_INVALID_ARG, "need 0 <= rank < world");
    if (e->comm) return fail(MBT_E_STATE, "the handle already belongs to a group");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(MBT_E_UNSUPPORTED, api->error);
    CU(cudaSetDevice(e->device));
    nccl_unique_id uid;
    memcpy(&uid, id, sizeof uid);
    nccl_comm_t comm = nullptr;
    NC(api->CommInitRank(&comm, world, uid, rank));
    e->comm = comm;
    e->g_rank = rank;
    e->g_world = world;
    CU(cudaStreamCreateWithFlags(&e->comm_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&e->ev_rollout, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->ev_gather, cudaEventDisableTiming));
    /* shard sizes of all ranks (the gather of returns places shard r at the sum of the sizes before it) */
    long long *d_counts = nullptr;
    CU(cudaMalloc((void **)&d_counts, (size_t)world * sizeof(long long)));
    const long long mine = e->N;
    CU(cudaMemcpyAsync(d_counts + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, e->stream));
    int r = api->AllGather(d_counts + rank, d_counts, 1, NCCL_INT64, comm, e->stream);
    e->g_counts.assign((size_t)world, 0);
    cudaError_t ce = cudaMemcpyAsync(e->g_counts.data(), d_counts, (size_t)world * sizeof(long long), cudaMemcpyDeviceToHost, e->stream);
    cudaError_t se = cudaStreamSynchronize(e->stream);
    cudaFree(d_counts);
    if (r != 0 || ce != cudaSuccess || se != cudaSuccess) {
        group_teardown(e);
        return fail(MBT_E_NCCL, std::string("mbt_group_create: exchanging shard sizes failed: ") +
                                    (r != 0 ? api->GetErrorString(r) : cudaGetErrorString(ce != cudaSuccess ? ce : se)));
    }
    return MBT_OK;
}

int mbt_group_destroy(mbt_env *e) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    group_teardown(e);
    return MBT_OK;
}

int mbt_group_info(mbt_env *e, int32_t *rank, int32_t *world, int64_t *total) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    if (rank) *rank = e->g_rank;
    if (world) *world = e->g_world;
    if (total) {
        long long t = 0;
        if (e->comm) for (long long c : e->g_counts) t += c;
        else t = e->N;
        *total = t;
    }
    return MBT_OK;
}

/* all-reduce (sum) of the device-resident summary on the handle's stream */
static int group_allreduce_summary(mbt_env *e) {
    NC(nccl_api()->AllReduce(e->d_summary, e->d_summary, MBT_SUMMARY_DOUBLES, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)e->comm, e->stream));
    return MBT_OK;
}

int mbt_group_rollout(mbt_env *e, const mbt_policy *policy, mbt_summary *summary_out, void *returns_local, void *returns_all) {
    NvtxRange nvtx("mbt_group_rollout");
    if (!e || !policy) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (!e->comm) return fail(MBT_E_STATE, "the handle belongs to no group (mbt_group_create)");
    if (!e->started) return fail(MBT_E_STATE, "mbt_group_rollout called before mbt_reset");
    if (e->t >= e->cfg.terminal_time - e->cfg.step_size / 2) return fail(MBT_E_STATE, "episode already finished; call mbt_reset");
    if (returns_all && !returns_local) return fail(MBT_E_INVALID_ARG, "returns_all needs returns_local (the shard's own returns)");
    CU(cudaSetDevice(e->device));
    if (e->gather_pending && returns_local == e->gather_src) {
        /* the previous episode's gather is still reading this buffer: order the rollout's writes behind it.  A caller
         * that alternates two returns buffers gets the gather overlapped with the next episode's rollout. */
        CU(cudaStreamWaitEvent(e->stream, e->ev_gather, 0));
        e->gather_pending = false;
    }
    int steps = 0;
    int rc = e->cfg.precision == MBT_F64 ? enqueue_rollout<double>(e, policy, returns_local, nullptr, nullptr, &steps)
                                         : enqueue_rollout<float>(e, policy, returns_local, nullptr, nullptr, &steps);
    if (rc) return rc;
    if (returns_all) {
        /* the gather runs on the group's own stream behind the rollout: it overlaps the summary's host round trip and
         * whatever the caller enqueues next on the handle's stream */
        CU(cudaEventRecord(e->ev_rollout, e->stream));
        CU(cudaStreamWaitEvent(e->comm_stream, e->ev_rollout, 0));
    }
    rc = group_allreduce_summary(e);
    if (rc) return rc;
    if (returns_all) {
        const int dt = e->cfg.precision == MBT_F64 ? NCCL_FLOAT64 : NCCL_FLOAT32;
        bool equal = true;
        for (long long c : e->g_counts) equal = equal && c == e->g_counts[0];
        NcclApi *api = nccl_api();
        if (equal) {
            NC(api->AllGather(returns_local, returns_all, (size_t)e->N, dt, (nccl_comm_t)e->comm, e->comm_stream));
        } else { /* ragged shards (N % world != 0): one broadcast per rank, grouped */
            NC(api->GroupStart());
            long long off = 0;
            for (int r = 0; r < e->g_world; ++r) {
                char *dst = (char *)returns_all + (size_t)off * e->esz;
                int br = api->Broadcast(r == e->g_rank ? returns_local : dst, dst, (size_t)e->g_counts[(size_t)r], dt, r, (nccl_comm_t)e->comm, e->comm_stream);
                if (br != 0) {
                    api->GroupEnd();
                    return fail(MBT_E_NCCL, std::string("ncclBroadcast failed: ") + api->GetErrorString(br));
                }
                off += e->g_counts[(size_t)r];
            }
            NC(api->GroupEnd());
        }
        CU(cudaEventRecord(e->ev_gather, e->comm_stream));
        e->gather_pending = true;
        e->gather_src = returns_local;
    }
    long long total = 0;
    for (long long c : e->g_counts) total += c;
    return fetch_summary(e, steps, total, summary_out);
}

int mbt_group_summary(mbt_env *e, const mbt_summary *local, mbt_summary *global_out) {
    if (!e || !local || !global_out) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (!e->comm) return fail(MBT_E_STATE, "the handle belongs to no group (mbt_group_create)");
    CU(cudaSetDevice(e->device));
    e->h_summary[0] = local->sum_return; e->h_summary[1] = local->sum_return_sq; e->h_summary[2] = local->sum_q;
    e->h_summary[3] = local->sum_q_sq; e->h_summary[4] = local->sum_action; e->h_summary[5] = local->sum_reward_sq;
    e->h_summary[6] = (double)local->clipped;
    CU(cudaMemcpyAsync(e->d_summary, e->h_summary, MBT_SUMMARY_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    int rc = group_allreduce_summary(e);
    if (rc) return rc;
    long long total = 0;
    for (long long c : e->g_counts) total += c;
    return fetch_summary(e, (int)local->steps, total, global_out);
}

int mbt_group_wait(mbt_env *e, int host_sync) {
    if (!e) return fail(MBT_E_INVALID_ARG, "env is NULL");
    if (!e->comm) return fail(MBT_E_STATE, "the handle belongs to no group (mbt_group_create)");
    CU(cudaSetDevice(e->device));
    if (e->gather_pending) {
        CU(cudaStreamWaitEvent(e->stream, e->ev_gather, 0));
        e->gather_pending = false;
    }
    if (host_sync) {
        CU(cudaStreamSynchronize(e->comm_stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return MBT_OK;
}

} /* extern "C" */

extern "C" {

int mbt_get_kernel_info(mbt_env *e, mbt_kernel_info *out) {
    if (!e || !out) return fail(MBT_E_INVALID_ARG, "NULL argument");
    memset(out, 0, sizeof *out);
    out->aot_variant = variant_of(e->cfg);
    out->jit_mode = mbt_jit::mode();
    if (e->jit_step) {
        out->step_is_jit = 1;
        out->step_registers = e->jit_step->regs;
        out->step_local_bytes = e->jit_step->local_bytes;
        out->jit_compile_ms = e->jit_step->compile_ms;
        out->jit_from_disk_cache = e->jit_step->from_disk ? 1 : 0;
        out->jit_hash = e->jit_step->hash;
    }
    for (int i = 0; i < 5; ++i) out->rollout_is_jit[i] = e->jit_roll_state[i] == 1;
    snprintf(out->message, sizeof out->message, "%s", e->jit_error.c_str());
    return MBT_OK;
}

int mbt_jit_precompile(const mbt_config *cfg, int32_t kind, int32_t policy_kind, int32_t record) {
    std::string err;
    int rc = mbt_validate_config(cfg, err);
    if (rc) return fail(rc, err);
    if (!wants_jit(*cfg)) return MBT_OK; /* served by a fully specialised ahead-of-time variant */
    const int io64 = cfg->precision == MBT_F64 && (kind != 0 || cfg->io_precision == MBT_IO_SAME);
    const mbt_jit::Key key = mbt_jit::key_of(*cfg, io64, kind == 0 ? mbt_jit::STEP : mbt_jit::ROLLOUT, record ? -1 : policy_kind, record ? 1 : 0);
    std::vector<char> cubin;
    unsigned long long h = 0;
    bool from_disk = false;
    double ms = 0;
    rc = mbt_jit::cubin_of(key, cubin, &h, &from_disk, &ms, err);
    if (rc) return fail(rc, err);
    return MBT_OK;
}

} /* extern "C" */

extern "C" {

int mbt_inventory_histogram(mbt_env *e, int64_t lo, int64_t hi, int64_t *counts_out, int group_sum) {
    if (!e || !counts_out) return fail(MBT_E_INVALID_ARG, "NULL argument");
    if (!(hi >= lo) || hi - lo + 1 > MBT_HIST_MAX_BINS) return fail(MBT_E_INVALID_ARG, "need lo <= hi and at most 4096 bins");
    if (!e->started) return fail(MBT_E_STATE, "mbt_inventory_histogram called before mbt_reset");
    if (group_sum && !e->comm) return fail(MBT_E_STATE, "the handle belongs to no group (mbt_group_create)");
    if (stream_is_capturing(e)) return fail(MBT_E_STATE, "mbt_inventory_histogram is not capturable");
    CU(cudaSetDevice(e->device));
    const int bins = (int)(hi - lo + 1);
    const size_t bytes = (size_t)(bins + 2) * sizeof(unsigned long long);
    unsigned long long *d_counts = nullptr;
    CU(cudaMallocAsync((void **)&d_counts, bytes, e->stream));
    cudaError_t ce = cudaMemsetAsync(d_counts, 0, bytes, e->stream);
    const unsigned blocks = std::min<unsigned>(grid_for(e->N), (unsigned)(e->sm_count * 8));
    const size_t smem = (size_t)(bins + 2) * sizeof(unsigned int);
    if (ce == cudaSuccess) {
        if (e->cfg.precision == MBT_F64)
            mbt_inventory_hist_kernel<double><<<blocks, MBT_BLOCK, smem, e->stream>>>((const double *)e->col[1], e->N, (long long)lo, bins, d_counts);
        else
            mbt_inventory_hist_kernel<float><<<blocks, MBT_BLOCK, smem, e->stream>>>((const float *)e->col[1], e->N, (long long)lo, bins, d_counts);
        ce = cudaGetLastError();
        e->launches += 1;
    }
    int rc = MBT_OK;
    if (ce == cudaSuccess && group_sum) { /* counts of all ranks: NCCL all-reduce (sum) on the handle's stream */
        int r = nccl_api()->AllReduce(d_counts, d_counts, (size_t)(bins + 2), NCCL_UINT64, NCCL_SUM, (nccl_comm_t)e->comm, e->stream);
        if (r != 0) rc = fail(MBT_E_NCCL, std::string("ncclAllReduce failed: ") + nccl_api()->GetErrorString(r));
    }
    if (ce == cudaSuccess && rc == MBT_OK) ce = cudaMemcpyAsync(counts_out, d_counts, bytes, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t se = cudaStreamSynchronize(e->stream);
    cudaFreeAsync(d_counts, e->stream);
    if (rc) return rc;
    if (ce != cudaSuccess || se != cudaSuccess)
        return fail(MBT_E_CUDA, std::string("mbt_inventory_histogram: ") + cudaGetErrorString(ce != cudaSuccess ? ce : se));
    return MBT_OK;
}

} /* extern "C" */
