/*
 * mbt_jit.h -- run-time specialisation of the step / rollout kernels for ONE configuration (NVRTC).
 *
 * The ahead-of-time table (mbt_variants.h) holds a dozen hand-picked instantiations; every other configuration used to
 * run the generic kernel, whose model switches cost registers (64 against 32-48), occupancy and instruction-cache space.
 * A handle's configuration is frozen between mbt_create / mbt_reconfigure calls, so the library compiles
 *     VariantFull<dynamics, midprice, arrival, impact, reward, fill, normalise_action, normalise_obs, normalise_rewards,
 *                 obs_select>
 * for exactly that configuration the first time it is needed: the SAME kernel bodies (mbt_step_body / mbt_rollout_body,
 * mbt_kernels.cuh) wrapped in extern "C" kernels, compiled with the same floating-point flags as the nvcc build
 * (--fmad=false, IEEE division / sqrt, no FTZ), so results are bit-identical to the ahead-of-time kernels.
 *
 *   source     the six headers are embedded in the library at build time (_jit_embed.inc, written by _build.py)
 *   compiler   libnvrtc.so.12, loaded with dlopen (no NVRTC = the ahead-of-time kernels keep running; MBT_JIT=0 turns the
 *              specialiser off, MBT_JIT=require makes a failure an error)
 *   cache      cubins are kept per process (hash map) and on disk (<library dir>/_jit_cache/<hash>.cubin, or
 *              $MBT_JIT_CACHE_DIR; written atomically), keyed by a hash of the generated source, the embedded headers,
 *              the compile options and the NVRTC version -- a typical compile takes 0.2-0.5 s, a cache hit ~1 ms
 *   loading    cudaLibraryLoadData / cudaLibraryGetKernel (context-independent kernel handles), launched with
 *              cudaLaunchKernelExC like the ahead-of-time kernels (programmatic dependent launch included)
 */
#ifndef MBT_JIT_H
#define MBT_JIT_H

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mbt_b200.h"

namespace mbt_jit {

struct EmbeddedHeader {
    const char *name;
    const char *text;
};
#include "_jit_embed.inc" /* static const EmbeddedHeader kHeaders[]; static const int kNumHeaders; */

enum Kind { STEP = 0, ROLLOUT = 1 };

struct Key {
    int dyn, mid, arr, imp, rew, fill, na, no, nr, sel;
    int f64;     /* arithmetic type: 1 double, 0 float */
    int io_f64;  /* caller-buffer element type of the step kernel */
    int kind;    /* STEP | ROLLOUT */
    int pol;     /* ROLLOUT: MBT_POL_* compiled in, or -1 with rec */
    int rec;     /* ROLLOUT: recording stores compiled in */
};

static inline Key key_of(const mbt_config &c, int io_f64, int kind, int pol, int rec) {
    Key k;
    const bool fills = c.dynamics == MBT_DYN_LIMIT || c.dynamics == MBT_DYN_LIMIT_AND_MARKET;
    k.dyn = c.dynamics; k.mid = c.midprice; k.arr = c.arrival; k.imp = c.impact; k.rew = c.reward;
    k.fill = fills ? c.fill : MBT_FILL_NONE; /* what mbt_make_params puts into StepParams::fill */
    k.na = c.normalise_action ? 1 : 0; k.no = c.normalise_obs ? 1 : 0; k.nr = c.normalise_rewards ? 1 : 0;
    k.sel = (int)c.obs_select;
    k.f64 = c.precision == MBT_F64; k.io_f64 = io_f64; k.kind = kind; k.pol = pol; k.rec = rec;
    return k;
}

struct Module {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k0 = nullptr, k1 = nullptr; /* STEP: (whole-row vector access, scalar access); ROLLOUT: k0 */
    int regs = 0, local_bytes = 0;
    double compile_ms = 0;
    bool from_disk = false;
    unsigned long long hash = 0;
};

/* ------------------------------------------------------------------ NVRTC through dlopen */
struct Nvrtc {
    void *so = nullptr;
    typedef void *program;
    int (*Version)(int *, int *) = nullptr;
    int (*CreateProgram)(program *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(program, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(program, size_t *) = nullptr;
    int (*GetCUBIN)(program, char *) = nullptr;
    int (*GetProgramLogSize)(program, size_t *) = nullptr;
    int (*GetProgramLog)(program, char *) = nullptr;
    int (*DestroyProgram)(program *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int major = 0, minor = 0;
    std::string error;
};

static inline Nvrtc *nvrtc() {
    static Nvrtc api = [] {
        Nvrtc a;
        /* the toolkit's own copy first: a process may already hold an older libnvrtc.so.12 (e.g. a pip wheel), and the
         * 256-bit vector accesses of the double-precision rows need the PTX ISA of CUDA 12.9 */
        const char *names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"};
        if (const char *p = getenv("MBT_NVRTC_PATH")) a.so = dlopen(p, RTLD_NOW | RTLD_LOCAL);
        for (const char *n : names)
            if (!a.so) a.so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (!a.so) {
            const char *de = dlerror();
            a.error = std::string("NVRTC not found (dlopen libnvrtc.so.12: ") + (de ? de : "?") + ")";
            return a;
        }
#define MBT_SYM(field, name)                          \
    *(void **)(&a.field) = dlsym(a.so, name);         \
    if (!a.field && a.error.empty()) a.error = std::string("libnvrtc lacks ") + name;
        MBT_SYM(Version, "nvrtcVersion")
        MBT_SYM(CreateProgram, "nvrtcCreateProgram")
        MBT_SYM(CompileProgram, "nvrtcCompileProgram")
        MBT_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        MBT_SYM(GetCUBIN, "nvrtcGetCUBIN")
        MBT_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        MBT_SYM(GetProgramLog, "nvrtcGetProgramLog")
        MBT_SYM(DestroyProgram, "nvrtcDestroyProgram")
        MBT_SYM(GetErrorString, "nvrtcGetErrorString")
#undef MBT_SYM
        if (a.error.empty()) a.Version(&a.major, &a.minor);
        return a;
    }();
    return &api;
}

/* ------------------------------------------------------------------ source, hash, cache */
static inline unsigned long long fnv1a(const void *data, size_t n, unsigned long long h = 1469598103934665603ull) {
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

static inline std::string source_of(const Key &k) {
    char buf[2048];
    const char *T = k.f64 ? "double" : "float", *E = k.io_f64 ? "double" : "float";
    int n = snprintf(buf, sizeof buf,
                     "#include \"mbt_kernels.cuh\"\n"
                     "typedef VariantFull<%d, %d, %d, %d, %d, %d, %d, %d, %d, %d> VJ;\n",
                     k.dyn, k.mid, k.arr, k.imp, k.rew, k.fill, k.na, k.no, k.nr, k.sel);
    if (k.kind == STEP)
        snprintf(buf + n, sizeof buf - n,
                 "extern \"C\" __global__ void __launch_bounds__(MBT_BLOCK) mbt_jit_step_vec(const __grid_constant__ StepArgs<%s, %s> g) "
                 "{ mbt_step_body<%s, %s, VJ, true>(g); }\n"
                 "extern \"C\" __global__ void __launch_bounds__(MBT_BLOCK) mbt_jit_step(const __grid_constant__ StepArgs<%s, %s> g) "
                 "{ mbt_step_body<%s, %s, VJ, false>(g); }\n",
                 T, E, T, E, T, E, T, E);
    else
        snprintf(buf + n, sizeof buf - n,
                 "extern \"C\" __global__ void __launch_bounds__(MBT_BLOCK) mbt_jit_rollout(const __grid_constant__ RolloutArgs<%s> g) "
                 "{ mbt_rollout_body<%s, VJ, %s, %d>(g); }\n",
                 T, T, k.rec ? "true" : "false", k.pol);
    return buf;
}

static inline std::string library_dir() {
    Dl_info info;
    if (dladdr((const void *)&fnv1a, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        return s == std::string::npos ? "." : p.substr(0, s);
    }
    return ".";
}

static inline std::string cache_dir() {
    if (const char *d = getenv("MBT_JIT_CACHE_DIR")) return d;
    return library_dir() + "/_jit_cache";
}

static inline bool read_file(const std::string &path, std::vector<char> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    bool ok = n > 0 && fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

static inline void write_file_atomic(const std::string &path, const std::vector<char> &data) {
    std::string dir = path.substr(0, path.find_last_of('/'));
    mkdir(dir.c_str(), 0755);
    char tmp[64];
    snprintf(tmp, sizeof tmp, ".tmp.%d.%p", (int)getpid(), (const void *)&data);
    std::string t = dir + "/" + tmp;
    FILE *f = fopen(t.c_str(), "wb");
    if (!f) return; /* a read-only tree: the cubin simply stays in memory */
    bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok || rename(t.c_str(), path.c_str()) != 0) unlink(t.c_str());
}

/* Compile (or fetch from the disk cache) the cubin of `key`.  No CUDA device is needed for this part. */
static inline int cubin_of(const Key &key, std::vector<char> &cubin, unsigned long long *hash_out, bool *from_disk, double *ms,
                           std::string &err) {
    const std::string src = source_of(key);
    Nvrtc *rt = nvrtc();
    static const char *base_opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "--prec-div=true",
                                      "--prec-sqrt=true", "--ftz=false", "-lineinfo", "-default-device"};
    unsigned long long h = fnv1a(src.data(), src.size());
    for (int i = 0; i < kNumHeaders; ++i) h = fnv1a(kHeaders[i].text, strlen(kHeaders[i].text), h);
    for (const char *o : base_opts) h = fnv1a(o, strlen(o), h);
    const int ver[2] = {rt->major, rt->minor};
    h = fnv1a(ver, sizeof ver, h);
    *hash_out = h;
    char name[64];
    snprintf(name, sizeof name, "/%016llx.cubin", h);
    const std::string path = cache_dir() + name;
    const auto t0 = std::chrono::steady_clock::now();
    if (!getenv("MBT_JIT_NO_DISK_CACHE") && read_file(path, cubin)) {
        *from_disk = true;
        *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return MBT_OK;
    }
    *from_disk = false;
    if (!rt->error.empty()) {
        err = rt->error;
        return MBT_E_UNSUPPORTED;
    }
    std::vector<const char *> names, texts;
    for (int i = 0; i < kNumHeaders; ++i) {
        names.push_back(kHeaders[i].name);
        texts.push_back(kHeaders[i].text);
    }
    /* second attempt without the 256-bit vector accesses, for an NVRTC whose PTX ISA predates them */
    for (int attempt = 0; attempt < 2; ++attempt) {
        Nvrtc::program prog = nullptr;
        int r = rt->CreateProgram(&prog, src.c_str(), "mbt_jit.cu", kNumHeaders, texts.data(), names.data());
        if (r != 0) {
            err = std::string("nvrtcCreateProgram: ") + rt->GetErrorString(r);
            return MBT_E_CUDA;
        }
        std::vector<const char *> opts(base_opts, base_opts + sizeof base_opts / sizeof *base_opts);
        if (attempt == 1) opts.push_back("-DMBT_NO_256BIT");
        r = rt->CompileProgram(prog, (int)opts.size(), opts.data());
        if (r != 0) {
            size_t ln = 0;
            rt->GetProgramLogSize(prog, &ln);
            std::string log(ln, '\0');
            if (ln) rt->GetProgramLog(prog, &log[0]);
            rt->DestroyProgram(&prog);
            err = std::string("nvrtcCompileProgram: ") + rt->GetErrorString(r) + "\n" + log.substr(0, 1500);
            if (attempt == 0 && log.find("Vector type too large") != std::string::npos) continue;
            return MBT_E_CUDA;
        }
        size_t n = 0;
        rt->GetCUBINSize(prog, &n);
        cubin.resize(n);
        r = rt->GetCUBIN(prog, cubin.data());
        rt->DestroyProgram(&prog);
        if (r != 0 || n == 0) {
            err = "nvrtcGetCUBIN failed";
            return MBT_E_CUDA;
        }
        err.clear();
        if (!getenv("MBT_JIT_NO_DISK_CACHE")) write_file_atomic(path, cubin);
        *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return MBT_OK;
    }
    return MBT_E_CUDA;
}

/* The loaded module of `key` (compiled / read / taken from the per-process table).  Needs a current CUDA context. */
static inline int module_of(const Key &key, const Module **out, std::string &err) {
    static std::mutex mu;
    static std::map<std::string, Module> table; /* keyed by the generated source text */
    const std::string src = source_of(key);
    std::lock_guard<std::mutex> lock(mu);
    auto it = table.find(src);
    if (it != table.end()) {
        *out = &it->second;
        return MBT_OK;
    }
    std::vector<char> cubin;
    Module m;
    int rc = cubin_of(key, cubin, &m.hash, &m.from_disk, &m.compile_ms, err);
    if (rc) return rc;
    cudaError_t ce = cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce == cudaSuccess) ce = cudaLibraryGetKernel(&m.k0, m.lib, key.kind == STEP ? "mbt_jit_step_vec" : "mbt_jit_rollout");
    if (ce == cudaSuccess && key.kind == STEP) ce = cudaLibraryGetKernel(&m.k1, m.lib, "mbt_jit_step");
    if (ce != cudaSuccess) {
        cudaGetLastError();
        err = std::string("loading the specialised kernel failed: ") + cudaGetErrorString(ce);
        if (m.lib) cudaLibraryUnload(m.lib);
        return MBT_E_CUDA;
    }
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void *)m.k0) == cudaSuccess) {
        m.regs = fa.numRegs;
        m.local_bytes = (int)fa.localSizeBytes;
    } else {
        cudaGetLastError();
    }
    auto ins = table.emplace(src, m);
    *out = &ins.first->second;
    return MBT_OK;
}

/* MBT_JIT: "0" off, "require" failures are errors, anything else / unset = on with the ahead-of-time kernels as fallback */
static inline int mode() {
    static const int m = [] {
        const char *v = getenv("MBT_JIT");
        if (!v) return 1;
        if (strcmp(v, "0") == 0 || strcmp(v, "off") == 0) return 0;
        if (strcmp(v, "require") == 0) return 2;
        return 1;
    }();
    return m;
}

} /* namespace mbt_jit */

#endif /* MBT_JIT_H */
