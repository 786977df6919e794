// HReduce: row reductions with a caller-supplied combining expression.
//
// The reference renders `hreduce.mako` with the C expression `op`, an `identity` expression
// and free-form `extra_code`, and compiles it at template construction (reduce.py:72-89).
// The same contract needs a run-time compiler here as well, so this is the one place where
// the library does not ship its kernel pre-built: the kernel text below is specialised with
// the caller's strings and compiled for sm_100a with NVRTC, which is loaded with dlopen on
// first use (the rest of the library has no dependency on it).  The resulting cubin is loaded
// through the driver entry points that the static CUDA runtime hands out.
//
// Kernel: one warp per row, 8 rows per block.  Lanes walk the column range with a stride of
// 32 (coalesced), combine their elements with op, then the 32 partial values are combined by
// a butterfly of word-wise shuffles, so any trivially copyable element type works.  Only
// commutative, associative ops are supported, as in the reference (reduce.py:45).
#include "common.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

namespace {

using namespace ksp;

// ---------------------------------------------------------------- NVRTC, loaded lazily
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *,
                         const char *const *);
    int (*CompileProgram)(nvrtcProgram, int, const char *const *);
    int (*GetCUBINSize)(nvrtcProgram, size_t *);
    int (*GetCUBIN)(nvrtcProgram, char *);
    int (*GetProgramLogSize)(nvrtcProgram, size_t *);
    int (*GetProgramLog)(nvrtcProgram, char *);
    int (*DestroyProgram)(nvrtcProgram *);
    bool ok = false;
};

const Nvrtc &nvrtc()
{
    static Nvrtc api = [] {
        Nvrtc a;
        const char *names[] = {getenv("KSP_B200_NVRTC"), "libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.13"};
        void *h = nullptr;
        for (const char *n : names)
            if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!h) return a;
#define KSP_SYM(field, sym) *(void **) (&a.field) = dlsym(h, sym)
        KSP_SYM(CreateProgram, "nvrtcCreateProgram");
        KSP_SYM(CompileProgram, "nvrtcCompileProgram");
        KSP_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        KSP_SYM(GetCUBIN, "nvrtcGetCUBIN");
        KSP_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        KSP_SYM(GetProgramLog, "nvrtcGetProgramLog");
        KSP_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef KSP_SYM
        a.ok = a.CreateProgram && a.CompileProgram && a.GetCUBINSize && a.GetCUBIN &&
               a.GetProgramLogSize && a.GetProgramLog && a.DestroyProgram;
        return a;
    }();
    return api;
}

// ---------------------------------------------------------------- driver entry points
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule *, const void *);
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void **, void **);
    bool ok = false;
};

const Driver &driver()
{
    static Driver api = [] {
        Driver d;
        auto get = [](const char *name) -> void * {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult status;
            if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &status) != cudaSuccess ||
                status != cudaDriverEntryPointSuccess)
                return nullptr;
            return p;
        };
        *(void **) (&d.ModuleLoadData) = get("cuModuleLoadData");
        *(void **) (&d.ModuleGetFunction) = get("cuModuleGetFunction");
        *(void **) (&d.ModuleUnload) = get("cuModuleUnload");
        *(void **) (&d.LaunchKernel) = get("cuLaunchKernel");
        d.ok = d.ModuleLoadData && d.ModuleGetFunction && d.ModuleUnload && d.LaunchKernel;
        return d;
    }();
    return api;
}

thread_local std::string g_build_log;

// ---------------------------------------------------------------- the kernel text
const char *HREDUCE_SOURCE = R"KSP(
typedef KSP_TYPE elem_t;
KSP_EXTRA_CODE
__device__ __forceinline__ elem_t ksp_op(elem_t a, elem_t b) { return (KSP_OP); }
__device__ __forceinline__ elem_t ksp_identity() { return (KSP_IDENTITY); }

// exchange a value of any size between lanes, 32 bits at a time
__device__ __forceinline__ elem_t ksp_shfl_xor(elem_t v, int d)
{
    constexpr int WORDS = (sizeof(elem_t) + 3) / 4;
    union { elem_t v; unsigned w[WORDS]; } u;
#pragma unroll
    for (int i = 0; i < WORDS; i++) u.w[i] = 0u;
    u.v = v;
#pragma unroll
    for (int i = 0; i < WORDS; i++) u.w[i] = __shfl_xor_sync(0xffffffffu, u.w[i], d);
    return u.v;
}

extern "C" __global__ void __launch_bounds__(256)
ksp_hreduce_kernel(const elem_t *__restrict__ src, elem_t *__restrict__ dest, long long rows,
                   long long src_stride, long long first_col, long long n_cols)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long) blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;                                  // whole warp
    const elem_t *p = src + row * src_stride + first_col;
    elem_t acc = ksp_identity();
    long long c = lane;
    // four independent loads in flight per lane
    for (; c + 96 < n_cols; c += 128) {
        const elem_t x0 = p[c], x1 = p[c + 32], x2 = p[c + 64], x3 = p[c + 96];
        acc = ksp_op(acc, ksp_op(ksp_op(x0, x1), ksp_op(x2, x3)));
    }
    for (; c < n_cols; c += 32) acc = ksp_op(acc, p[c]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc = ksp_op(acc, ksp_shfl_xor(acc, d));
    if (lane == 0) dest[row] = acc;
}
)KSP";

struct HReduce {
    CUmodule module = nullptr;
    CUfunction function = nullptr;
    size_t elem_size = 0;
};

}  // namespace

extern "C" const char *ksp_jit_log(void) { return g_build_log.c_str(); }

extern "C" int ksp_hreduce_create(const char *ctype, const char *op, const char *identity,
                                  const char *extra_code, size_t elem_size, void **handle)
{
    g_build_log.clear();
    if (!ctype || !op || !identity || !handle || elem_size == 0 || elem_size > 64) return KSP_EINVAL;
    *handle = nullptr;
    const Nvrtc &rtc = nvrtc();
    if (!rtc.ok) {
        g_build_log = "NVRTC (libnvrtc.so.12) could not be loaded; set KSP_B200_NVRTC to its path";
        return KSP_ENOJIT;
    }
    KSP_CUDA(cudaFree(nullptr));                              // make sure the primary context exists
    const Driver &drv = driver();
    if (!drv.ok) {
        g_build_log = "CUDA driver entry points for module loading are not available";
        return KSP_ENOJIT;
    }
    std::string src;
    // NVRTC has no <math.h> / <float.h>: the constants user expressions commonly need
    src += "#ifndef INFINITY\n#define INFINITY __int_as_float(0x7f800000)\n#endif\n"
           "#ifndef NAN\n#define NAN __int_as_float(0x7fc00000)\n#endif\n"
           "#ifndef FLT_MAX\n#define FLT_MAX 3.402823466e+38f\n#endif\n"
           "#ifndef DBL_MAX\n#define DBL_MAX 1.7976931348623157e+308\n#endif\n";
    src += "#define KSP_TYPE " + std::string(ctype) + "\n";
    src += "#define KSP_OP " + std::string(op) + "\n";
    src += "#define KSP_IDENTITY " + std::string(identity) + "\n";
    src += "#define KSP_EXTRA_CODE\n";
    if (extra_code && *extra_code) src += std::string(extra_code) + "\n";
    src += "static_assert(sizeof(KSP_TYPE) == " + std::to_string(elem_size) +
           ", \"dtype and ctype sizes differ\");\n";
    src += HREDUCE_SOURCE;

    nvrtcProgram prog = nullptr;
    if (rtc.CreateProgram(&prog, src.c_str(), "ksp_hreduce.cu", 0, nullptr, nullptr) != 0) {
        g_build_log = "nvrtcCreateProgram failed";
        return KSP_EJIT;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    const int rc = rtc.CompileProgram(prog, 3, opts);
    size_t log_size = 0;
    if (rtc.GetProgramLogSize(prog, &log_size) == 0 && log_size > 1) {
        g_build_log.resize(log_size);
        rtc.GetProgramLog(prog, &g_build_log[0]);
    }
    if (rc != 0) {
        rtc.DestroyProgram(&prog);
        return KSP_EJIT;
    }
    size_t cubin_size = 0;
    std::vector<char> cubin;
    if (rtc.GetCUBINSize(prog, &cubin_size) != 0 || cubin_size == 0) {
        rtc.DestroyProgram(&prog);
        g_build_log += "\nno cubin produced";
        return KSP_EJIT;
    }
    cubin.resize(cubin_size);
    rtc.GetCUBIN(prog, cubin.data());
    rtc.DestroyProgram(&prog);

    HReduce *h = new HReduce;
    h->elem_size = elem_size;
    CUresult cr = drv.ModuleLoadData(&h->module, cubin.data());
    if (cr == CUDA_SUCCESS) cr = drv.ModuleGetFunction(&h->function, h->module, "ksp_hreduce_kernel");
    if (cr != CUDA_SUCCESS) {
        if (h->module) drv.ModuleUnload(h->module);
        delete h;
        g_build_log += "\nloading the compiled module failed (CUresult " + std::to_string((int) cr) + ")";
        return KSP_EJIT;
    }
    *handle = h;
    return 0;
}

extern "C" int ksp_hreduce_destroy(void *handle)
{
    if (!handle) return 0;
    HReduce *h = static_cast<HReduce *>(handle);
    if (h->module && driver().ok) driver().ModuleUnload(h->module);
    delete h;
    return 0;
}

extern "C" int ksp_hreduce(void *stream, void *handle, const void *src, void *dest, int64_t rows,
                           int64_t src_stride, int64_t first_col, int64_t n_cols)
{
    if (!handle) return KSP_EINVAL;
    HReduce *h = static_cast<HReduce *>(handle);
    if (rows < 0 || n_cols <= 0 || first_col < 0 || src_stride < first_col + n_cols) return KSP_EINVAL;
    if (rows == 0) return 0;
    if (!src || !dest) return KSP_EINVAL;
    if ((uintptr_t) src % h->elem_size || (uintptr_t) dest % h->elem_size) {
        // element alignment only matters for power-of-two sizes; others are read bytewise by the compiler
        if ((h->elem_size & (h->elem_size - 1)) == 0) return KSP_EALIGN;
    }
    const int64_t blocks = ksp_divup(rows, (int64_t) 8);
    if (blocks > 0x7fffffff) return KSP_ETOOLARGE;
    long long a_rows = rows, a_stride = src_stride, a_first = first_col, a_cols = n_cols;
    void *args[] = {(void *) &src, (void *) &dest, &a_rows, &a_stride, &a_first, &a_cols};
    const CUresult cr = driver().LaunchKernel(h->function, (unsigned) blocks, 1, 1, 256, 1, 1, 0,
                                              (CUstream) stream, args, nullptr);
    if (cr != CUDA_SUCCESS) return (int) cudaErrorLaunchFailure;
    ksp_count_launch();
    return 0;
}
