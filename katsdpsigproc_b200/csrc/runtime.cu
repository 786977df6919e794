// Runtime shims of the C ABI: thin, synchronisation-free wrappers over the CUDA
// runtime that stand in for the PyCUDA calls of reference cuda.py (see
// include/ksp_b200.h for the line-by-line mapping).
#include "common.cuh"
#include <string.h>

extern "C" {

int ksp_abi_version(void) { return KSP_ABI_VERSION; }

const char *ksp_error_string(int code)
{
    switch (code) {
    case 0: return "success";
    case KSP_EINVAL: return "invalid argument (size, null pointer or unsupported combination)";
    case KSP_EALIGN: return "pointer or stride is not aligned as the C ABI requires";
    case KSP_ETOOLARGE: return "argument exceeds a documented limit";
    case KSP_ESCRATCH: return "scratch buffer too small";
    case KSP_ENOJIT: return "run-time compiler (NVRTC) not available";
    case KSP_EJIT: return "run-time compilation failed (see ksp_jit_log)";
    case KSP_ETIMEOUT: return "the dataflow flagger abandoned a launch (a wait exceeded its limit)";
    default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t) code);
    return "unknown error";
}

int ksp_device_count(int *count)
{
    if (!count) return KSP_EINVAL;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        *count = 0;
        cudaGetLastError();
        return 0;
    }
    return (int) e;
}

int ksp_device_name(int device, char *buf, int buf_len)
{
    if (!buf || buf_len <= 0) return KSP_EINVAL;
    cudaDeviceProp prop;
    KSP_CUDA(cudaGetDeviceProperties(&prop, device));
    strncpy(buf, prop.name, (size_t) buf_len - 1);
    buf[buf_len - 1] = 0;
    return 0;
}

int ksp_device_pci_bus_id(int device, char *buf, int buf_len)
{
    if (!buf || buf_len < 13) return KSP_EINVAL;
    KSP_CUDA(cudaDeviceGetPCIBusId(buf, buf_len, device));
    return 0;
}

int ksp_device_attributes(int device, int *cc_major, int *cc_minor, int *sm_count, int *warp_size,
                          size_t *total_mem, int *l2_bytes)
{
    int v;
    if (cc_major) { KSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); *cc_major = v; }
    if (cc_minor) { KSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); *cc_minor = v; }
    if (sm_count) { KSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
    if (warp_size) { KSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrWarpSize, device)); *warp_size = v; }
    if (l2_bytes) { KSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); *l2_bytes = v; }
    if (total_mem) {
        cudaDeviceProp prop;
        KSP_CUDA(cudaGetDeviceProperties(&prop, device));
        *total_mem = prop.totalGlobalMem;
    }
    return 0;
}

int ksp_device_set(int device) { return (int) cudaSetDevice(device); }
int ksp_device_get(int *device) { return device ? (int) cudaGetDevice(device) : KSP_EINVAL; }

int ksp_versions(int *runtime_version, int *driver_version)
{
    if (runtime_version) KSP_CUDA(cudaRuntimeGetVersion(runtime_version));
    if (driver_version) KSP_CUDA(cudaDriverGetVersion(driver_version));
    return 0;
}

int ksp_malloc(void **ptr, size_t bytes)
{
    if (!ptr) return KSP_EINVAL;
    return (int) cudaMalloc(ptr, bytes ? bytes : 1);
}
int ksp_free(void *ptr) { return (int) cudaFree(ptr); }
int ksp_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return KSP_EINVAL;
    return (int) cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
}
int ksp_host_free(void *ptr) { return (int) cudaFreeHost(ptr); }

int ksp_stream_create(void **stream)
{
    if (!stream) return KSP_EINVAL;
    cudaStream_t s;
    KSP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *) s;
    return 0;
}
int ksp_stream_destroy(void *stream) { return (int) cudaStreamDestroy((cudaStream_t) stream); }
int ksp_stream_synchronize(void *stream) { return (int) cudaStreamSynchronize((cudaStream_t) stream); }
int ksp_stream_query(void *stream)
{
    cudaError_t e = cudaStreamQuery((cudaStream_t) stream);
    if (e == cudaErrorNotReady) cudaGetLastError();
    return (int) e;
}
int ksp_stream_wait_event(void *stream, void *event)
{
    return (int) cudaStreamWaitEvent((cudaStream_t) stream, (cudaEvent_t) event, 0);
}

int ksp_event_create(void **event, int timing)
{
    if (!event) return KSP_EINVAL;
    cudaEvent_t e;
    // BLOCKING_SYNC as in reference cuda.py:463 so that Event.wait() sleeps instead of spinning
    unsigned flags = cudaEventBlockingSync | (timing ? 0u : (unsigned) cudaEventDisableTiming);
    KSP_CUDA(cudaEventCreateWithFlags(&e, flags));
    *event = (void *) e;
    return 0;
}
int ksp_event_destroy(void *event) { return (int) cudaEventDestroy((cudaEvent_t) event); }
int ksp_event_record(void *event, void *stream)
{
    return (int) cudaEventRecord((cudaEvent_t) event, (cudaStream_t) stream);
}
int ksp_event_synchronize(void *event) { return (int) cudaEventSynchronize((cudaEvent_t) event); }
int ksp_event_elapsed_ms(void *start, void *end, float *ms)
{
    if (!ms) return KSP_EINVAL;
    return (int) cudaEventElapsedTime(ms, (cudaEvent_t) start, (cudaEvent_t) end);
}

static int copy_kind(int kind, cudaMemcpyKind *out)
{
    switch (kind) {
    case 1: *out = cudaMemcpyHostToDevice; return 0;
    case 2: *out = cudaMemcpyDeviceToHost; return 0;
    case 3: *out = cudaMemcpyDeviceToDevice; return 0;
    default: return KSP_EINVAL;
    }
}

int ksp_memcpy_async(void *dst, const void *src, size_t bytes, int kind, void *stream)
{
    cudaMemcpyKind k;
    if (copy_kind(kind, &k)) return KSP_EINVAL;
    if (bytes == 0) return 0;
    return (int) cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t) stream);
}

int ksp_memcpy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                        size_t width_bytes, size_t height, int kind, void *stream)
{
    cudaMemcpyKind k;
    if (copy_kind(kind, &k)) return KSP_EINVAL;
    if (width_bytes == 0 || height == 0) return 0;
    return (int) cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, k,
                                   (cudaStream_t) stream);
}

int ksp_memset_async(void *dst, int value, size_t bytes, void *stream)
{
    if (bytes == 0) return 0;
    return (int) cudaMemsetAsync(dst, value, bytes, (cudaStream_t) stream);
}

}  // extern "C"

// ---------------------------------------------------------------- launch counter / profiling
#include <atomic>
#include <mutex>
#include <vector>

static std::atomic<unsigned long long> g_launches{0};
void ksp_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct StageSpan { int stage; cudaEvent_t begin, end; };
std::atomic<bool> g_profile{false};
std::mutex g_profile_mutex;
std::vector<StageSpan> g_spans;
std::vector<cudaEvent_t> g_open(KSP_STAGE_COUNT, nullptr);
}  // namespace

bool ksp_profile_active() { return g_profile.load(std::memory_order_relaxed); }

void ksp_profile_begin(int stage, cudaStream_t s)
{
    if (!ksp_profile_active() || stage < 0 || stage >= KSP_STAGE_COUNT) return;
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_open[stage] = e;
}

void ksp_profile_end(int stage, cudaStream_t s)
{
    if (!ksp_profile_active() || stage < 0 || stage >= KSP_STAGE_COUNT) return;
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    if (!g_open[stage]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_spans.push_back({stage, g_open[stage], e});
    g_open[stage] = nullptr;
}

extern "C" int ksp_kernel_launch_count(unsigned long long *count)
{
    if (!count) return KSP_EINVAL;
    *count = g_launches.load(std::memory_order_relaxed);
    return 0;
}

extern "C" int ksp_profile_enable(int on)
{
    g_profile.store(on != 0);
    return 0;
}

extern "C" int ksp_profile_read(double *stage_ms, int *stage_launches, int n_stages)
{
    if (!stage_ms || !stage_launches || n_stages < KSP_STAGE_COUNT) return KSP_EINVAL;
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    for (int i = 0; i < n_stages; i++) { stage_ms[i] = 0.0; stage_launches[i] = 0; }
    int rc = 0;
    for (const StageSpan &sp : g_spans) {
        float ms = 0.0f;
        cudaError_t e = cudaEventSynchronize(sp.end);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.begin, sp.end);
        if (e != cudaSuccess) rc = (int) e;
        stage_ms[sp.stage] += ms;
        stage_launches[sp.stage] += 1;
        cudaEventDestroy(sp.begin);
        cudaEventDestroy(sp.end);
    }
    g_spans.clear();
    return rc;
}

int ksp_sm_count()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = v;
    }
    return cached[dev];
}

int ksp_l2_bytes()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (!cached[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev);
        cached[dev] = v ? v : 1;
    }
    return cached[dev];
}
