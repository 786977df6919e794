// Fused flagger: the standard median + MAD + SumThreshold combination of
// reference rfi/device.py:1111-1166 (5 launches, 31 B/vis of HBM traffic).  Two forms: the
// dataflow kernel of dataflow.cu (one persistent launch per dump, the default where it applies)
// and, in this file, the chunked form, 4 launches per CHUNK of baselines:
//
//   vis[:, chunk] --bg13_kernel--> dev_t (chunk x C float32, scratch)
//                 --madnz_stream_kernel--> noise[chunk]
//                 --threshold_sum_kernel (packed)--> bits_t (chunk x C/32 words, scratch)
//                 --expand_flags_kernel--> flags[:, chunk]
//
// The float transpose of the reference is folded into the background kernel's
// store, the uchar transpose into the bit -> byte expansion.  Consecutive chunks
// run on separate internal streams ("lanes", each with its own scratch) that fork
// from and join into the caller's stream, so kernel tails overlap.
#include "common.cuh"
#include "dataflow.h"
#include <stdlib.h>

int ksp_threshold_sum_packed(cudaStream_t s, const float *dev_t, const float *noise,
                             uint32_t *bits_t, int64_t channels, int64_t baselines,
                             int64_t dev_stride, int64_t words_stride, int n_windows,
                             double n_sigma, const double *scales, uint32_t *work);
size_t ksp_threshold_work_bytes(int64_t channels, int64_t baselines);
int ksp_expand_flags(cudaStream_t s, const uint32_t *bits_t, uint8_t *flags, int64_t channels,
                     int64_t baselines, int64_t words_stride, int64_t flags_stride, int flag_value);

namespace {

struct Layout {
    int64_t chunk;        // baselines per chunk (multiple of 32)
    int64_t dev_stride;   // floats per baseline row of dev_t
    int64_t words_stride; // words per baseline row of bits_t
    size_t dev_bytes, bits_bytes, work_bytes;   // per lane
    int lanes;            // chunks in flight (1 = everything on the caller's stream)
};

constexpr int MAX_LANES = 8;

// Internal streams and events for running several chunks at a time (per thread and device,
// created on first use, released when the thread ends).
struct LanePool {
    bool ready = false;
    int device = -1;
    cudaStream_t stream[MAX_LANES];
    cudaEvent_t start, done[MAX_LANES];
    // a thread that goes away gives its streams and events back (at process exit the runtime may
    // already be unloading: the calls then fail harmlessly)
    ~LanePool()
    {
        if (!ready) return;
        int current = 0;
        if (cudaGetDevice(&current) != cudaSuccess) return;
        if (cudaSetDevice(device) != cudaSuccess) return;
        for (int i = 0; i < MAX_LANES; i++) {
            cudaStreamDestroy(stream[i]);
            cudaEventDestroy(done[i]);
        }
        cudaEventDestroy(start);
        cudaSetDevice(current);
    }
};

int get_pool(LanePool **out)
{
    static thread_local LanePool pools[64];   // per thread: calls from different threads never share events
    int dev = 0;
    KSP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return KSP_EINVAL;
    LanePool &pool = pools[dev];
    if (!pool.ready) {
        for (int i = 0; i < MAX_LANES; i++) {
            KSP_CUDA(cudaStreamCreateWithFlags(&pool.stream[i], cudaStreamNonBlocking));
            KSP_CUDA(cudaEventCreateWithFlags(&pool.done[i], cudaEventDisableTiming));
        }
        KSP_CUDA(cudaEventCreateWithFlags(&pool.start, cudaEventDisableTiming));
        pool.device = dev;
        pool.ready = true;
    }
    *out = &pool;
    return 0;
}

Layout make_layout(const ksp_flagger_params *p)
{
    Layout l;
    l.dev_stride = ksp_divup(p->channels, 32) * 32;
    l.words_stride = ksp_divup(ksp_divup(p->channels, 32), 4) * 4;
    // Chunks in flight ("lanes", KSP_LANES, default 4): consecutive chunks run on separate
    // internal streams so that the tail of one kernel overlaps the next chunk's kernels.
    // Measured on B200 (profiles/): the stages are issue-bound rather than HBM-bound and launches
    // cost ~9 us each, so few large chunks beat many L2-sized ones (KSP_CHUNK / chunk_baselines
    // override the default below).
    // (the environment is read ONCE per process: the scratch size, queried when the operation is
    // instantiated, and every later launch must see the same layout)
    static const int env_lanes = [] {
        const char *e = getenv("KSP_LANES");
        int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > MAX_LANES ? MAX_LANES : v);
    }();
    static const int64_t env_chunk = [] {
        const char *c = getenv("KSP_CHUNK");
        return c ? (int64_t) atoll(c) : (int64_t) 0;
    }();
    int lanes = env_lanes;
    int64_t chunk = p->chunk_baselines;
    if (chunk <= 0) {
        chunk = env_chunk;
        if (chunk <= 0) {
            // Whole waves: with 16 baselines per SM a chunk is exactly 16 waves of the
            // background kernel (4 blocks of 32 baselines x 256 channels per SM, 32768 channels)
            // and 4 waves of the noise kernel (4 one-row blocks per SM), so no launch ends on a
            // nearly empty wave.  Take the multiple of that unit nearest to an even split over
            // the lanes; shards smaller than one unit are split evenly instead.
            const int64_t unit = 16 * (int64_t) ksp_sm_count();
            if (p->baselines >= unit) {
                // as many units per chunk as it takes for every chunk to get a lane of its own
                // (then the background filter is one launch, see ksp_flagger), up to 4 units
                int64_t k = ksp_divup(p->baselines, (int64_t) lanes * unit);
                chunk = unit * (k < 1 ? 1 : (k > 4 ? 4 : k));
            } else {
                // a shard smaller than one unit (e.g. 1620 baselines of a 12960-baseline dump on
                // 8 GPUs): one or two launches per stage, not `lanes` sub-wave ones
                const int n = p->baselines * 2 > unit ? 2 : 1;
                chunk = ksp_divup(ksp_divup(p->baselines, n), 32) * 32;
            }
        }
    }
    chunk = (chunk / 32) * 32;
    if (chunk < 32) chunk = 32;
    const int64_t bl_pad = ksp_divup(p->baselines, 32) * 32;
    if (chunk > bl_pad) chunk = bl_pad > 0 ? bl_pad : 32;
    l.chunk = chunk;
    l.dev_bytes = (size_t) chunk * (size_t) l.dev_stride * 4;
    l.bits_bytes = (size_t) chunk * (size_t) l.words_stride * 4;
    l.work_bytes = (ksp_threshold_work_bytes(p->channels, chunk) + 1023) / 1024 * 1024;   // keeps every lane 1 KB aligned
    const int64_t n_chunks = ksp_divup(p->baselines, chunk);
    if (lanes > n_chunks) lanes = (int) n_chunks;
    l.lanes = lanes;
    return l;
}

}  // namespace

extern "C" size_t ksp_flagger_scratch_bytes(const ksp_flagger_params *p)
{
    if (!p || p->channels <= 0 || p->baselines <= 0) return 0;
    if (ksp_dataflow_applies(p)) return ksp_dataflow_scratch_bytes(p);
    Layout l = make_layout(p);
    return (l.dev_bytes + l.bits_bytes + l.work_bytes) * (size_t) l.lanes;
}

extern "C" int ksp_flagger_is_dataflow(const ksp_flagger_params *p)
{
    return (p && p->channels > 0 && p->baselines > 0 && ksp_dataflow_applies(p)) ? 1 : 0;
}

extern "C" int ksp_flagger_stats(void *stream, const ksp_flagger_params *p, const void *scratch,
                                 unsigned long long *out, int n)
{
    if (!p || !out || n < 0) return KSP_EINVAL;
    for (int i = 0; i < n; i++) out[i] = 0;
    if (p->channels <= 0 || p->baselines <= 0 || !ksp_dataflow_applies(p)) return 0;
    if (!scratch) return KSP_EINVAL;
    return ksp_dataflow_stats((cudaStream_t) stream, scratch, out, n);
}

extern "C" int64_t ksp_flagger_chunk_baselines(const ksp_flagger_params *p)
{
    if (!p || p->channels <= 0 || p->baselines <= 0) return 0;
    if (ksp_dataflow_applies(p)) return 32;                 // the dataflow kernel's strip
    return make_layout(p).chunk;
}

extern "C" int ksp_flagger(void *stream, const ksp_flagger_params *p, const void *vis,
                           const uint8_t *input_flags, float *noise, uint8_t *flags, void *scratch,
                           size_t scratch_bytes)
{
    if (!p) return KSP_EINVAL;
    if (p->channels < 0 || p->baselines < 0) return KSP_EINVAL;
    if (p->channels == 0 || p->baselines == 0) return 0;
    if (!vis || !noise || !flags || !scratch) return KSP_EINVAL;
    if (p->n_windows < 1) return KSP_EINVAL;
    if (p->flag_mode != KSP_FLAGS_NONE && !input_flags) return KSP_EINVAL;
    if (p->flags_stride < p->baselines || p->vis_stride < p->baselines) return KSP_EINVAL;
    if (p->chunk_baselines < 0 && !ksp_dataflow_legal(p)) return KSP_EINVAL;
    if (ksp_dataflow_applies(p))
        return ksp_dataflow_flagger((cudaStream_t) stream, p, vis, input_flags, noise, flags, scratch,
                                    scratch_bytes);
    Layout l = make_layout(p);
    const size_t lane_bytes = l.dev_bytes + l.bits_bytes + l.work_bytes;
    if (scratch_bytes < lane_bytes * (size_t) l.lanes) return KSP_ESCRATCH;
    if ((uintptr_t) scratch % 16) return KSP_EALIGN;
    cudaStream_t user = (cudaStream_t) stream;
    const size_t vis_elem = p->is_amplitude ? 4 : 8;

    // With several lanes, chunk i runs on internal stream i % lanes with its own scratch; the
    // internal streams fork from and join back into the caller's stream through events, so
    // the call keeps plain stream semantics.
    // When every chunk has a lane of its own, the background filter of ALL chunks can be ONE launch
    // on the caller's stream - the lanes' deviations are contiguous in the scratch - with only the
    // later stages chunk by chunk on the lanes: one ramp and one tail instead of one per chunk
    // (0.61 against 0.72 ms at 8320 baselines with the stages timed one after the other).  Measured
    // with 4 lanes (profiles/r02g_bg_whole.txt): the same at 8320 baselines (1.107 ms, 13 launches
    // instead of 16), 1.5 % faster at 12960, but 1 - 6 % SLOWER for the shards of a dump split over
    // 2 - 8 GPUs (6496 ... 1632 baselines), where the first chunk's noise estimate had better
    // start under the second chunk's background tiles.  So: dumps of at least 3 chunk units with
    // the library's own chunking; a caller who fixes chunk_baselines gets it whenever the chunks
    // fit the lanes.  KSP_BG_WHOLE=0 / 1: never / wherever the chunks fit the lanes.
    static const int env_bg_whole = [] {
        const char *e = getenv("KSP_BG_WHOLE");
        return e ? atoi(e) : -1;
    }();
    const bool fits = ksp_divup(p->baselines, l.chunk) <= l.lanes;
    const bool large = p->chunk_baselines > 0 || p->baselines >= 3 * 16 * (int64_t) ksp_sm_count();
    const bool bg_whole = fits && (env_bg_whole < 0 ? large : env_bg_whole != 0);
    if (bg_whole) {
        ksp_profile_begin(KSP_STAGE_BACKGROUND, user);
        int rc = ksp_background_median_filter_t(user, vis, (float *) scratch, input_flags, p->channels,
                                                p->baselines, p->vis_stride, l.dev_stride,
                                                p->input_flags_stride, p->width, p->is_amplitude,
                                                p->flag_mode, p->abs_mode);
        ksp_profile_end(KSP_STAGE_BACKGROUND, user);
        if (rc) return rc;
    }
    LanePool *pool = nullptr;
    if (l.lanes > 1 && !ksp_profile_active()) {   // stage timing needs the stages one after another: no lanes
        int rc = get_pool(&pool);
        if (rc) return rc;
        KSP_CUDA(cudaEventRecord(pool->start, user));
        for (int i = 0; i < l.lanes; i++) KSP_CUDA(cudaStreamWaitEvent(pool->stream[i], pool->start, 0));
    }

    int64_t index = 0;
    int status = 0;                        // first error; the lanes are joined before it is returned
    for (int64_t b0 = 0; b0 < p->baselines && !status; b0 += l.chunk, index++) {
        const int lane = (int) (index % l.lanes);
        cudaStream_t s = pool ? pool->stream[lane] : user;
        // scratch: the lanes' deviations one after the other, then (flag words, tile list) per lane
        float *dev_t = (float *) ((char *) scratch + l.dev_bytes * (size_t) lane);
        uint32_t *bits_t = (uint32_t *) ((char *) scratch + l.dev_bytes * (size_t) l.lanes +
                                         (l.bits_bytes + l.work_bytes) * (size_t) lane);
        uint32_t *work = (uint32_t *) ((char *) bits_t + l.bits_bytes);
        const int64_t nb = (p->baselines - b0 < l.chunk) ? p->baselines - b0 : l.chunk;
        const void *vis_c = (const char *) vis + (size_t) b0 * vis_elem;
        const uint8_t *in_fl = input_flags;
        if (p->flag_mode == KSP_FLAGS_FULL) in_fl = input_flags + b0;
        int rc = 0;
        if (!bg_whole) {
            ksp_profile_begin(KSP_STAGE_BACKGROUND, s);
            rc = ksp_background_median_filter_t(s, vis_c, dev_t, in_fl, p->channels, nb,
                                                p->vis_stride, l.dev_stride, p->input_flags_stride,
                                                p->width, p->is_amplitude, p->flag_mode,
                                                p->abs_mode);
            ksp_profile_end(KSP_STAGE_BACKGROUND, s);
            if (rc) { status = rc; break; }
        }
        ksp_profile_begin(KSP_STAGE_NOISE, s);
        rc = ksp_madnz_t(s, dev_t, noise + b0, p->channels, nb, l.dev_stride);
        ksp_profile_end(KSP_STAGE_NOISE, s);
        if (rc) { status = rc; break; }
        ksp_profile_begin(KSP_STAGE_THRESHOLD, s);
        rc = ksp_threshold_sum_packed(s, dev_t, noise + b0, bits_t, p->channels, nb, l.dev_stride,
                                      l.words_stride, p->n_windows, p->n_sigma, p->scales, work);
        ksp_profile_end(KSP_STAGE_THRESHOLD, s);
        if (rc) { status = rc; break; }
        ksp_profile_begin(KSP_STAGE_EXPAND, s);
        rc = ksp_expand_flags(s, bits_t, flags + b0, p->channels, nb, l.words_stride,
                              p->flags_stride, p->flag_value);
        ksp_profile_end(KSP_STAGE_EXPAND, s);
        if (rc) { status = rc; break; }
    }
    if (pool) {
        // always: later work on the caller's stream must not overtake kernels still running on a lane
        for (int i = 0; i < l.lanes; i++) {
            cudaError_t e = cudaEventRecord(pool->done[i], pool->stream[i]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(user, pool->done[i], 0);
            if (e != cudaSuccess && !status) status = (int) e;
        }
    }
    return status;
}
