// EXPERIMENT, NOT BUILT (kept for the record).  Measured on B200
// (r02g, 32768 x 8320): the chunked flagger with this kernel in place of madnz_stream_kernel +
// the two threshold passes takes 1.462 ms per dump against 1.108 ms - flags and noise identical
// (116 GPU parity tests green with KSP_ROWFUSE=1).  The row's second read does hit the L2, but the
// threshold phase keeps only 160 of the block's 256 threads busy, 4 blocks per SM, against nine
// 128-thread blocks per SM of the stand-alone lean pass, and every span pays the block-wide
// barriers of both phases.  To try it again: put the file back into csrc/, add it to the Makefile,
// export the fallback counter's address from madnz.cu (ksp_selection_fallbacks_ptr) and call
// ksp_noise_threshold_packed from flagger.cu in place of the noise and threshold stages.
//
// Noise estimate and SumThreshold of a baseline-major row in ONE kernel (fused flagger).
//
// Replaces the pair reference rfi/madnz_t.mako:72-87 + rfi/threshold_sum.mako:49-132 as the
// chunked flagger launches them (madnz_stream_kernel, then the two threshold passes): both walk
// the same row of deviations, the second needs the first's single number.  Here one block of 256
// threads takes a row, streams it once from device memory for the noise (madnz_stream_row, the
// device code of madnz_stream_kernel) and goes over it again, span by span, for the thresholds
// (ts_process_tile, the device code of threshold_sum_kernel) while its 128 KB are still in the
// L2: the spans are staged by TMA tile loads (128-byte swizzle, two buffers, the next span in
// flight while this one is worked on) and the rare span where a larger window might fire runs the
// full algorithm in place.  One launch instead of three per chunk, one read of the deviations from
// device memory instead of two.
#include "common.cuh"
#include "tma.cuh"
#include <stdlib.h>
#include <string.h>

#include "madnz_stream.cuh"
#include "threshold_tile.cuh"

int ksp_selection_fallbacks_ptr(unsigned long long **out);

namespace {

constexpr int RF_THREADS = 256;
constexpr int RF_BLOCKS_PER_SM = 4;
constexpr int RF_SPAN_RUNS = 160;        // runs of 32 channels per staged span: two buffers fit

struct RfArgs {
    const float *dev_t;
    float *noise;
    uint32_t *bits_t;
    unsigned long long *fallbacks;
    int64_t dev_stride, words_stride;
    int channels;
    int T, n_chunks, chunk_valid, edge;
    int n_windows;
    double n_sigma;
    double scales[TS_MAX_WINDOWS];
};

constexpr size_t rf_thr_smem(int T)
{
    return 2 * (size_t) ((((T + 2) * PITCH + 255) / 256) * 256) * 4 + (size_t) (RF_THREADS + 2) * 16 +
           (size_t) (RF_THREADS + 2) * 4 + (size_t) RF_THREADS * 8 + TS_THR_WORDS * 4;
}
constexpr size_t rf_max(size_t x, size_t y) { return x > y ? x : y; }
constexpr size_t RF_SMEM = 1024 + rf_max((size_t) MS_SMEM_WORDS * 4, rf_thr_smem(RF_SPAN_RUNS));
static_assert(RF_SMEM <= 56 * 1024, "four blocks per SM");

__shared__ RfArgs s_rf;
__shared__ uint64_t s_mbar[2];
extern __shared__ __align__(1024) uint8_t rf_sm_raw[];

__device__ __forceinline__ uint8_t *rf_smem()
{
    return rf_sm_raw + ((1024u - (smem_u32(rf_sm_raw) & 1023u)) & 1023u);
}

// The phases are real calls: each gets the block's whole register budget.
__device__ __noinline__ void rf_noise(int64_t row)
{
    const RfArgs &a = s_rf;
    const bool vec_ok = ((a.dev_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.dev_t) & 15) == 0);
    madnz_stream_row<false>(a.dev_t + row * a.dev_stride, a.noise + row, a.channels, vec_ok,
                            reinterpret_cast<uint32_t *>(rf_smem()), a.fallbacks);
}

__device__ __noinline__ uint32_t rf_threshold_full(const TsTile &tl)
{
    uint32_t F = 0;
    __syncthreads();                       // the vote's readers are done with stat / Fsm
    ts_process_tile<false>(tl, F);
    return F;
}

__device__ __noinline__ void rf_threshold(const CUtensorMap *tmap, int64_t row)
{
    const RfArgs &a = s_rf;
    uint8_t *sm = rf_smem();
    const int tid = threadIdx.x;
    const int T = a.T;
    // layout: two spans of (T + 2) runs, statistics, flag words, carries, thresholds
    const int buf_floats = (((T + 2) * PITCH + 255) / 256) * 256;
    float *rowbuf0 = reinterpret_cast<float *>(sm);
    float4 *stat = reinterpret_cast<float4 *>(rowbuf0 + 2 * buf_floats);         // T + 2
    uint32_t *Fsm = reinterpret_cast<uint32_t *>(stat + RF_THREADS + 2);         // T + 2
    uint32_t *car1 = Fsm + RF_THREADS + 2;
    uint32_t *car2 = car1 + RF_THREADS;
    float *thr = reinterpret_cast<float *>(car2 + RF_THREADS);                   // TS_THR_WORDS
    // this shared memory was last written with ordinary stores (the noise phase): order them
    // before the TMA writes of the spans
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();                       // also: the row's noise (global memory) is visible to the block
    if (tid == 0) {
        for (int y = 0; y < 2 && y < a.n_chunks; y++) {
            mbar_expect_tx(&s_mbar[y], (uint32_t) T * RUN * 4u);
            tma_load_3d(rowbuf0 + y * buf_floats, tmap, 0, (y * a.chunk_valid - a.edge) >> 5, (int) row,
                        &s_mbar[y]);
        }
    }
    if (tid >= 32 && tid < 34) {
        Fsm[T + tid - 32] = 0u;
        stat[T + tid - 32] = make_float4(-__int_as_float(0x7f800000), 0.0f, 0.0f, 0.0f);
    }
    if (tid >= 64 && tid < 64 + 2 * PITCH) {               // two runs of zeros past each span
        rowbuf0[T * PITCH + tid - 64] = 0.0f;
        rowbuf0[buf_floats + T * PITCH + tid - 64] = 0.0f;
    }
    if (tid < 32)
        ts_thresholds(thr, tid, a.n_windows, a.n_sigma, __ldcg(a.noise + row), a.scales, a.channels);
    const int C = a.channels;
    uint32_t *bits_row = a.bits_t + row * a.words_stride;
    uint32_t parity = 0;
    for (int y = 0; y < a.n_chunks; y++) {
        float *rowbuf = rowbuf0 + (y & 1) * buf_floats;
        mbar_wait(&s_mbar[y & 1], (parity >> (y & 1)) & 1u);
        parity ^= 1u << (y & 1);
        __syncthreads();                   // thr[] and the presets visible; the previous span's statistics are free
        const int base = y * a.chunk_valid - a.edge;               // row channel of slot 0
        const int64_t pos0 = (int64_t) base + (int64_t) tid * RUN;
        uint32_t F = 0;
        TsTile tl;
        tl.rowbuf = rowbuf; tl.stat = stat; tl.Fsm = Fsm; tl.car1 = car1; tl.car2 = car2; tl.thr = thr;
        tl.T = T; tl.span = T * RUN; tl.C = C; tl.n_windows = a.n_windows; tl.pos0 = pos0;
        if (ts_process_tile<true>(tl, F)) F = rf_threshold_full(tl);   // block-uniform
        const int64_t out_lo = (int64_t) y * a.chunk_valid;
        const int64_t out_hi = min((int64_t) C, out_lo + (int64_t) a.chunk_valid);
        if (tid < T && pos0 >= out_lo && pos0 < out_hi)
            bits_row[pos0 >> 5] = F & bit_range(-pos0, (int64_t) C - pos0);
        __syncthreads();                   // everybody is done with this buffer
        if (tid == 0 && y + 2 < a.n_chunks) {
            mbar_expect_tx(&s_mbar[y & 1], (uint32_t) T * RUN * 4u);
            tma_load_3d(rowbuf, tmap, 0, ((y + 2) * a.chunk_valid - a.edge) >> 5, (int) row, &s_mbar[y & 1]);
        }
    }
}

__global__ void __launch_bounds__(RF_THREADS, RF_BLOCKS_PER_SM)
noise_threshold_kernel(const RfArgs a, const __grid_constant__ CUtensorMap tmap)
{
    if (threadIdx.x == 0) {
        s_rf = a;
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
    }
    __syncthreads();
    const int64_t row = blockIdx.x;
    rf_noise(row);
    rf_threshold(&tmap, row);
}

}  // namespace

bool ksp_rowfuse_legal(int64_t channels, int64_t baselines, const float *dev_t, int64_t dev_stride,
                       int n_windows)
{
    if (n_windows < 1 || n_windows > TS_MAX_WINDOWS) return false;
    if (channels % RUN != 0 || channels < RUN || channels > (int64_t) 1 << 24) return false;
    if (baselines < 1 || baselines > 0x7fffffff) return false;
    if (dev_stride % 4 != 0 || (uintptr_t) dev_t % 16 != 0) return false;
    return tensor_map_encoder() != nullptr;
}

// noise[b] and the bit-packed flags of every row in one launch (see the head of the file)
int ksp_noise_threshold_packed(cudaStream_t s, const float *dev_t, float *noise, uint32_t *bits_t,
                               int64_t channels, int64_t baselines, int64_t dev_stride,
                               int64_t words_stride, int n_windows, double n_sigma, const double *scales)
{
    if (!ksp_rowfuse_legal(channels, baselines, dev_t, dev_stride, n_windows)) return KSP_EINVAL;
    if (!dev_t || !noise || !bits_t || !scales) return KSP_EINVAL;
    if (dev_stride < channels || words_stride < channels / 32) return KSP_EINVAL;
    RfArgs a;
    memset(&a, 0, sizeof(a));
    a.dev_t = dev_t; a.noise = noise; a.bits_t = bits_t;
    {
        int rc = ksp_selection_fallbacks_ptr(&a.fallbacks);
        if (rc) return rc;
    }
    a.dev_stride = dev_stride; a.words_stride = words_stride;
    a.channels = (int) channels;
    a.n_windows = n_windows; a.n_sigma = n_sigma;
    for (int w = 0; w < TS_MAX_WINDOWS; w++) a.scales[w] = w < n_windows ? scales[w] : 0.0;
    const int64_t runs = channels / RUN;
    if (runs <= RF_SPAN_RUNS) {
        a.T = (int) (ksp_divup(runs, 32) * 32);
        a.edge = 0;
        a.chunk_valid = a.T * RUN;
        a.n_chunks = 1;
    } else {
        const int reach = (1 << n_windows) - n_windows - 1;          // influence radius of a sample
        a.edge = (int) (ksp_divup(reach, RUN) * RUN);
        const int max_valid = RF_SPAN_RUNS * RUN - 2 * a.edge;
        const int n = (int) ksp_divup(channels, max_valid);
        const int valid = (int) (ksp_divup(ksp_divup(channels, n), RUN) * RUN);
        a.T = (int) (ksp_divup((valid + 2 * a.edge) / RUN, 32) * 32);
        a.chunk_valid = a.T * RUN - 2 * a.edge;
        a.n_chunks = (int) ksp_divup(channels, a.chunk_valid);
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    {
        const cuuint64_t dims[3] = {(cuuint64_t) RUN, (cuuint64_t) runs, (cuuint64_t) baselines};
        const cuuint64_t strides[2] = {RUN * sizeof(float), (cuuint64_t) dev_stride * sizeof(float)};
        const cuuint32_t box[3] = {RUN, (cuuint32_t) a.T, 1};
        const cuuint32_t elem[3] = {1, 1, 1};
        CUresult rc = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *) dev_t, dims,
                                           strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return KSP_EINVAL;
    }
    static bool configured[64];
    int dev = 0;
    KSP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        KSP_CUDA(cudaFuncSetAttribute(noise_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) RF_SMEM));
        configured[dev] = true;
    }
    noise_threshold_kernel<<<(unsigned) baselines, RF_THREADS, RF_SMEM, s>>>(a, tmap);
    KSP_CHECK_LAUNCH();
    return 0;
}
