// Channel-axis sliding-median background filter.
//
// Replaces reference rfi/background_median_filter.mako:200-220 (+ the float
// transpose of rfi/device.py:1152-1157 in the _t variant).  Semantics are those
// of the HOST class (rfi/host.py:133-151): windows clipped to
// the band, flagged / NaN samples skipped, even counts averaged in float64,
// flagged centre -> 0.
//
// Width 13 (the MeerKAT setting) has a dedicated tiled kernel, described at
// bg13_kernel below; other widths use a simple per-thread ring buffer.
#include "common.cuh"
#include "median13.cuh"
#include "bg13.cuh"

namespace {

template <int IN_MODE, int FLAG_MODE, bool TRANSPOSED, int TC>
__global__ void __launch_bounds__(BG_THREADS, BG_MIN_BLOCKS)
bg13_kernel(const BgArgs a)
{
    extern __shared__ __align__(16) float amp_sm[];     // [TILE_B][P]
    bg13_tile<IN_MODE, FLAG_MODE, TRANSPOSED, TC>(a, BlockTile(), amp_sm);
}

// ---------------------------------------------------------------- any odd width
// One thread per (baseline, channel segment); window kept in local memory.
// Only used when width != 13.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(BG_THREADS)
bg_generic_kernel(const BgArgs a, int width, int in_mode, int flag_mode)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.baselines) return;
    const int64_t c_begin = (int64_t) blockIdx.y * a.seg;
    const int64_t c_end = min(a.channels, c_begin + (int64_t) a.seg);
    const int half = width / 2;
    float win[KSP_MAX_WIDTH];   // ring buffer of amplitudes (NaN = unusable)
    float sorted[KSP_MAX_WIDTH];
    const float nan = __int_as_float(0x7fc00000);

    auto fetch = [&](int64_t c) -> float {
        if (c < 0 || c >= a.channels) return nan;
        float amp;
        int64_t idx = c * a.vis_stride + b;
        if (in_mode == IN_AMP) {
            amp = reinterpret_cast<const float *>(a.vis)[idx];
        } else {
            float2 v = reinterpret_cast<const float2 *>(a.vis)[idx];
            amp = abs_c64_rt(v.x, v.y, in_mode == IN_NUMPY ? KSP_ABS_NUMPY : KSP_ABS_HYPOT);
        }
        if (flag_mode == KSP_FLAGS_CHANNEL && a.flags[c]) amp = nan;
        if (flag_mode == KSP_FLAGS_FULL && a.flags[c * a.flags_stride + b]) amp = nan;
        return amp;
    };

    // ring slot of channel c is (c - (c_begin - half)) % width
    for (int k = 0; k < width - 1; k++) win[k] = fetch(c_begin - half + k);
    int head = width - 1;  // slot that receives channel c + half
    for (int64_t c = c_begin; c < c_end; c++) {
        win[head] = fetch(c + half);
        head = (head + 1 == width) ? 0 : head + 1;
        // centre sample: slot of channel c
        int centre = head + half;  // head now indexes channel c - half
        if (centre >= width) centre -= width;
        float amp = win[centre];
        float out = 0.0f;
        if (amp == amp) {
            int n = 0;
            for (int k = 0; k < width; k++) {
                float v = win[k];
                if (v == v) {
                    int i = n;
                    while (i > 0 && sorted[i - 1] > v) {
                        sorted[i] = sorted[i - 1];
                        i--;
                    }
                    sorted[i] = v;
                    n++;
                }
            }
            double med = (n & 1) ? (double) sorted[n / 2]
                                 : ((double) sorted[n / 2 - 1] + (double) sorted[n / 2]) * 0.5;
            out = __double2float_rn((double) amp - med);
        }
        if (TRANSPOSED)
            a.out[b * a.out_stride + c] = out;
        else
            a.out[c * a.out_stride + b] = out;
    }
}

// ---------------------------------------------------------------- odd widths up to 31, tiled
// Same tiling as bg13_kernel (32 baselines x 256 channels of amplitudes in shared memory, NaN
// marks unusable samples), but phase 2 keeps a SORTED WINDOW in registers and slides it:
// thread (baseline, segment of 32 channels) sorts the first window once, then for every output
// removes the sample that leaves and inserts the one that enters, branch-free in 4 operations
// per element (t_i = a_i < out ? a_i : a_(i+1);  new_i = max(t_(i-1), min(in, t_i))), and reads
// the median off the middle register.  Threads whose channel range touches an unusable sample
// (band edges, flags, NaN) use a per-output insertion sort over the usable samples instead.
// The deviations go through a second shared tile so that global stores are coalesced in both
// output layouts.  ~4 W operations per output: 30-100x faster than bg_generic_kernel.
constexpr int BGW_TC = 256;            // channels per tile
constexpr int BGW_SEG = 32;            // consecutive outputs per thread

template <int W>
struct WGeom {
    static constexpr int H = W / 2;
    static constexpr int HP = ((H + 3) / 4) * 4;                   // halo, whole float4s
    static constexpr int RAW = BGW_TC + 2 * HP;
    // odd pitches: all shared-memory traffic here is scalar with lane <-> baseline
    static constexpr int P = RAW | 1;
    static constexpr int OP = BGW_TC + 1;
    static constexpr int SMEM_BYTES = TILE_B * (P + OP) * 4;
};

template <int W, bool TRANSPOSED>
__global__ void __launch_bounds__(BG_THREADS)
bgw_kernel(const BgArgs a, int in_mode, int flag_mode)
{
    using G = WGeom<W>;
    constexpr int H = G::H;
    extern __shared__ __align__(16) float bgw_sm[];
    float *amp = bgw_sm;                                  // [TILE_B][P]: channel c0 - HP + s
    float *dev = bgw_sm + TILE_B * G::P;                  // [TILE_B][OP]
    const int64_t b0 = (int64_t) blockIdx.x * TILE_B;
    const int c0 = (int) blockIdx.y * BGW_TC;
    const int C = (int) a.channels;
    const float nan = __int_as_float(0x7fc00000);

    // ---- phase 1: amplitudes of the tile and its halos (lane <-> baseline: coalesced rows)
    for (int idx = threadIdx.x; idx < TILE_B * G::RAW; idx += BG_THREADS) {
        const int bl = idx & (TILE_B - 1), sidx = idx / TILE_B;
        const int c = c0 - G::HP + sidx;
        const int64_t b = b0 + bl;
        float v = nan;
        if (c >= 0 && c < C && b < a.baselines) {
            const int64_t at = (int64_t) c * a.vis_stride + b;
            if (in_mode == IN_AMP) {
                v = reinterpret_cast<const float *>(a.vis)[at];
            } else {
                const float2 z = reinterpret_cast<const float2 *>(a.vis)[at];
                v = abs_c64_rt(z.x, z.y, in_mode == IN_NUMPY ? KSP_ABS_NUMPY : KSP_ABS_HYPOT);
            }
            if (flag_mode == KSP_FLAGS_CHANNEL && a.flags[c]) v = nan;
            if (flag_mode == KSP_FLAGS_FULL && a.flags[(int64_t) c * a.flags_stride + b]) v = nan;
        }
        amp[bl * G::P + sidx] = v;
    }
    __syncthreads();

    // ---- phase 2: thread = (baseline bl, segment seg of 32 outputs)
    {
        const int bl = threadIdx.x & (TILE_B - 1), seg = threadIdx.x / TILE_B;
        const float *row = amp + bl * G::P + G::HP + seg * BGW_SEG;      // row[k] = channel c0 + 32 seg + k
        float *out = dev + bl * G::OP + seg * BGW_SEG;
        // every sample the segment touches: row[-H .. SEG - 1 + H]
        bool clean = true;
        for (int k = -H; k < BGW_SEG + H; k++) clean &= (row[k] == row[k]);
        if (clean) {
            float sw[W];                                               // sorted window
#pragma unroll
            for (int i = 0; i < W; i++) sw[i] = __int_as_float(0x7f800000);
#pragma unroll 1
            for (int k = -H; k <= H; k++) {                            // insertion of the first window
                const float in = row[k];
                float prev = -__int_as_float(0x7f800000);
#pragma unroll
                for (int i = 0; i < W; i++) {
                    const float cur = sw[i];
                    sw[i] = fmaxf(prev, fminf(in, cur));
                    prev = cur;
                }
            }
#pragma unroll 1
            for (int k = 0; k < BGW_SEG; k++) {
                out[k] = row[k] - sw[H];
                if (k + 1 < BGW_SEG) {                                 // slide: row[k - H] leaves, row[k + H + 1] enters
                    const float gone = row[k - H], in = row[k + H + 1];
                    float t[W - 1];
#pragma unroll
                    for (int i = 0; i < W - 1; i++) t[i] = (sw[i] < gone) ? sw[i] : sw[i + 1];
                    float prev = -__int_as_float(0x7f800000);
#pragma unroll
                    for (int i = 0; i < W - 1; i++) {
                        sw[i] = fmaxf(prev, fminf(in, t[i]));
                        prev = t[i];
                    }
                    sw[W - 1] = fmaxf(prev, in);
                }
            }
        } else {
            // usable samples of each window, insertion-sorted (band edges, flags, NaN)
            for (int k = 0; k < BGW_SEG; k++) {
                const float centre = row[k];
                float o = 0.0f;
                if (centre == centre) {
                    float srt[W];
                    int n = 0;
                    for (int j = -H; j <= H; j++) {
                        const float v = row[k + j];
                        if (v == v) {
                            int i = n;
                            while (i > 0 && srt[i - 1] > v) {
                                srt[i] = srt[i - 1];
                                i--;
                            }
                            srt[i] = v;
                            n++;
                        }
                    }
                    const double med = (n & 1) ? (double) srt[n / 2]
                                               : ((double) srt[n / 2 - 1] + (double) srt[n / 2]) * 0.5;
                    o = __double2float_rn((double) centre - med);
                }
                out[k] = o;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: coalesced stores
    if (TRANSPOSED) {
        for (int idx = threadIdx.x; idx < TILE_B * BGW_TC; idx += BG_THREADS) {
            const int bl = idx / BGW_TC, k = idx % BGW_TC;                 // lanes along channels
            const int64_t b = b0 + bl;
            const int c = c0 + k;
            if (b < a.baselines && c < C) a.out[b * a.out_stride + c] = dev[bl * G::OP + k];
        }
    } else {
        for (int idx = threadIdx.x; idx < TILE_B * BGW_TC; idx += BG_THREADS) {
            const int bl = idx % TILE_B, k = idx / TILE_B;                 // lanes along baselines
            const int64_t b = b0 + bl;
            const int c = c0 + k;
            if (b < a.baselines && c < C) a.out[(int64_t) c * a.out_stride + b] = dev[bl * G::OP + k];
        }
    }
}

template <int W, bool TRANSPOSED>
int launch_bgw(cudaStream_t s, const BgArgs &a, int in_mode, int flag_mode)
{
    dim3 grid((unsigned) ksp_divup(a.baselines, TILE_B), (unsigned) ksp_divup(a.channels, BGW_TC));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    const size_t smem = WGeom<W>::SMEM_BYTES;
    KSP_CUDA(cudaFuncSetAttribute(bgw_kernel<W, TRANSPOSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int) smem));
    bgw_kernel<W, TRANSPOSED><<<grid, BG_THREADS, smem, s>>>(a, in_mode, flag_mode);
    KSP_CHECK_LAUNCH();
    return 0;
}

int pick_segment(int64_t channels, int64_t baselines)
{
    // Aim for >= 4 waves of 148 SMs x 1024 resident threads, but keep the 12-sample halo
    // per segment below ~5 %: segments of 256..1024 channels, multiples of 32.
    const int64_t want_threads = (int64_t) ksp_sm_count() * 1024 * 4;
    int64_t bl = ksp_divup(baselines, 32) * 32;
    int64_t nseg = ksp_divup(want_threads, bl);
    int64_t seg = channels / (nseg > 0 ? nseg : 1);
    seg = (seg / 32) * 32;
    if (seg < 256) seg = 256;
    if (seg > 1024) seg = 1024;
    return (int) seg;
}

template <bool TRANSPOSED>
int launch_bg(cudaStream_t s, const void *vis, float *out, const uint8_t *flags, int64_t channels,
              int64_t baselines, int64_t vis_stride, int64_t out_stride, int64_t flags_stride,
              int width, int is_amplitude, int flag_mode, int abs_mode)
{
    if (channels < 0 || baselines < 0) return KSP_EINVAL;
    if (width < 1 || !(width & 1)) return KSP_EINVAL;
    if (width > KSP_MAX_WIDTH) return KSP_ETOOLARGE;
    if (flag_mode < KSP_FLAGS_NONE || flag_mode > KSP_FLAGS_FULL) return KSP_EINVAL;
    if (abs_mode != KSP_ABS_NUMPY && abs_mode != KSP_ABS_HYPOT) return KSP_EINVAL;
    if (channels > 0x7fff0000) return KSP_ETOOLARGE;
    if (channels == 0 || baselines == 0) return 0;
    if (!vis || !out || (flag_mode != KSP_FLAGS_NONE && !flags)) return KSP_EINVAL;
    if (vis_stride < baselines) return KSP_EINVAL;
    if (TRANSPOSED ? out_stride < channels : out_stride < baselines) return KSP_EINVAL;
    if (flag_mode == KSP_FLAGS_FULL && flags_stride < baselines) return KSP_EINVAL;
    if ((uintptr_t) vis % (is_amplitude ? 4 : 8) || (uintptr_t) out % 4) return KSP_EALIGN;

    BgArgs a;
    a.vis = vis; a.out = out; a.flags = flags;
    a.channels = channels; a.baselines = baselines;
    a.vis_stride = vis_stride; a.out_stride = out_stride; a.flags_stride = flags_stride;
    a.seg = pick_segment(channels, baselines);
    dim3 grid((unsigned) ksp_divup(baselines, TILE_B), (unsigned) ksp_divup(channels, BG_TC));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    const int in_mode = is_amplitude ? IN_AMP : (abs_mode == KSP_ABS_NUMPY ? IN_NUMPY : IN_HYPOT);

    if (width != 13 && width <= 31) {
        switch (width) {
#define KSP_BGW(Wd) case Wd: return launch_bgw<Wd, TRANSPOSED>(s, a, in_mode, flag_mode);
        KSP_BGW(3) KSP_BGW(5) KSP_BGW(7) KSP_BGW(9) KSP_BGW(11) KSP_BGW(15) KSP_BGW(17)
        KSP_BGW(19) KSP_BGW(21) KSP_BGW(23) KSP_BGW(25) KSP_BGW(27) KSP_BGW(29) KSP_BGW(31)
#undef KSP_BGW
        default: break;
        }
    }
    if (width != 13) {
        dim3 ggrid((unsigned) ksp_divup(baselines, BG_THREADS), (unsigned) ksp_divup(channels, a.seg));
        if (ggrid.y > 65535) return KSP_ETOOLARGE;
        bg_generic_kernel<TRANSPOSED><<<ggrid, BG_THREADS, 0, s>>>(a, width, in_mode, flag_mode);
        KSP_CHECK_LAUNCH();
        return 0;
    }
#define KSP_BG_CASE(IM, FM)                                                        \
    if (in_mode == IM && flag_mode == FM) {                                        \
        if (TileGeom<BG_TC>::SMEM_BYTES > 48 * 1024)                               \
            KSP_CUDA(cudaFuncSetAttribute(bg13_kernel<IM, FM, TRANSPOSED, BG_TC>,  \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          TileGeom<BG_TC>::SMEM_BYTES));           \
        bg13_kernel<IM, FM, TRANSPOSED, BG_TC>                                     \
            <<<grid, BG_THREADS, TileGeom<BG_TC>::SMEM_BYTES, s>>>(a);             \
        KSP_CHECK_LAUNCH();                                                        \
        return 0;                                                                  \
    }
    KSP_BG_CASE(IN_NUMPY, KSP_FLAGS_NONE)
    KSP_BG_CASE(IN_NUMPY, KSP_FLAGS_CHANNEL)
    KSP_BG_CASE(IN_NUMPY, KSP_FLAGS_FULL)
    KSP_BG_CASE(IN_HYPOT, KSP_FLAGS_NONE)
    KSP_BG_CASE(IN_HYPOT, KSP_FLAGS_CHANNEL)
    KSP_BG_CASE(IN_HYPOT, KSP_FLAGS_FULL)
    KSP_BG_CASE(IN_AMP, KSP_FLAGS_NONE)
    KSP_BG_CASE(IN_AMP, KSP_FLAGS_CHANNEL)
    KSP_BG_CASE(IN_AMP, KSP_FLAGS_FULL)
#undef KSP_BG_CASE
    return KSP_EINVAL;
}

}  // namespace

extern "C" int ksp_background_median_filter(void *stream, const void *vis, float *dev,
                                            const uint8_t *flags, int64_t channels,
                                            int64_t baselines, int64_t vis_stride,
                                            int64_t dev_stride, int64_t flags_stride, int width,
                                            int is_amplitude, int flag_mode, int abs_mode)
{
    return launch_bg<false>((cudaStream_t) stream, vis, dev, flags, channels, baselines, vis_stride,
                            dev_stride, flags_stride, width, is_amplitude, flag_mode, abs_mode);
}

extern "C" int ksp_background_median_filter_t(void *stream, const void *vis, float *dev_t,
                                              const uint8_t *flags, int64_t channels,
                                              int64_t baselines, int64_t vis_stride,
                                              int64_t dev_t_stride, int64_t flags_stride,
                                              int width, int is_amplitude, int flag_mode,
                                              int abs_mode)
{
    return launch_bg<true>((cudaStream_t) stream, vis, dev_t, flags, channels, baselines,
                           vis_stride, dev_t_stride, flags_stride, width, is_amplitude, flag_mode,
                           abs_mode);
}
