// Channel-axis sliding-median background filter.
//
// Replaces reference rfi/background_median_filter.mako:200-220 (+ the float
// transpose of rfi/device.py:1152-1157 in the _t variant).  Semantics are those
// of the HOST class (rfi/host.py:133-151): windows clipped to
// the band, flagged / NaN samples skipped, even counts averaged in float64,
// flagged centre -> 0.
//
// Layout and mapping.  vis is channel-major, so one warp reads 32 neighbouring
// baselines of one channel as a single 256-byte request and every lane walks
// down the channel axis of its own baseline.  A lane keeps the last 16
// amplitudes in registers and emits 4 medians per step with a shared selection
// network (median13.cuh): the 10 samples common to the 4 windows are reduced
// to their 4 middle ranks once, each output then only merges its 3 private
// samples (23 min/max per output instead of ~70 for a sort).  Work is split
// along channels into segments (halo of 12 re-read per segment) so that the
// grid is several waves of 148 SMs.
//
// Slow path (warp-divergent, rare): any of the 16 samples is outside the band,
// flagged or NaN -> masked median by a small sort (median_masked13).
//
// The _t variant stages 32 baselines x 32 channels per warp in shared memory
// (pitch 36 floats: conflict-free 128-bit writes by lane and 128-bit reads by
// row) and stores baseline-major rows with 128-bit coalesced stores.
#include "common.cuh"
#include "median13.cuh"

namespace {

using namespace ksp;

constexpr int IN_NUMPY = 0;   // complex64, numpy AVX-512 amplitude rule
constexpr int IN_HYPOT = 1;   // complex64, correctly rounded hypot
constexpr int IN_AMP = 2;     // float32 amplitudes

constexpr int BG_THREADS = 128;
constexpr int TILE_PITCH = 36;

struct BgArgs {
    const void *vis;
    float *out;
    const uint8_t *flags;
    int64_t channels, baselines;
    int64_t vis_stride, out_stride, flags_stride;
    int seg;  // channels per segment (multiple of 32)
};

template <int IN_MODE>
__device__ __forceinline__ float load_amp(const void *vis, int64_t idx)
{
    if (IN_MODE == IN_AMP) {
        return ldg_stream_f(reinterpret_cast<const float *>(vis) + idx);
    } else {
        float2 v = ldg_stream_f2(reinterpret_cast<const float2 *>(vis) + idx);
        return abs_c64<IN_MODE == IN_NUMPY ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(v.x, v.y);
    }
}

// Masked medians for the 4 outputs of one step.  e is in logical order.
__device__ __noinline__ float4 slow_step(const float *e, unsigned bad)
{
    float r[4];
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        float out = 0.0f;
        if (!((bad >> (j + 6)) & 1u)) {
            float w[13];
#pragma unroll
            for (int k = 0; k < 13; k++) w[k] = e[j + k];
            float lo, hi;
            unsigned valid = ~(bad >> j) & 0x1fffu;
            if (median_masked13(w, valid, lo, hi)) {
                double med = (lo == hi) ? (double) lo : ((double) lo + (double) hi) * 0.5;
                out = __double2float_rn((double) e[j + 6] - med);
            }
        }
        r[j] = out;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

template <int IN_MODE, int FLAG_MODE, bool TRANSPOSED>
__global__ void __launch_bounds__(BG_THREADS, 6)
bg13_kernel(const BgArgs a)
{
    __shared__ __align__(16) float tile[TRANSPOSED ? (BG_THREADS / 32) * 32 * TILE_PITCH : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // a block is 4 warps = 4 channel segments of the same 32 baselines
    const int64_t b_raw = (int64_t) blockIdx.x * 32 + lane;
    const bool b_ok = b_raw < a.baselines;
    const int64_t b = b_ok ? b_raw : a.baselines - 1;
    const int64_t c_begin = ((int64_t) blockIdx.y * (BG_THREADS / 32) + warp) * a.seg;
    if (c_begin >= a.channels) return;   // whole warp; no block-level barriers in this kernel
    const int64_t c_end = min(a.channels, c_begin + (int64_t) a.seg);
    const int64_t C = a.channels;

    float e[16];
    unsigned bad = 0;  // logical slot k unusable (outside band / flagged / NaN)

    // fetch one sample: amplitude + "unusable" bit
    auto fetch = [&](int64_t c, float &amp) -> unsigned {
        if (c < 0 || c >= C) {
            amp = 0.0f;
            return 1u;
        }
        amp = load_amp<IN_MODE>(a.vis, c * a.vis_stride + b);
        unsigned u = (amp != amp) ? 1u : 0u;
        if (FLAG_MODE == KSP_FLAGS_CHANNEL) u |= (a.flags[c] != 0);
        if (FLAG_MODE == KSP_FLAGS_FULL) u |= (a.flags[c * a.flags_stride + b] != 0);
        return u;
    };

    // prologue: logical slots 0..11 <- channels c_begin-6 .. c_begin+5
#pragma unroll
    for (int k = 0; k < 12; k++) bad |= fetch(c_begin - 6 + k, e[k]) << k;

    float *my_tile = tile + warp * 32 * TILE_PITCH;

    for (int64_t c16 = c_begin; c16 < c_end; c16 += 16) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int rot = 4 * g;  // logical slot k lives in e[(k + rot) & 15]
            const int64_t c = c16 + 4 * g;
            if (c < c_end) {  // warp-uniform
                // new samples: logical slots 12..15 <- channels c+6 .. c+9
                unsigned nb = 0;
                if (c + 9 < C && FLAG_MODE == KSP_FLAGS_NONE) {
                    float s = 0.0f;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float v = load_amp<IN_MODE>(a.vis, (c + 6 + k) * a.vis_stride + b);
                        e[(12 + k + rot) & 15] = v;
                        s += v;
                    }
                    if (s != s) {  // some NaN (or inf - inf): look at each one
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float v = e[(12 + k + rot) & 15];
                            nb |= (v != v ? 1u : 0u) << k;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) nb |= fetch(c + 6 + k, e[(12 + k + rot) & 15]) << k;
                }
                bad |= nb << 12;

                float o0, o1, o2, o3;
                if (bad == 0) {
                    float m0, m1, m2, m3;
                    median13x4(e, rot, m0, m1, m2, m3);
                    o0 = e[(6 + rot) & 15] - m0;
                    o1 = e[(7 + rot) & 15] - m1;
                    o2 = e[(8 + rot) & 15] - m2;
                    o3 = e[(9 + rot) & 15] - m3;
                } else {
                    float lin[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) lin[k] = e[(k + rot) & 15];
                    float4 r = slow_step(lin, bad);
                    o0 = r.x; o1 = r.y; o2 = r.z; o3 = r.w;
                }
                bad >>= 4;

                if (TRANSPOSED) {
                    int col = (int) ((c - c_begin) & 31);
                    *reinterpret_cast<float4 *>(my_tile + lane * TILE_PITCH + col) =
                        make_float4(o0, o1, o2, o3);
                    if (col == 28 || c + 4 >= c_end) {
                        __syncwarp();
                        const int64_t tc0 = c - col;            // first channel of the tile
                        const int64_t tb0 = b_raw - lane;       // first baseline of the tile
                        const int ncols = (int) min((int64_t) 32, c_end - tc0);
                        const bool vec = (ncols == 32) && ((a.out_stride & 3) == 0) &&
                                         ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
                        if (vec) {
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                int row = (lane >> 3) + 4 * i;
                                int cc = 4 * (lane & 7);
                                float4 v = *reinterpret_cast<const float4 *>(
                                    my_tile + row * TILE_PITCH + cc);
                                if (tb0 + row < a.baselines)
                                    *reinterpret_cast<float4 *>(
                                        a.out + (tb0 + row) * a.out_stride + tc0 + cc) = v;
                            }
                        } else {
                            for (int row = 0; row < 32; row++) {
                                if (lane < ncols && tb0 + row < a.baselines)
                                    a.out[(tb0 + row) * a.out_stride + tc0 + lane] =
                                        my_tile[row * TILE_PITCH + lane];
                            }
                        }
                        __syncwarp();
                    }
                } else if (b_ok) {
                    float *o = a.out + c * a.out_stride + b;
                    o[0] = o0;
                    if (c + 1 < c_end) o[a.out_stride] = o1;
                    if (c + 2 < c_end) o[2 * a.out_stride] = o2;
                    if (c + 3 < c_end) o[3 * a.out_stride] = o3;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- any odd width
// One thread per (baseline, channel segment); window kept in local memory.
// Only used when width != 13.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(128)
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
    dim3 grid((unsigned) ksp_divup(baselines, 32),
              (unsigned) ksp_divup(ksp_divup(channels, a.seg), BG_THREADS / 32));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    const int in_mode = is_amplitude ? IN_AMP : (abs_mode == KSP_ABS_NUMPY ? IN_NUMPY : IN_HYPOT);

    if (width != 13) {
        dim3 ggrid((unsigned) ksp_divup(baselines, BG_THREADS), (unsigned) ksp_divup(channels, a.seg));
        if (ggrid.y > 65535) return KSP_ETOOLARGE;
        bg_generic_kernel<TRANSPOSED><<<ggrid, BG_THREADS, 0, s>>>(a, width, in_mode, flag_mode);
        KSP_CHECK_LAUNCH();
        return 0;
    }
#define KSP_BG_CASE(IM, FM)                                                        \
    if (in_mode == IM && flag_mode == FM) {                                        \
        bg13_kernel<IM, FM, TRANSPOSED><<<grid, BG_THREADS, 0, s>>>(a);            \
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
