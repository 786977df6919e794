// 2-D SumThreshold flagger (time x frequency per baseline).
//
// Replaces reference rfi/twodflag.py:67-890 (numba, CPU only): _average_freq, _time_median,
// _get_background2d (masked box-Gaussian filters, per-chunk MAD rejection, NaN interpolation),
// _sum_threshold along both axes, _combine_flags, _unaverage_freq, _get_flags_impl.
//
// The results of that code depend on the ORDER of its float64 running sums (box filters,
// cumulative sums of the SumThreshold) and on which intermediate is float32, so this file keeps
// every such recurrence serial along its axis, in the reference's order and precision, and takes
// its parallelism from everything the reference loops over independently: baselines (one
// 256-thread block per baseline, blocks persistent over the batch), and inside a baseline the
// columns of a time-axis recurrence, the rows (x frequency chunks) of a frequency-axis one.
// oracle/twodflag_numpy.py states the same arithmetic in numpy and is pinned bit for bit against
// the reference; the tests compare this file's flags with both.
//
// Three launches per batch of baselines:
//   twod_average_kernel   (time, freq, baseline) input -> baseline-major averaged magnitudes + flags
//   twod_baseline_kernel  everything of rfi/twodflag.py:768-881 for one baseline per block
//   twod_output_kernel    baseline-major flags -> (time, freq, baseline), OR isnan(input)
#include "common.cuh"
#include "select.cuh"
#include <math.h>
#include <string.h>

namespace {

using namespace ksp;

constexpr int TD_THREADS = 256;
constexpr int TD_PASSES = 4;                 // box filters per Gaussian (reference default, twodflag.py:313)
constexpr double TD_MAD_NORMAL = 1.4826;     // rfi/__init__.py:31

struct TdArgs {
    ksp_twodflag_params p;
    const void *data;            // (time, freq, baseline)
    const uint8_t *in_flags;     // same shape, non-zero = flagged
    uint8_t *out_flags;          // same shape
    int64_t bl0, nb;             // batch of baselines
    int a_freq;                  // averaged channels
    int max_rt, max_rf;          // largest box radii (they size the padded work array)
    int max_wf;                  // largest frequency window
    char *scratch;
    size_t per_bl;               // scratch bytes per baseline
};

// ---- per-baseline scratch layout (A = n_time * a_freq elements)
struct TdBuffers {
    float *data, *bg, *weight, *pad, *vals;
    int64_t pad_n;               // floats per work area of `pad` (there are two)
    float *spec_data, *spec_bg, *spec_weight;
    uint8_t *flags, *work, *tfl, *ffl, *pos, *neg, *hp, *hn, *spec_flags, *spec_work, *spec_out, *comb;
    uint8_t *outb;               // (n_time, n_freq) flags of this baseline at the original resolution
    uint8_t *row_full, *col_full; // n_time, n_freq
    float *thr;                  // thresholds per (row, chunk)
};

__host__ __device__ inline size_t td_align(size_t x) { return (x + 255) / 256 * 256; }

__host__ __device__ inline size_t td_row_work(const ksp_twodflag_params &p, int a_freq, int max_wf);

__host__ __device__ inline size_t td_layout(const ksp_twodflag_params &p, int a_freq, int max_r_t, int max_r_f,
                                            int max_wf, char *base, TdBuffers *b)
{
    const size_t T = (size_t) p.n_time, F = (size_t) a_freq, A = T * F;
    const size_t pad_t = (T + (size_t) max_r_t * TD_PASSES) * F, pad_f = T * (F + (size_t) max_r_f * TD_PASSES);
    const size_t pad_n = pad_t > pad_f ? pad_t : pad_f;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += td_align(bytes); return base ? base + o : (char *) nullptr; };
    float *f_data = (float *) take(A * 4), *f_bg = (float *) take(A * 4), *f_w = (float *) take(A * 4);
    float *f_pad = (float *) take(2 * pad_n * 4), *f_vals = (float *) take(A * 4);
    float *s_data = (float *) take(F * 4), *s_bg = (float *) take(F * 4), *s_w = (float *) take(F * 4);
    const size_t W = T * td_row_work(p, a_freq, max_wf);       // >= A
    uint8_t *u[12];
    for (int i = 0; i < 4; i++) u[i] = (uint8_t *) take(A);
    for (int i = 4; i < 8; i++) u[i] = (uint8_t *) take(W);    // pos, neg, hits
    for (int i = 8; i < 11; i++) u[i] = (uint8_t *) take(F);
    u[11] = (uint8_t *) take(A);
    uint8_t *outb = (uint8_t *) take(T * (size_t) p.n_freq);
    uint8_t *row_full = (uint8_t *) take(T), *col_full = (uint8_t *) take((size_t) p.n_freq);
    float *thr = (float *) take((T + 1) * (size_t) (p.n_chunks > 0 ? p.n_chunks : 1) * 4);
    if (b) {
        b->data = f_data; b->bg = f_bg; b->weight = f_w; b->pad = f_pad; b->vals = f_vals;
        b->pad_n = (int64_t) pad_n;
        b->spec_data = s_data; b->spec_bg = s_bg; b->spec_weight = s_w;
        b->flags = u[0]; b->work = u[1]; b->tfl = u[2]; b->ffl = u[3]; b->pos = u[4]; b->neg = u[5];
        b->hp = u[6]; b->hn = u[7]; b->spec_flags = u[8]; b->spec_work = u[9]; b->spec_out = u[10];
        b->comb = u[11]; b->outb = outb; b->thr = thr; b->row_full = row_full; b->col_full = col_full;
    }
    return off;
}

__host__ __device__ inline size_t td_row_work(const ksp_twodflag_params &p, int a_freq, int max_wf)
{
    // work bytes of one row of a frequency-axis SumThreshold: every (row, chunk) has its own
    // padded slice (chunk + 2 (largest window - 1))
    return (size_t) a_freq + (size_t) p.n_chunks * 2 * (size_t) max_wf;
}

__device__ __forceinline__ void td_buffers(const TdArgs &a, int64_t blr, TdBuffers *b)
{
    td_layout(a.p, a.a_freq, a.max_rt, a.max_rf, a.max_wf, a.scratch + (size_t) blr * a.per_bl, b);
}

// ------------------------------------------------------------------ averaging (twodflag.py:68-116)
template <bool COMPLEX>
__global__ void __launch_bounds__(256)
twod_average_kernel(const TdArgs a)
{
    // thread <-> (baseline of the batch, time, averaged channel); baseline fastest so that the
    // reads of a warp are contiguous
    const int64_t total = a.nb * a.p.n_time * a.a_freq;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t blr = i % a.nb;
    const int64_t rest = i / a.nb;
    const int jo = (int) (rest % a.a_freq);
    const int t = (int) (rest / a.a_freq);
    const int factor = a.p.average_freq;
    float sum = 0.0f;
    int count = 0;
    for (int j = jo * factor; j < (jo + 1) * factor && j < a.p.n_freq; j++) {
        const int64_t idx = ((int64_t) t * a.p.n_freq + j) * a.p.n_bl + a.bl0 + blr;
        float mag;
        if (COMPLEX) {
            const float2 v = reinterpret_cast<const float2 *>(a.data)[idx];
            // numba's abs(complex64): the correctly rounded hypot
            mag = (isnan(v.x) || isnan(v.y)) ? __int_as_float(0x7fc00000)
                                             : abs_slow(fabsf(v.x), fabsf(v.y), KSP_ABS_HYPOT);
        } else {
            mag = fabsf(reinterpret_cast<const float *>(a.data)[idx]);
        }
        if (!a.in_flags[idx] && !isnan(mag)) {
            sum = __fadd_rn(sum, mag);
            count++;
        }
    }
    TdBuffers b;
    td_buffers(a, blr, &b);
    const size_t o = (size_t) t * a.a_freq + jo;
    b.data[o] = count ? __fdiv_rn(sum, (float) count) : 0.0f;
    b.flags[o] = count == 0;
}

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ float median_sorted_f32(const float *s, int n)
{
    if (n & 1) return s[n / 2];
    return __fmul_rn(__fadd_rn(s[n / 2 - 1], s[n / 2]), 0.5f);
}

// insertion sort of a thread-private segment (n is a few hundred at most: the time axis)
__device__ void sort_small(float *v, int n)
{
    for (int i = 1; i < n; i++) {
        const float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) {
            v[j + 1] = v[j];
            j--;
        }
        v[j + 1] = x;
    }
}

// Median of |x| over the unflagged elements of an index set, by the whole block (exact radix
// select on the float bit patterns).  key_at(i) returns KEY_SKIP for flagged elements.  NaN if none.
template <typename KeyAt>
__device__ float block_median_abs(const KeyAt &key_at, int n, const SelectScratch &sc, uint32_t *s_count)
{
    const int tid = threadIdx.x;
    if (tid == 0) *s_count = 0u;
    __syncthreads();
    uint32_t mine = 0;
    for (int i = tid; i < n; i += TD_THREADS) mine += key_at(i) != KEY_SKIP;
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((tid & 31) == 0 && mine) atomicAdd(s_count, mine);
    __syncthreads();
    const uint32_t n_valid = *s_count;
    __syncthreads();
    if (n_valid == 0) return __int_as_float(0x7fc00000);
    const uint32_t k = (n_valid - 1) >> 1;
    const uint32_t lo = block_radix_select<TD_THREADS>(key_at, n, k, sc);
    uint32_t hi = lo;
    if (!(n_valid & 1u)) {
        uint32_t next, count_le;
        block_next_above<TD_THREADS>(key_at, n, lo, next, count_le, sc);
        if (count_le < k + 2) hi = next;
    }
    const float a = __uint_as_float(lo), b = __uint_as_float(hi);
    return lo == hi ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
}

// ------------------------------------------------------------------ box filter (twodflag.py:255-309)
// One line: n samples at data[k * stride], work array padded[(k) * pstride] of n + r * K floats.
// float64 running sum, float32 stores, in the reference's order; result / float32(d^K).
__device__ void box_line(const float *data, int64_t stride, float *padded, int64_t pstride, int n, int r,
                         float *out, int64_t ostride, float divisor)
{
    const int K = TD_PASSES;
    const int padding = r * K, len = n + padding;
    for (int i = 0; i < padding; i++) padded[(int64_t) i * pstride] = 0.0f;
    for (int i = 0; i < n; i++) padded[(int64_t) (padding + i) * pstride] = data[(int64_t) i * stride];
    int prev_start = padding;
    for (int p = 1; p <= K; p++) {
        double s = 0.0;
        int start = padding - 2 * r * p;
        int stop = start + n + 2 * padding;
        start = max(start, 0);
        stop = min(stop, len);
        const int tail = min(stop, len - 2 * r);
        for (int i = prev_start; i < min(start + 2 * r, len); i++) s += (double) padded[(int64_t) i * pstride];
        // Step i reads padded[i + 2 r] and padded[i] and writes padded[i]: every read is of an
        // element no earlier step wrote, so the loads of a batch of steps are issued together and
        // only the two float64 additions per step stay serial.
        int i = start;
        for (; i + 8 <= tail; i += 8) {
            float in[8], prev[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                in[k] = padded[(int64_t) (i + k + 2 * r) * pstride];
                prev[k] = padded[(int64_t) (i + k) * pstride];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                s += (double) in[k];
                padded[(int64_t) (i + k) * pstride] = (float) s;
                s -= (double) prev[k];
            }
        }
        for (; i < tail; i++) {
            s += (double) padded[(int64_t) (i + 2 * r) * pstride];
            const float prev = padded[(int64_t) i * pstride];
            padded[(int64_t) i * pstride] = (float) s;
            s -= (double) prev;
        }
        for (int i = tail; i < stop; i++) {
            const float prev = padded[(int64_t) i * pstride];
            padded[(int64_t) i * pstride] = (float) s;
            s -= (double) prev;
        }
        prev_start = start;
    }
    for (int i = 0; i < n; i++) out[(int64_t) i * ostride] = __fdiv_rn(padded[(int64_t) i * pstride], divisor);
}

// float32(2 r + 1) ** K as numba evaluates it: binary exponentiation in float32 (K = 4: the
// square of the square)
__device__ __forceinline__ float box_divisor(int r)
{
    static_assert(TD_PASSES == 4, "square of the square");
    const float d = (float) (2 * r + 1);
    const float d2 = __fmul_rn(d, d);
    return __fmul_rn(d2, d2);
}

// In-place 2-D filter of TWO arrays (T x F each; they are independent, so their lines run side
// by side): time axis (one thread per array and column), then frequency (per array and row).
// pad: two work areas of pad_n floats.
__device__ void box_filter_2d_pair(float *arr0, float *arr1, int T, int F, int r_t, int r_f, float *pad,
                                   int64_t pad_n)
{
    const int tid = threadIdx.x;
    if (r_t > 0) {
        const float div = box_divisor(r_t);
        for (int i = tid; i < 2 * F; i += TD_THREADS) {
            const int which = i >= F, f = which ? i - F : i;
            float *arr = which ? arr1 : arr0;
            box_line(arr + f, F, pad + which * pad_n + f, F, T, r_t, arr + f, F, div);
        }
        __syncthreads();
    }
    if (r_f > 0) {
        const float div = box_divisor(r_f);
        const int64_t plen = F + (int64_t) r_f * TD_PASSES;
        for (int i = tid; i < 2 * T; i += TD_THREADS) {
            const int which = i >= T, t = which ? i - T : i;
            float *arr = which ? arr1 : arr0;
            box_line(arr + (int64_t) t * F, 1, pad + which * pad_n + t * plen, 1, F, r_f, arr + (int64_t) t * F, 1, div);
        }
        __syncthreads();
    }
}

// twodflag.py:360-400
__device__ void masked_gaussian(const float *data, const uint8_t *flags, int T, int F, int r_t, int r_f,
                                float *out, float *weight, float *pad, int64_t pad_n)
{
    const int tid = threadIdx.x, A = T * F;
    for (int i = tid; i < A; i += TD_THREADS) {
        weight[i] = flags[i] ? 0.0f : 1.0f;
        out[i] = flags[i] ? 0.0f : data[i];
    }
    __syncthreads();
    box_filter_2d_pair(weight, out, T, F, r_t, r_f, pad, pad_n);
    for (int i = tid; i < A; i += TD_THREADS)
        out[i] = (weight[i] == 0.0f) ? __int_as_float(0x7fc00000) : __fdiv_rn(out[i], weight[i]);
    __syncthreads();
}

// twodflag.py:200-251, one row
__device__ void interpolate_row(float *row, int n)
{
    int p = 0;
    while (p < n && isnan(row[p])) p++;
    if (p == n) {
        for (int i = 0; i < n; i++) row[i] = 0.0f;
        return;
    }
    for (int i = 0; i < p; i++) row[i] = row[p];
    p++;
    while (p < n) {
        if (isnan(row[p])) {
            int q = p + 1;
            while (q < n && isnan(row[q])) q++;
            if (q == n) {
                for (int i = p; i < n; i++) row[i] = row[p - 1];
            } else {
                const float start = row[p - 1];
                const double grad = (double) __fsub_rn(row[q], start) / (double) (q - (p - 1));
                for (int i = p; i < q; i++)
                    row[i] = (float) __dadd_rn((double) start, __dmul_rn((double) (i - (p - 1)), grad));
            }
            p = q;
        } else {
            p++;
        }
    }
}

// twodflag.py:404-463.  flags_in is not modified; `work` receives the growing mask.
__device__ void background2d(const TdArgs &a, const float *data, const uint8_t *flags_in, int T, int F,
                             const int *r_t, const int *r_f, float *bg, uint8_t *work, float *weight,
                             float *pad, int64_t pad_n, const SelectScratch &sc, uint32_t *s_count)
{
    const int tid = threadIdx.x, A = T * F;
    for (int i = tid; i < A; i += TD_THREADS) work[i] = flags_in[i] != 0;
    __syncthreads();
    for (int ef = a.p.background_iterations; ef >= 1; ef--) {
        masked_gaussian(data, work, T, F, r_t ? r_t[ef] : 0, r_f[ef], bg, weight, pad, pad_n);
        for (int c = 0; c < a.p.n_chunks; c++) {
            const int c0 = (int) a.p.chunk_ends[c], c1 = (int) a.p.chunk_ends[c + 1], cl = c1 - c0;
            for (int i = tid; i < T * cl; i += TD_THREADS) {
                const int o = (i / cl) * F + c0 + i % cl;
                bg[o] = fabsf(__fsub_rn(data[o], bg[o]));
            }
            __syncthreads();
            auto key_at = [=](int i) -> uint32_t {
                const int o = (i / cl) * F + c0 + i % cl;
                return work[o] ? KEY_SKIP : __float_as_uint(bg[o]);       // residuals are >= 0
            };
            const float med = block_median_abs(key_at, T * cl, sc, s_count);
            const double threshold = __dmul_rn((double) med, __dmul_rn(TD_MAD_NORMAL, a.p.background_reject));
            for (int i = tid; i < T * cl; i += TD_THREADS) {
                const int o = (i / cl) * F + c0 + i % cl;
                if ((double) bg[o] > threshold) work[o] = 1;
            }
            __syncthreads();
        }
    }
    masked_gaussian(data, work, T, F, r_t ? r_t[1] : 0, r_f[1], bg, weight, pad, pad_n);
    for (int t = tid; t < T; t += TD_THREADS) interpolate_row(bg + (int64_t) t * F, F);
    __syncthreads();
}

// ------------------------------------------------------------------ SumThreshold, one line (twodflag.py:493-560)
// A chunk of `clen` samples starting `coff` samples into its padded slice of `len` samples
// (the chunk extended by the largest window - 1 on both sides, clipped to the array):
// line[i * stride], i < len.  thr32 = the chunk's threshold (already scaled; inf without data).
// pos / neg / hp / hn: work bytes, [i * wstride].  Writes out[k * ostride], k < clen.
__device__ void sum_threshold_line(const float *line, int64_t stride, int len, int coff, int clen, float thr32,
                                   const int *windows, int n_windows, const double *tf, uint8_t *pos,
                                   uint8_t *neg, uint8_t *hp, uint8_t *hn, int64_t wstride, uint8_t *out,
                                   int64_t ostride)
{
    for (int i = 0; i < len; i++) pos[i * wstride] = neg[i * wstride] = 0;
    for (int wi = 0; wi < n_windows; wi++) {
        const int w = windows[wi];
        const float lim = (float) ((double) thr32 / tf[wi]);
        const double scale = (double) (float) (1.0 / (double) w);       // np.float32(1.0 / window)
        const double dlim = (double) lim;
        double ring[KSP_TWOD_MAX_WINDOW + 1];                           // cum[j] at j mod (w + 1)
        double cum = 0.0;
        ring[0] = 0.0;
        for (int i = 0; i < len; i++) {
            float x = line[i * stride];
            if (pos[i * wstride] && x > lim) x = lim;
            else if (neg[i * wstride] && x < -lim) x = -lim;
            cum += (double) x;
            ring[(i + 1) % (w + 1)] = cum;
            hp[i * wstride] = hn[i * wstride] = 0;
            if (i + 1 >= w) {
                const int k = i + 1 - w;                                 // the window k .. k + w - 1
                const double avg = cum - ring[k % (w + 1)];
                hp[k * wstride] = (avg * scale) > dlim;
                hn[k * wstride] = (avg * -scale) > dlim;
            }
        }
        // every hit flags the samples of its window
        int run_p = 0, run_n = 0;
        for (int i = 0; i < len; i++) {
            if (hp[i * wstride]) run_p = w;
            if (hn[i * wstride]) run_n = w;
            if (run_p > 0) { pos[i * wstride] = 1; run_p--; }
            if (run_n > 0) { neg[i * wstride] = 1; run_n--; }
        }
    }
    for (int k = 0; k < clen; k++) out[k * ostride] = pos[(coff + k) * wstride] | neg[(coff + k) * wstride];
}

__device__ __forceinline__ float scaled_threshold(float med, double outlier_nsigma)
{
    if (isnan(med)) return __int_as_float(0x7f800000);
    return (float) __dmul_rn((double) med, __dmul_rn(outlier_nsigma, TD_MAD_NORMAL));
}

// SumThreshold along frequency with per-chunk thresholds (axis 1): T rows x n_chunks chunks,
// one thread per (row, chunk), each with its own work slice.
__device__ void sum_threshold_freq(const TdArgs &a, const float *data, const uint8_t *flags, int T, int F,
                                   uint8_t *out, const TdBuffers &b, const SelectScratch &sc, uint32_t *s_count)
{
    const int tid = threadIdx.x;
    const int nc = a.p.n_chunks;
    const int max_w = a.max_wf;
    // thresholds: median of |data| over the unflagged samples of (row, chunk), the whole block at a time
    for (int t = 0; t < T; t++)
        for (int c = 0; c < nc; c++) {
            const int c0 = (int) a.p.chunk_ends[c], cl = (int) a.p.chunk_ends[c + 1] - c0;
            const float *row = data + (int64_t) t * F + c0;
            const uint8_t *frow = flags + (int64_t) t * F + c0;
            auto key_at = [=](int i) -> uint32_t {
                return frow[i] ? KEY_SKIP : (__float_as_uint(row[i]) & 0x7fffffffu);
            };
            const float med = block_median_abs(key_at, cl, sc, s_count);
            if (tid == 0) b.thr[t * nc + c] = scaled_threshold(med, a.p.outlier_nsigma);
        }
    __syncthreads();
    const int64_t row_work = (int64_t) td_row_work(a.p, F, max_w);
    for (int i = tid; i < T * nc; i += TD_THREADS) {
        const int t = i / nc, c = i % nc;
        const int c0 = (int) a.p.chunk_ends[c], c1 = (int) a.p.chunk_ends[c + 1];
        if (c1 <= c0) continue;
        const int p0 = max(c0 - max_w + 1, 0), p1 = min(c1 + max_w - 1, F);
        const int64_t wo = (int64_t) t * row_work + c0 + (int64_t) c * 2 * max_w;   // slices never overlap
        sum_threshold_line(data + (int64_t) t * F + p0, 1, p1 - p0, c0 - p0, c1 - c0, b.thr[t * nc + c],
                           a.p.windows_freq, a.p.n_windows_freq, a.p.tf_freq, b.pos + wo, b.neg + wo,
                           b.hp + wo, b.hn + wo, 1, out + (int64_t) t * F + c0, 1);
    }
    __syncthreads();
}

// ------------------------------------------------------------------ one baseline (twodflag.py:768-881)
__global__ void __launch_bounds__(TD_THREADS, 4)
twod_baseline_kernel(const TdArgs a)
{
    __shared__ uint32_t s_hist[SELECT_HIST_WORDS];
    __shared__ uint32_t s_misc[64];
    __shared__ uint32_t s_count;
    SelectScratch sc;
    sc.hist = s_hist;
    sc.misc = s_misc;
    const int tid = threadIdx.x;
    const int T = (int) a.p.n_time, F = a.a_freq, A = T * F;
    for (int64_t blr = blockIdx.x; blr < a.nb; blr += gridDim.x) {
        TdBuffers b;
        td_buffers(a, blr, &b);

        // ---- median spectrum over time (twodflag.py:120-158)
        for (int f = tid; f < F; f += TD_THREADS) {
            float *v = b.vals + (int64_t) f * T;
            int n = 0;
            for (int t = 0; t < T; t++)
                if (!b.flags[(int64_t) t * F + f]) v[n++] = b.data[(int64_t) t * F + f];
            if (n == 0) {
                b.spec_data[f] = 0.0f;
                b.spec_flags[f] = 1;
            } else {
                sort_small(v, n);
                b.spec_data[f] = median_sorted_f32(v, n);
                b.spec_flags[f] = 0;
            }
        }
        __syncthreads();
        // ---- background and SumThreshold of the spectrum
        background2d(a, b.spec_data, b.spec_flags, 1, F, nullptr, a.p.r_freq, b.spec_bg, b.spec_work,
                     b.spec_weight, b.pad, b.pad_n, sc, &s_count);
        for (int f = tid; f < F; f += TD_THREADS) b.spec_data[f] = __fsub_rn(b.spec_data[f], b.spec_bg[f]);
        __syncthreads();
        sum_threshold_freq(a, b.spec_data, b.spec_flags, 1, F, b.spec_out, b, sc, &s_count);
        for (int i = tid; i < A; i += TD_THREADS) b.flags[i] |= b.spec_out[i % F];
        __syncthreads();
        // ---- 2-D background
        background2d(a, b.data, b.flags, T, F, a.p.r_time, a.p.r_freq, b.bg, b.work, b.weight, b.pad, b.pad_n, sc, &s_count);
        for (int i = tid; i < A; i += TD_THREADS) b.data[i] = __fsub_rn(b.data[i], b.bg[i]);
        __syncthreads();
        // ---- SumThreshold along time: one column per thread, one chunk [0, T)
        for (int f = tid; f < F; f += TD_THREADS) {
            float *v = b.vals + (int64_t) f * T;
            int n = 0;
            for (int t = 0; t < T; t++)
                if (!b.flags[(int64_t) t * F + f]) v[n++] = fabsf(b.data[(int64_t) t * F + f]);
            float med = __int_as_float(0x7fc00000);
            if (n > 0) {
                sort_small(v, n);
                med = median_sorted_f32(v, n);
            }
            sum_threshold_line(b.data + f, F, T, 0, T, scaled_threshold(med, a.p.outlier_nsigma),
                               a.p.windows_time, a.p.n_windows_time, a.p.tf_time, b.pos + f, b.neg + f,
                               b.hp + f, b.hn + f, F, b.tfl + f, F);
        }
        __syncthreads();
        for (int i = tid; i < A; i += TD_THREADS) b.flags[i] |= b.tfl[i];
        __syncthreads();
        // ---- SumThreshold along frequency
        sum_threshold_freq(a, b.data, b.flags, T, F, b.ffl, b, sc, &s_count);
        // ---- combine and smear in time (twodflag.py:691-722)
        {
            const int lo = -(a.p.time_extend / 2), hi = lo + a.p.time_extend;
            for (int f = tid; f < F; f += TD_THREADS) {
                // any flag in rows [t + lo, t + hi) clipped to the array
                for (int t = 0; t < T; t++) {
                    const int t0 = max(t + lo, 0), t1 = min(t + hi, T);
                    uint8_t any = 0;
                    for (int k = t0; k < t1; k++)
                        any |= b.spec_out[f] | b.tfl[(int64_t) k * F + f] | b.ffl[(int64_t) k * F + f];
                    b.comb[(int64_t) t * F + f] = any;
                }
            }
        }
        __syncthreads();
        // ---- back to the original channels, smear in frequency, fill rows / columns (twodflag.py:726-764)
        {
            const int OF = (int) a.p.n_freq, avg = a.p.average_freq;
            const int lo = -(a.p.freq_extend / 2), hi = lo + a.p.freq_extend;
            for (int i = tid; i < T * OF; i += TD_THREADS) {
                const int t = i / OF, f = i % OF;
                const int f0 = max(f + lo, 0), f1 = min(f + hi, OF);
                uint8_t any = 0;
                for (int k = f0; k < f1; k++) any |= b.comb[(int64_t) t * F + k / avg];
                b.outb[i] = any;
            }
            __syncthreads();
            // rows with too many flags (counted before any filling), then columns likewise
            uint8_t *row_full = b.row_full, *col_full = b.col_full;
            for (int t = tid; t < T; t += TD_THREADS) {
                int tot = 0;
                for (int f = 0; f < OF; f++) tot += b.outb[(int64_t) t * OF + f];
                row_full[t] = (double) tot > a.p.flag_all_freq_frac * (double) OF;
            }
            for (int f = tid; f < OF; f += TD_THREADS) {
                int tot = 0;
                for (int t = 0; t < T; t++) tot += b.outb[(int64_t) t * OF + f];
                col_full[f] = (double) tot > (double) T * a.p.flag_all_time_frac;
            }
            __syncthreads();
            for (int i = tid; i < T * OF; i += TD_THREADS)
                if (row_full[i / OF] || col_full[i % OF]) b.outb[i] = 1;
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ output (twodflag.py:680-688)
template <bool COMPLEX>
__global__ void __launch_bounds__(256)
twod_output_kernel(const TdArgs a)
{
    const int64_t total = a.nb * a.p.n_time * a.p.n_freq;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t blr = i % a.nb, rest = i / a.nb;             // rest = t * n_freq + f
    TdBuffers b;
    td_buffers(a, blr, &b);
    const int64_t idx = rest * a.p.n_bl + a.bl0 + blr;
    bool nan_in;
    if (COMPLEX) {
        const float2 v = reinterpret_cast<const float2 *>(a.data)[idx];
        nan_in = isnan(v.x) || isnan(v.y);
    } else {
        nan_in = isnan(reinterpret_cast<const float *>(a.data)[idx]);
    }
    a.out_flags[idx] = (b.outb[rest] || nan_in) ? 1 : 0;
}

void fill_derived(TdArgs &a)
{
    a.a_freq = (int) ((a.p.n_freq + a.p.average_freq - 1) / a.p.average_freq);
    a.max_rt = a.max_rf = 0;
    for (int e = 1; e <= a.p.background_iterations; e++) {
        if (a.p.r_time[e] > a.max_rt) a.max_rt = a.p.r_time[e];
        if (a.p.r_freq[e] > a.max_rf) a.max_rf = a.p.r_freq[e];
    }
    a.max_wf = 1;
    for (int i = 0; i < a.p.n_windows_freq; i++)
        if (a.p.windows_freq[i] > a.max_wf) a.max_wf = a.p.windows_freq[i];
    a.per_bl = td_layout(a.p, a.a_freq, a.max_rt, a.max_rf, a.max_wf, nullptr, nullptr);
}

int check_params(const ksp_twodflag_params *p)
{
    if (!p) return KSP_EINVAL;
    if (p->n_time < 1 || p->n_freq < 1 || p->n_bl < 0) return KSP_EINVAL;
    if (p->n_time > 4096 || p->n_freq > (1 << 22)) return KSP_ETOOLARGE;
    if (p->average_freq < 1 || p->background_iterations < 1 || p->background_iterations > 63) return KSP_EINVAL;
    if (p->n_windows_time < 0 || p->n_windows_time > KSP_TWOD_MAX_WINDOWS) return KSP_EINVAL;
    if (p->n_windows_freq < 0 || p->n_windows_freq > KSP_TWOD_MAX_WINDOWS) return KSP_EINVAL;
    for (int i = 0; i < p->n_windows_time; i++)
        if (p->windows_time[i] < 1 || p->windows_time[i] > KSP_TWOD_MAX_WINDOW) return KSP_ETOOLARGE;
    for (int i = 0; i < p->n_windows_freq; i++)
        if (p->windows_freq[i] < 1 || p->windows_freq[i] > KSP_TWOD_MAX_WINDOW) return KSP_ETOOLARGE;
    if (p->n_chunks < 1 || p->n_chunks > KSP_TWOD_MAX_CHUNKS) return KSP_EINVAL;
    const int64_t a_freq = (p->n_freq + p->average_freq - 1) / p->average_freq;
    if (p->chunk_ends[0] != 0 || p->chunk_ends[p->n_chunks] != a_freq) return KSP_EINVAL;
    for (int c = 0; c < p->n_chunks; c++)
        if (p->chunk_ends[c + 1] < p->chunk_ends[c]) return KSP_EINVAL;
    if (p->time_extend < 1 || p->freq_extend < 1) return KSP_EINVAL;
    for (int e = 1; e <= p->background_iterations; e++)
        if (p->r_time[e] < 0 || p->r_freq[e] < 0 || p->r_time[e] > 4096 || p->r_freq[e] > 4096) return KSP_EINVAL;
    if (p->n_time * a_freq > 0x3fffffff) return KSP_ETOOLARGE;
    return 0;
}

}  // namespace

extern "C" size_t ksp_twodflag_scratch_bytes(const ksp_twodflag_params *p, int64_t batch_baselines)
{
    if (check_params(p) || batch_baselines < 1) return 0;
    TdArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    fill_derived(a);
    return a.per_bl * (size_t) batch_baselines;
}

extern "C" int ksp_twodflag(void *stream, const ksp_twodflag_params *p, const void *data,
                            const uint8_t *in_flags, uint8_t *out_flags, void *scratch, size_t scratch_bytes,
                            int64_t batch_baselines)
{
    int rc = check_params(p);
    if (rc) return rc;
    if (p->n_bl == 0) return 0;
    if (!data || !in_flags || !out_flags || !scratch || batch_baselines < 1) return KSP_EINVAL;
    if ((uintptr_t) scratch % 256) return KSP_EALIGN;
    cudaStream_t s = (cudaStream_t) stream;
    TdArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.data = data; a.in_flags = in_flags; a.out_flags = out_flags;
    fill_derived(a);
    if (scratch_bytes < a.per_bl * (size_t) batch_baselines) return KSP_ESCRATCH;
    a.scratch = (char *) scratch;
    const int resident = 4 * ksp_sm_count();
    for (int64_t bl0 = 0; bl0 < p->n_bl; bl0 += batch_baselines) {
        a.bl0 = bl0;
        a.nb = p->n_bl - bl0 < batch_baselines ? p->n_bl - bl0 : batch_baselines;
        const int64_t n_avg = a.nb * p->n_time * a.a_freq, n_out = a.nb * p->n_time * p->n_freq;
        if (ksp_divup(n_avg, 256) > 0x7fffffff || ksp_divup(n_out, 256) > 0x7fffffff) return KSP_ETOOLARGE;
        if (p->is_complex) twod_average_kernel<true><<<(unsigned) ksp_divup(n_avg, 256), 256, 0, s>>>(a);
        else twod_average_kernel<false><<<(unsigned) ksp_divup(n_avg, 256), 256, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
        const int blocks = (int) (a.nb < resident ? a.nb : resident);
        twod_baseline_kernel<<<blocks, TD_THREADS, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
        if (p->is_complex) twod_output_kernel<true><<<(unsigned) ksp_divup(n_out, 256), 256, 0, s>>>(a);
        else twod_output_kernel<false><<<(unsigned) ksp_divup(n_out, 256), 256, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
    }
    return 0;
}
