/*
 * contract.c -- plain-C statement of the float32 DEVICE CONTRACT of the RFI
 * flagging hot path.  TEST INFRASTRUCTURE ONLY (checker for the CUDA kernels;
 * optionally the timed "port" CPU baseline in bench.py).  Nothing under
 * katsdpsigproc_b200/ links, loads or calls this file.
 *
 * The reference's CPU path (src/katsdpsigproc/rfi/host.py) works in float64
 * through numpy/pandas; its GPU path (rfi/ *.mako) works in float32 with
 * slightly different rounding.  The north star asks for results that equal the
 * HOST path, so the contract below is "float32 storage, host semantics":
 * every function cites the reference lines it follows and states where it
 * rounds.  Rules R1..R10 are those of SURVEY.md section 8(c').
 *
 * Parity status: PINNED -- tests/test_oracle_contract.py checks this file
 * against oracle/host_numpy.py (itself pinned to the reference's known-answer
 * tests and to reference-generated fixtures) on every golden case.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: a*b+c must not be fused behind our back.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_ABS_NUMPY 0 /* numpy's AVX-512F complex64 absolute value (R1) */
#define KO_ABS_HYPOT 1 /* correctly rounded hypot via float64 (non-AVX-512 numpy) */

#define KO_FLAGS_NONE 0
#define KO_FLAGS_CHANNEL 1
#define KO_FLAGS_FULL 2

/* ------------------------------------------------------------------ R1 --- */
/* |re + i*im| as numpy computes it for complex64 (reference host.py:137 calls
 * np.abs).  On AVX-512F hosts numpy evaluates, in float32 with one rounding per
 * step: L = max(|re|,|im|), r = min/L, L * sqrt(r*r + 1) with the multiply-add
 * fused; zero gives 0, an infinite part gives +inf, otherwise NaN propagates.
 * Verified against np.abs on 4M samples incl. denormals. */
float ko_abs(float re, float im, int mode)
{
    float x = fabsf(re), y = fabsf(im);
    if (isinf(x) || isinf(y))
        return INFINITY;
    if (mode == KO_ABS_HYPOT)
        return (float) sqrt((double) x * (double) x + (double) y * (double) y);
    float big = x > y ? x : y;
    float small = x > y ? y : x;
    if (isnan(x) || isnan(y))
        return NAN;
    if (big == 0.0f)
        return 0.0f;
    float r = small / big;
    float t = fmaf(r, r, 1.0f);
    return sqrtf(t) * big;
}

void ko_amplitude(const float *ri, float *out, long n, int mode)
{
    for (long i = 0; i < n; i++)
        out[i] = ko_abs(ri[2 * i], ri[2 * i + 1], mode);
}

/* ------------------------------------------------------------ R2, R3 ---- */
/* Sliding-median background (reference host.py:133-151; device twin
 * rfi/background_median_filter.mako:110-189).  Element (c, b) of vis lives at
 * vis[c*vis_stride + b] (complex64 = two floats, or one float when
 * is_amplitude).  A sample is usable when it is not flagged and its amplitude
 * is not NaN.  The window is [c-H, c+H] clipped to the band.  The median of an
 * even number of usable samples is the float64 mean of the two middle ones.
 * dev = float32( float64(amp) - median64 ); flagged / unusable centre -> 0. */
static void insert_sorted(float *buf, int *n, float v)
{
    int i = *n;
    while (i > 0 && buf[i - 1] > v) {
        buf[i] = buf[i - 1];
        i--;
    }
    buf[i] = v;
    (*n)++;
}

int ko_background(const float *vis, int is_amplitude, const uint8_t *flags, int flag_mode,
                  long flag_stride, float *dev, long channels, long baselines, long vis_stride,
                  long dev_stride, int width, int abs_mode)
{
    if (width < 1 || (width & 1) == 0 || width > 255)
        return -1;
    const int half = width / 2;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < baselines; b++) {
        float *amp = (float *) malloc(sizeof(float) * (size_t) channels);
        uint8_t *ok = (uint8_t *) malloc((size_t) channels);
        for (long c = 0; c < channels; c++) {
            float a;
            if (is_amplitude)
                a = vis[c * vis_stride + b];
            else
                a = ko_abs(vis[2 * (c * vis_stride + b)], vis[2 * (c * vis_stride + b) + 1],
                           abs_mode);
            int bad = 0;
            if (flag_mode == KO_FLAGS_CHANNEL)
                bad = flags[c] != 0;
            else if (flag_mode == KO_FLAGS_FULL)
                bad = flags[c * flag_stride + b] != 0;
            amp[c] = a;
            ok[c] = !bad && !isnan(a);
        }
        float window[256];
        for (long c = 0; c < channels; c++) {
            long lo = c - half < 0 ? 0 : c - half;
            long hi = c + half >= channels ? channels - 1 : c + half;
            int n = 0;
            for (long k = lo; k <= hi; k++)
                if (ok[k])
                    insert_sorted(window, &n, amp[k]);
            float out = 0.0f;
            if (ok[c] && n > 0) {
                double med;
                if (n & 1)
                    med = (double) window[n / 2];
                else
                    med = ((double) window[n / 2 - 1] + (double) window[n / 2]) * 0.5;
                out = (float) ((double) amp[c] - med);
            }
            dev[c * dev_stride + b] = out;
        }
        free(amp);
        free(ok);
    }
    return 0;
}

/* ---------------------------------------------------------------- R4 ---- */
/* noise[b] = float32( 1.4826 * median64{ |dev| : |dev| > 0 } ) (reference
 * host.py:157-163, rfi/__init__.py:31; device twin rfi/madnz_t.mako:72-87 and
 * rank.mako:236-266).  Element (c, b) is dev[c*stride_c + b*stride_b], so both
 * the channel-major and the baseline-major ("T") layouts are covered.  An even
 * count averages the two middle values in float64.  No usable sample -> NaN.
 * median_out (optional) receives float32(median64), the pure selection. */
static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *) a, y = *(const float *) b;
    return (x > y) - (x < y);
}

void ko_noise_mad(const float *dev, long channels, long baselines, long stride_c, long stride_b,
                  float *noise, float *median_out)
{
#pragma omp parallel for schedule(static)
    for (long b = 0; b < baselines; b++) {
        float *mag = (float *) malloc(sizeof(float) * (size_t) (channels > 0 ? channels : 1));
        long n = 0;
        for (long c = 0; c < channels; c++) {
            float a = fabsf(dev[c * stride_c + b * stride_b]);
            if (a > 0.0f)
                mag[n++] = a;
        }
        double med = NAN;
        if (n > 0) {
            qsort(mag, (size_t) n, sizeof(float), cmp_float);
            if (n & 1)
                med = (double) mag[n / 2];
            else
                med = ((double) mag[n / 2 - 1] + (double) mag[n / 2]) * 0.5;
        }
        noise[b] = (float) (1.4826 * med);
        if (median_out)
            median_out[b] = (float) med;
        free(mag);
    }
}

/* ---------------------------------------------------------- R5..R8 ------ */
/* Per-window thresholds (R5, reference host.py:215,235,252):
 *   thr_w = float32( (n_sigma * float64(noise)) * scales[w] ),  scales[w] = rho^-w
 * computed by the caller in Python exactly as the reference does. */
static float level_threshold(double n_sigma, float noise, double scale)
{
    return (float) ((n_sigma * (double) noise) * scale);
}

/* SumThreshold on one baseline (reference host.py:218-246; device twin
 * rfi/threshold_sum.mako:49-132).  x is contiguous scratch of length n.
 *
 * Contract for the window sums (R6): with F = samples flagged by earlier
 * window sizes, u[j] = F[j] ? 0 : x[j].  D_0 = u and D_{k+1}[i] =
 * fl32(D_k[i] + D_k[i + 2^k]) (a doubling tree in channel order, float32).
 * Window i of size 2^w fires iff
 *      float64(D_w[i]) > float64(thr_w) * (2^w - #F in window)
 * which is the reference's "sum with flagged samples replaced by thr_w exceeds
 * thr_w * window" with the replaced part moved to the right-hand side exactly.
 * Only windows fully inside the band exist (R7, np.convolve mode='valid'); a
 * window size larger than the band is skipped. */
static void sum_threshold_row(const float *x, long n, const float *thr, int n_windows,
                              uint8_t *flagged, float *tree, int32_t *count, uint8_t *fire)
{
    memset(flagged, 0, (size_t) n);
    for (int w = 0; w < n_windows; w++) {
        long win = 1L << w;
        if (win > n)
            break;
        for (long j = 0; j < n; j++) {
            tree[j] = flagged[j] ? 0.0f : x[j];
            count[j] = flagged[j];
        }
        long len = n;
        for (long step = 1; step < win; step <<= 1) {
            len -= step;
            for (long j = 0; j < len; j++) {
                tree[j] = tree[j] + tree[j + step];
                count[j] = count[j] + count[j + step];
            }
        }
        long n_pos = n - win + 1;
        for (long j = 0; j < n_pos; j++)
            fire[j] = (double) tree[j] > (double) thr[w] * (double) (win - count[j]);
        for (long j = 0; j < n_pos; j++)
            if (fire[j])
                for (long k = j; k < j + win; k++)
                    flagged[k] = 1;
    }
}

void ko_threshold_sum(const float *dev, const float *noise, uint8_t *flags, long channels,
                      long baselines, long stride_c, long stride_b, long fstride_c,
                      long fstride_b, int n_windows, double n_sigma, const double *scales,
                      int flag_value)
{
#pragma omp parallel for schedule(static)
    for (long b = 0; b < baselines; b++) {
        size_t n = (size_t) (channels > 0 ? channels : 1);
        float *x = (float *) malloc(sizeof(float) * n);
        float *tree = (float *) malloc(sizeof(float) * n);
        int32_t *count = (int32_t *) malloc(sizeof(int32_t) * n);
        uint8_t *flagged = (uint8_t *) malloc(n);
        uint8_t *fire = (uint8_t *) malloc(n);
        float thr[32];
        for (int w = 0; w < n_windows && w < 32; w++)
            thr[w] = level_threshold(n_sigma, noise[b], scales[w]);
        for (long c = 0; c < channels; c++)
            x[c] = dev[c * stride_c + b * stride_b];
        sum_threshold_row(x, channels, thr, n_windows, flagged, tree, count, fire);
        for (long c = 0; c < channels; c++)
            flags[c * fstride_c + b * fstride_b] = flagged[c] ? (uint8_t) flag_value : 0;
        free(x);
        free(tree);
        free(count);
        free(flagged);
        free(fire);
    }
}

/* ThresholdSimple (reference host.py:177-183): flag iff dev > thr_0. */
void ko_threshold_simple(const float *dev, const float *noise, uint8_t *flags, long channels,
                         long baselines, long stride_c, long stride_b, long fstride_c,
                         long fstride_b, double n_sigma, int flag_value)
{
#pragma omp parallel for schedule(static)
    for (long b = 0; b < baselines; b++) {
        float thr = level_threshold(n_sigma, noise[b], 1.0);
        for (long c = 0; c < channels; c++)
            flags[c * fstride_c + b * fstride_b] =
                dev[c * stride_c + b * stride_b] > thr ? (uint8_t) flag_value : 0;
    }
}

/* ---------------------------------------------------------------- R9 ---- */
/* Percentile5 (reference percentile.mako:115-140, oracle expression
 * test/test_percentile.py:79-84): per row of src, over columns
 * [first_col, first_col + n_cols): min, max and the order statistics of rank
 * (n-1)/4, 3(n-1)/4, (n-1)/2 of |value|; complex input takes the numpy
 * amplitude (R1) first, so the result is a pure selection of np.abs values. */
void ko_percentile5(const float *src, int is_amplitude, long rows, long stride, long first_col,
                    long n_cols, float *dest, long dest_stride, int abs_mode)
{
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) {
        float *v = (float *) malloc(sizeof(float) * (size_t) n_cols);
        for (long c = 0; c < n_cols; c++) {
            long idx = r * stride + first_col + c;
            v[c] = is_amplitude ? fabsf(src[idx]) : ko_abs(src[2 * idx], src[2 * idx + 1], abs_mode);
        }
        qsort(v, (size_t) n_cols, sizeof(float), cmp_float);
        dest[0 * dest_stride + r] = v[0];
        dest[1 * dest_stride + r] = v[n_cols - 1];
        dest[2 * dest_stride + r] = v[(n_cols - 1) / 4];
        dest[3 * dest_stride + r] = v[((n_cols - 1) * 3) / 4];
        dest[4 * dest_stride + r] = v[(n_cols - 1) / 2];
        free(v);
    }
}

/* --------------------------------------------------------------- R10 ---- */
/* MaskedSum (reference maskedsum.mako:38-68, oracle expression
 * test/test_maskedsum.py:62-67): dest[col] = sum over rows of
 * mask[row] * src[row, col] (complex), or of mask[row] * |src[row, col]|.
 * The reference device chains float32 fma in row order and numpy sums
 * pairwise; neither is bit-defined, so the contract is the float64-accumulated
 * sum rounded once to float32 (within 1e-6 relative of both). */
void ko_maskedsum(const float *src, const float *mask, long rows, long cols, long stride,
                  int use_amplitudes, float *dest, int abs_mode)
{
#pragma omp parallel for schedule(static)
    for (long c = 0; c < cols; c++) {
        double re = 0.0, im = 0.0;
        for (long r = 0; r < rows; r++) {
            double m = (double) mask[r];
            float x = src[2 * (r * stride + c)], y = src[2 * (r * stride + c) + 1];
            if (use_amplitudes)
                re += m * (double) ko_abs(x, y, abs_mode);
            else {
                re += m * (double) x;
                im += m * (double) y;
            }
        }
        if (use_amplitudes)
            dest[c] = (float) re;
        else {
            dest[2 * c] = (float) re;
            dest[2 * c + 1] = (float) im;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* The composed flagger (reference host.py:270-273; device rfi/device.py:
 * 1111-1166): background -> MAD noise -> SumThreshold on channel-major data.
 * dev (channels*baselines floats) and noise are caller-provided outputs. */
int ko_flagger(const float *vis, int is_amplitude, const uint8_t *in_flags, int flag_mode,
               float *dev, float *noise, uint8_t *flags, long channels, long baselines, int width,
               int n_windows, double n_sigma, const double *scales, int flag_value, int abs_mode)
{
    int rc = ko_background(vis, is_amplitude, in_flags, flag_mode, baselines, dev, channels,
                           baselines, baselines, baselines, width, abs_mode);
    if (rc)
        return rc;
    ko_noise_mad(dev, channels, baselines, baselines, 1, noise, NULL);
    if (n_windows <= 0)
        ko_threshold_simple(dev, noise, flags, channels, baselines, baselines, 1, baselines, 1,
                            n_sigma, flag_value);
    else
        ko_threshold_sum(dev, noise, flags, channels, baselines, baselines, 1, baselines, 1,
                         n_windows, n_sigma, scales, flag_value);
    return 0;
}
