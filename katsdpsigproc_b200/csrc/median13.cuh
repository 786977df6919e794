// Min/max programs for the sliding median of 13 (4 outputs per step) and the
// generic small-window median used at band edges and around flagged samples.
// __host__ __device__ so that tools/test_median13.cu can check them on the CPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define KSP_HD __host__ __device__ __forceinline__
#else
#define KSP_HD inline
#endif

namespace ksp {

#define KSP_CE(a, b)            \
    {                           \
        float lo_ = fminf(a, b); \
        float hi_ = fmaxf(a, b); \
        a = lo_;                \
        b = hi_;                \
    }

// Ranks 3..6 (0-based, ascending) of ten values.  A 29-comparator sorting
// network for 10 inputs pruned to the 52 min/max operations that reach wires
// 3..6 (tools/median_network.py derives and verifies it with the 0-1 principle).
KSP_HD void mid4_of_10(float v0, float v1, float v2, float v3, float v4, float v5, float v6,
                       float v7, float v8, float v9, float &m0, float &m1, float &m2, float &m3)
{
    KSP_CE(v0, v5) KSP_CE(v1, v6) KSP_CE(v2, v7) KSP_CE(v3, v8) KSP_CE(v4, v9)
    KSP_CE(v0, v3) KSP_CE(v1, v4) KSP_CE(v5, v8) KSP_CE(v6, v9)
    KSP_CE(v0, v2) KSP_CE(v3, v6) KSP_CE(v7, v9)
    v1 = fmaxf(v0, v1);
    KSP_CE(v2, v4) KSP_CE(v5, v7)
    v8 = fminf(v8, v9);
    KSP_CE(v1, v2) KSP_CE(v3, v5) KSP_CE(v4, v6) KSP_CE(v7, v8)
    v3 = fmaxf(v1, v3);
    KSP_CE(v2, v5) KSP_CE(v4, v7)
    v6 = fminf(v6, v8);
    v3 = fmaxf(v2, v3);
    v6 = fminf(v6, v7);
    KSP_CE(v3, v4) KSP_CE(v5, v6) KSP_CE(v4, v5)
    m0 = v3; m1 = v4; m2 = v5; m3 = v6;
}

// 3-input min / max: one FMNMX3 on sm_100a (the host build composes two).
KSP_HD float ksp_min3(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
#else
    return fminf(fminf(a, b), c);
#endif
}
KSP_HD float ksp_max3(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}

// Median of the 7 values {m0<=m1<=m2<=m3} U {x, p0<=p1}: the 4th smallest.
// With b0<=b1<=b2 the sorted extras (b0 = min(x,p0), b1 = clamp(x,p0,p1), b2 = max(x,p1)) it is
//   min( max(m0,b2), max(m1,b1), max(m2,b0), m3 )
// and the extras never need to be formed: 7 operations with 3-input min/max.
KSP_HD float median7_core4_extras3(float m0, float m1, float m2, float m3, float x, float p0,
                                   float p1)
{
    const float t0 = ksp_max3(m0, x, p1);                  // max(m0, b2)
    const float t1 = ksp_max3(m1, p0, fminf(x, p1));       // max(m1, b1)
    const float t2 = fmaxf(m2, fminf(x, p0));              // max(m2, b0)
    return fminf(ksp_min3(t0, t1, t2), m3);
}

// Two medians of 7 that share six of their values: {m0<=m1<=m2<=m3} U {p0<=p1} U {x} for
// x = xa, xb.  With r2 <= r3 the 3rd and 4th smallest of the six, the median is clamp(x, r2, r3):
// 6 operations for the pair of bounds, 2 per output.
KSP_HD void median7_pair(float m0, float m1, float m2, float m3, float p0, float p1, float xa,
                         float xb, float &oa, float &ob)
{
    const float r2 = ksp_min3(fmaxf(m0, p1), fmaxf(m1, p0), m2);
    const float r3 = ksp_min3(fmaxf(m1, p1), fmaxf(m2, p0), m3);
    oa = fmaxf(r2, fminf(xa, r3));
    ob = fmaxf(r2, fminf(xb, r3));
}

// Four medians of 13 from 16 consecutive samples: out[j] = median(e[j .. j+12]).
// All 16 samples must be ordinary numbers (no NaN).
KSP_HD void median13x4(const float (&e)[16], int rot, float &o0, float &o1, float &o2, float &o3)
{
#define E(k) e[((k) + rot) & 15]
    float m0, m1, m2, m3;
    mid4_of_10(E(3), E(4), E(5), E(6), E(7), E(8), E(9), E(10), E(11), E(12), m0, m1, m2, m3);
    // left extras {e0,e1,e2} / {e1,e2}; right extras {e13,e14} / {e13,e14,e15}
    const float p0 = fminf(E(1), E(2)), p1 = fmaxf(E(1), E(2));
    const float q0 = fminf(E(13), E(14)), q1 = fmaxf(E(13), E(14));
    median7_pair(m0, m1, m2, m3, p0, p1, E(0), E(13), o0, o1);   // extras e0,e1,e2 / e1,e2,e13
    median7_pair(m0, m1, m2, m3, q0, q1, E(2), E(15), o2, o3);   // extras e2,e13,e14 / e13,e14,e15
#undef E
}

// Ranks 3..6 of the union of a sorted six a0<=..<=a5 and a sorted four b0<=..<=b3: Batcher's
// odd-even merge pruned to the 16 min/max operations that reach those four outputs (derived
// and verified over all sorted 0-1 inputs by tools/median_network.py).
KSP_HD void mid4_of_sorted_6_4(float a0, float a1, float a2, float a3, float a4, float a5,
                               float b0, float b1, float b2, float b3, float &m0, float &m1,
                               float &m2, float &m3)
{
    b0 = fmaxf(a0, b0);
    KSP_CE(a4, b0)
    KSP_CE(a2, b2)
    a4 = fmaxf(a2, a4);
    b2 = fminf(b2, b0);
    b1 = fmaxf(a1, b1);
    a5 = fminf(a5, b1);
    a3 = fminf(a3, b3);
    KSP_CE(a3, a5)
    KSP_CE(a3, a4)
    KSP_CE(a5, b2)
    m0 = a3; m1 = a4; m2 = a5; m3 = b2;
}

// Eight medians of 13 from 20 consecutive samples: out[j] = median(e[j .. j+12]), j = 0..7.
// The windows of outputs 0..3 share e3..e12, those of outputs 4..7 share e7..e16; the six
// samples e7..e12 common to all eight are sorted once, each group adds its own four:
// the three samples particular to an output are a sorted pair shared with its neighbour plus
// one more: 24 + 2 x (10 + 16) + 2 x 2 + 4 x 6 + 8 x 2 = 120 operations, 15 per output
// (median13x4: 21).
// All 20 samples must be ordinary numbers (no NaN).
KSP_HD void median13x8(const float (&e)[20], float (&o)[8])
{
    // sorted six e7..e12 (12 comparators)
    float s0 = e[7], s1 = e[8], s2 = e[9], s3 = e[10], s4 = e[11], s5 = e[12];
    KSP_CE(s0, s5) KSP_CE(s1, s3) KSP_CE(s2, s4)
    KSP_CE(s1, s2) KSP_CE(s3, s4)
    KSP_CE(s0, s3) KSP_CE(s2, s5)
    KSP_CE(s0, s1) KSP_CE(s2, s3) KSP_CE(s4, s5)
    KSP_CE(s1, s2) KSP_CE(s3, s4)
    // group A: own four e3..e6; its first layer also sorts the pair (e5, e6) that group B needs
    float a0 = e[3], a1 = e[4], a2 = e[5], a3 = e[6];
    KSP_CE(a0, a1) KSP_CE(a2, a3)
    const float pb0 = a2, pb1 = a3;                       // sorted (e5, e6)
    KSP_CE(a0, a2) KSP_CE(a1, a3) KSP_CE(a1, a2)
    // group B: own four e13..e16; its first layer sorts the pair (e13, e14) that group A needs
    float b0 = e[13], b1 = e[14], b2 = e[15], b3 = e[16];
    KSP_CE(b0, b1) KSP_CE(b2, b3)
    const float qa0 = b0, qa1 = b1;                       // sorted (e13, e14)
    KSP_CE(b0, b2) KSP_CE(b1, b3) KSP_CE(b1, b2)
    float m0, m1, m2, m3;
    mid4_of_sorted_6_4(s0, s1, s2, s3, s4, s5, a0, a1, a2, a3, m0, m1, m2, m3);
    const float pa0 = fminf(e[1], e[2]), pa1 = fmaxf(e[1], e[2]);
    median7_pair(m0, m1, m2, m3, pa0, pa1, e[0], e[13], o[0], o[1]);  // extras e0,e1,e2 / e1,e2,e13
    median7_pair(m0, m1, m2, m3, qa0, qa1, e[2], e[15], o[2], o[3]);  // extras e2,e13,e14 / e13,e14,e15
    mid4_of_sorted_6_4(s0, s1, s2, s3, s4, s5, b0, b1, b2, b3, m0, m1, m2, m3);
    const float qb0 = fminf(e[17], e[18]), qb1 = fmaxf(e[17], e[18]);
    median7_pair(m0, m1, m2, m3, pb0, pb1, e[4], e[17], o[4], o[5]);  // extras e4,e5,e6 / e5,e6,e17
    median7_pair(m0, m1, m2, m3, qb0, qb1, e[6], e[19], o[6], o[7]);  // extras e6,e17,e18 / e17,e18,e19
}

// Generic median of up to 13 samples with a validity mask (bit k <-> w[k]).
// Returns false if no sample is valid.  lo/hi are the lower/upper medians (equal
// when the count is odd).  Slow path: ~250 operations.
KSP_HD bool median_masked13(const float (&w)[13], unsigned valid, float &lo, float &hi)
{
    const float inf = INFINITY;
    float s[13];
#pragma unroll
    for (int k = 0; k < 13; k++) s[k] = ((valid >> k) & 1u) ? w[k] : inf;
    // odd-even transposition sort, 13 rounds
#pragma unroll
    for (int r = 0; r < 13; r++) {
#pragma unroll
        for (int k = (r & 1); k + 1 < 13; k += 2) KSP_CE(s[k], s[k + 1])
    }
#ifdef __CUDA_ARCH__
    int n = __popc(valid & 0x1fffu);
#else
    int n = __builtin_popcount(valid & 0x1fffu);
#endif
    if (n == 0) return false;
    int il = (n - 1) >> 1, ih = n >> 1;
    lo = s[0];
    hi = s[0];
#pragma unroll
    for (int k = 1; k < 13; k++) {
        lo = (k == il) ? s[k] : lo;
        hi = (k == ih) ? s[k] : hi;
    }
    return true;
}

}  // namespace ksp
