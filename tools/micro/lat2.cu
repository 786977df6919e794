// The frequency-axis box-filter stage loop of csrc/twodflag.cu in isolation (developer tool):
// one warp, lane = 4 * line + stage, shuffle hand-over, shared-memory ring.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int VARIANT> __global__ void k(float b, long long *out, float *sink, int r2)
{
    __shared__ float ring_mem[64 * 128];
    __shared__ float tile[32 * 33];
    const int tid = threadIdx.x, lane = tid & 31, p = tid & 3, ll = tid >> 2;
    for (int i = tid; i < 64 * 128; i += blockDim.x) ring_mem[i] = 0.0f;
    for (int i = tid; i < 32 * 33; i += blockDim.x) tile[i] = b + i;
    __syncthreads();
    float *ring = ring_mem + tid, *rp = ring;
    const int rstride = blockDim.x;
    double s = 0.0; float e = 0.0f; int slot = 0;
    const bool head = p == 0, tail = p == 3;
    float *cell = tile + ll * 33;
    long long t0 = clock64();
    for (int it0 = 0; it0 < N; it0 += 32) {
#pragma unroll 8
        for (int kk = 0; kk < 32; kk++) {
            const float from_prev = __shfl_up_sync(0xffffffffu, e, 1);
            float x = from_prev;
            if (VARIANT != 2) { if (head) x = cell[kk]; }
            const double sa = __dadd_rn(s, (double) x);
            e = (float) sa;
            if (VARIANT == 0 || VARIANT == 2) { s = __dsub_rn(sa, (double) *rp); *rp = x; }
            else s = sa;
            if (VARIANT != 2) { if (tail) cell[kk] = e; }
            slot++; rp += rstride;
            if (slot == r2) { slot = 0; rp = ring; }
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    sink[tid] = (float) s + e;
}
int main()
{
    long long *out; float *sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 4096);
    const char *names[] = {"full stage (tile in/out, ring)", "no ring", "ring, no tile"};
#define RUN(V, threads) k<V><<<1, threads>>>(1.5f, out, sink, 18); cudaDeviceSynchronize(); printf("%-32s %4d threads: %7.1f cycles per iteration\n", names[V], threads, out[0] / (double) N);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(0, 128) RUN(1, 128) RUN(2, 128)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
