// Dependent-chain latencies of the instructions the 2-D flagger's recurrences are made of
// (developer tool): nvcc -arch=sm_100a -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(double a, float b, long long *out, double *sink)
{
    double x = a; float y = b; int z = (int) b;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __dadd_rn(x, a);
        if (OP == 1) { y = (float) x; x = (double) y; }                 // F2F.F32.F64 + F2F.F64.F32
        if (OP == 2) { y = (float) x; x = __dadd_rn((double) y, a); }   // the recurrences' e -> s hop
        if (OP == 3) y = __fadd_rn(y, b);
        if (OP == 4) y = __shfl_up_sync(0xffffffffu, y, 1) + b;
        if (OP == 5) z = z + (int) b + i;
        if (OP == 6) x = __dmul_rn(x, a);
        if (OP == 7) { z = __double2int_rn(x); x = (double) z + a; }
        if (OP == 8) { x = __dadd_rn(x, a); x = __dsub_rn(x, (double) b); }
        if (OP == 9) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x)); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x) : "f"(y)); x = __dadd_rn(x, a); y = (float) x; asm volatile("" ::: "memory"); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[threadIdx.x] = x + y + z;
}
int main()
{
    long long *out; double *sink;
    cudaMallocManaged(&out, 64 * 8); cudaMalloc(&sink, 1024 * 8);
    const char *names[] = {"DADD", "F2F.F32.F64 + F2F.F64.F32", "F2F + F2F + DADD", "FADD", "SHFL + FADD", "IADD x2", "DMUL", "D2I + I2D + DADD", "DADD + DSUB(cvt const)", "cvt+cvt+DADD+cvt"};
#define RUN(OP, threads) k<OP><<<1, threads>>>(1.0000001, 1.5f, out, sink); cudaDeviceSynchronize(); printf("%-32s %4d threads: %7.1f cycles per iteration\n", names[OP], threads, out[0] / (double) N);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32) RUN(9, 32)
    RUN(0, 128) RUN(1, 128) RUN(2, 128) RUN(0, 512) RUN(1, 512) RUN(2, 512) RUN(0, 1024) RUN(1, 1024) RUN(2, 1024)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
