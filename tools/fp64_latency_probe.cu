// Dependent-issue latencies and single-warp throughputs of the instructions the forward sweep
// lives on (B200, sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
//   -o fp64_latency_probe tools/fp64_latency_probe.cu ; prints cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>

#define REP 512

template <int MODE>
__global__ void probe(double *out, long long *cyc, double seed, int ilp_unused)
{
    __shared__ double sh[64];
    double a = seed + threadIdx.x, b = 1.0000001, c = 0.9999999, d0 = a, d1 = a + 1, d2 = a + 2, d3 = a + 3;
    double e0 = a + 4, e1 = a + 5, e2 = a + 6, e3 = a + 7;
    sh[threadIdx.x & 63] = a;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < REP / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) a = a + b;                                   // DADD chain
            if (MODE == 1) a = a * b;                                   // DMUL chain
            if (MODE == 2) a = __fma_rn(a, b, c);                       // DFMA chain
            if (MODE == 3) { d0 += b; d1 += b; d2 += b; d3 += b; e0 += b; e1 += b; e2 += b; e3 += b; }   // 8 independent DADD
            if (MODE == 4) { double s; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(a)); a = s; }   // MUFU.RCP64H chain
            if (MODE == 5) { sh[threadIdx.x & 63] = a; __syncwarp(); a = sh[(threadIdx.x + 1) & 63]; }   // STS -> LDS round trip
            if (MODE == 6) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);                       // 2 SHFL
            if (MODE == 7) a = (a > b) ? a : c;                         // DSETP + 2 FSEL chain
            if (MODE == 8) { a = sh[((int)__double2loint(a)) & 63]; }   // dependent LDS chain
            if (MODE == 9) { a = a * b; a = a + c; }                    // DMUL -> DADD
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d0 + d1 + d2 + d3 + e0 + e1 + e2 + e3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter, int warps)
{
    double *out; long long *cyc, h[4];
    cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, 4 * sizeof(long long));
    probe<MODE><<<1, 32 * warps>>>(out, cyc, 1.5, 0);
    probe<MODE><<<1, 32 * warps>>>(out, cyc, 1.5, 0);
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM %2d: %7.2f cycles per instruction (%d per iteration)\n", name, warps, (double)h[0] / (REP * per_iter), per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 4, 8}) {
        if (w == 1) {
            run<0>("DADD dependent", 1, 1); run<1>("DMUL dependent", 1, 1); run<2>("DFMA dependent", 1, 1);
            run<9>("DMUL->DADD dependent", 2, 1);
            run<4>("MUFU.RCP64H dependent", 1, 1); run<5>("STS+syncwarp+LDS round trip", 1, 1);
            run<6>("SHFL (64-bit = 2) dependent", 1, 1); run<7>("DSETP+FSEL+FSEL dependent", 1, 1);
            run<8>("LDS dependent", 1, 1);
        }
        // one block: w warps spread over the 4 sub-partitions (w = 4: one per SMSP, w = 8: two per SMSP)
        switch (w) {
        case 1: run<3>("8 independent DADD", 8, 1); break;
        case 4: run<3>("8 independent DADD", 8, 4); break;
        default: run<3>("8 independent DADD", 8, 8); break;
        }
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
