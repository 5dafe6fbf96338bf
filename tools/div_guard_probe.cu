// Prints operand pairs for which the poison-guarded shared-divisor quotient (osc2_div.cuh) is not
// refused (NaN) but differs from a / b.  Expected output: "mismatches: 0".
#include <cstdio>
#include <cuda_runtime.h>
#include "../opensc2_b200/csrc/osc2_div.cuh"

__device__ unsigned long long sm64(unsigned long long &s)
{
    unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ double rnd(unsigned long long &s, int decades)
{
    const double m = 1.0 + (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0);
    const int e = (int)(sm64(s) % (unsigned)(2 * decades + 1)) - decades;
    const double v = m * pow(10.0, (double)e);
    return (sm64(s) & 1) ? -v : v;
}
__global__ void k(unsigned long long seed, int per, unsigned long long *count, double *ex)
{
    unsigned long long s = seed + 0x1234567ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
    for (int i = 0; i < per; ++i) {
        const int mode = (int)(sm64(s) & 3);
        double a = rnd(s, mode == 0 ? 320 : (mode == 1 ? 275 : 40)), b = rnd(s, mode == 3 ? 31 : 29);
        if ((sm64(s) & 63) == 0) a = 0.0;
        const Osc2Rcp r = osc2_rcp_prepare(b);
        const bool div_ok = osc2_abs_hi(b) - OSC2_HI_DIV_LO < OSC2_HI_DIV_SPAN;
        const double q = osc2_div_p(a, r), ref = a / b;      // NaN = refused
        if (div_ok && q == q && osc2_abs_hi(a) < OSC2_HI_2P900 &&
            __double_as_longlong(q) != __double_as_longlong(ref) && !(a == 0.0 && q == 0.0)) {
            const unsigned long long n = atomicAdd(count, 1ull);
            if (n < 8) { ex[4 * n] = a; ex[4 * n + 1] = b; ex[4 * n + 2] = q; ex[4 * n + 3] = ref; }
        }
    }
}
int main()
{
    unsigned long long *c, h = 0; double *ex, hx[32];
    cudaMalloc(&c, 8); cudaMemset(c, 0, 8); cudaMalloc(&ex, sizeof(hx));
    k<<<148 * 8, 128>>>(20261020ull, 4000, c, ex);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hx, ex, sizeof(hx), cudaMemcpyDeviceToHost);
    printf("mismatches: %llu of %llu\n", h, 148ull * 8 * 128 * 4000);
    for (unsigned long long i = 0; i < (h < 8 ? h : 8); ++i) printf("  a=%a b=%a q=%a ref=%a\n", hx[4 * i], hx[4 * i + 1], hx[4 * i + 2], hx[4 * i + 3]);
    return 0;
}
