// Bit-identity of the constant-bank exp/expm1 (brian2_b200/csrc/b200_functions.cuh) with CUDA's
// library functions, the ones the parity tolerances were established with.  Prints
// "<function> <number of arguments> <number of mismatches>" per function.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "b200_functions.cuh"

__device__ __forceinline__ double make_arg(unsigned long long i, int mode) {
    // splitmix64
    unsigned long long z = i * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    if (mode == 0) return __longlong_as_double((long long)z);                   // any bit pattern
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);            // [0,1)
    if (mode == 1) return (u - 0.5) * 100.0;                                    // HH range
    if (mode == 2) return (u - 0.5) * 1500.0;                                   // incl. overflow
    return (u - 0.5) * 1e-3 * ((z & 1) ? 1.0 : 1e-12);                          // tiny
}

__global__ void check(unsigned long long n, int mode, unsigned long long* bad_exp,
                      unsigned long long* bad_expm1, unsigned long long* bad_exprel) {
    unsigned long long be = 0, bm = 0, br = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double x = make_arg(i, mode);
        const double a = exp(x), b = _b200_exp(x);
        if (__double_as_longlong(a) != __double_as_longlong(b) && !(isnan(a) && isnan(b))) be++;
        const double c = expm1(x), d = _b200_expm1(x);
        if (__double_as_longlong(c) != __double_as_longlong(d) && !(isnan(c) && isnan(d))) bm++;
        const double ref = fabs(x) < 1e-16 ? 1.0 : (x > 717 ? INFINITY : expm1(x) / x), e = _exprel(x);
        if (__double_as_longlong(ref) != __double_as_longlong(e) && !(isnan(ref) && isnan(e))) br++;
    }
    if (be) atomicAdd(bad_exp, be);
    if (bm) atomicAdd(bad_expm1, bm);
    if (br) atomicAdd(bad_exprel, br);
}

int main() {
    unsigned long long* d;
    if (cudaMalloc(&d, 3 * sizeof(unsigned long long)) != cudaSuccess) { printf("no device\n"); return 2; }
    const unsigned long long n = 1ULL << 24;
    unsigned long long tot[3] = {0, 0, 0}, args = 0;
    for (int mode = 0; mode < 4; ++mode) {
        cudaMemset(d, 0, 3 * sizeof(unsigned long long));
        check<<<592, 256>>>(n, mode, d, d + 1, d + 2);
        unsigned long long h[3];
        if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("kernel failed\n"); return 3; }
        for (int k = 0; k < 3; ++k) tot[k] += h[k];
        args += n;
    }
    printf("exp %llu %llu\nexpm1 %llu %llu\nexprel %llu %llu\n", args, tot[0], args, tot[1], args, tot[2]);
    return (tot[0] | tot[1] | tot[2]) ? 1 : 0;
}
