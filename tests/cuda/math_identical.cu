// Bit-identity of the constant-bank exp/expm1 (brian2_b200/csrc/b200_functions.cuh) with CUDA's
// library functions, the ones the parity tolerances were established with.  Prints
// "<function> <number of arguments> <number of mismatches>" per function, then the distance to
// the HOST's glibc (the arithmetic of the reference's cpp_standalone, i.e. of the oracle) over
// 10^7 arguments of the Hodgkin-Huxley range: "<function>_vs_glibc <n> <differing> <max ulp>".
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
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

// ---- against the ORACLE's arithmetic: the host's glibc (what cpp_standalone links) -----------
__global__ void eval_hh(unsigned long long n, double* x, double* e, double* m) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        x[i] = make_arg(i, 1);                 // the Hodgkin-Huxley argument range, |x| < 50
        e[i] = _b200_exp(x[i]);
        m[i] = _b200_expm1(x[i]);
    }
}

static long long ulp_distance(double a, double b) {
    long long ia, ib;
    memcpy(&ia, &a, 8);
    memcpy(&ib, &b, 8);
    if (ia < 0) ia = (long long)0x8000000000000000ULL - ia;
    if (ib < 0) ib = (long long)0x8000000000000000ULL - ib;
    return ia > ib ? ia - ib : ib - ia;
}

static void against_glibc() {
    const unsigned long long n = 10000000ULL;          // 10^7 arguments
    double *dx, *de, *dm;
    cudaMalloc(&dx, n * 8); cudaMalloc(&de, n * 8); cudaMalloc(&dm, n * 8);
    eval_hh<<<592, 256>>>(n, dx, de, dm);
    double* hx = (double*)malloc(n * 8); double* he = (double*)malloc(n * 8); double* hm = (double*)malloc(n * 8);
    cudaMemcpy(hx, dx, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(he, de, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hm, dm, n * 8, cudaMemcpyDeviceToHost);
    long long max_e = 0, max_m = 0;
    unsigned long long diff_e = 0, diff_m = 0;
    for (unsigned long long i = 0; i < n; ++i) {
        const long long ue = ulp_distance(he[i], exp(hx[i])), um = ulp_distance(hm[i], expm1(hx[i]));
        if (ue) diff_e++;
        if (um) diff_m++;
        if (ue > max_e) max_e = ue;
        if (um > max_m) max_m = um;
    }
    // "<function>_vs_glibc <arguments> <results that differ> <largest distance in ulp>"
    printf("exp_vs_glibc %llu %llu %lld\nexpm1_vs_glibc %llu %llu %lld\n", n, diff_e, max_e, n, diff_m, max_m);
    free(hx); free(he); free(hm); cudaFree(dx); cudaFree(de); cudaFree(dm);
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
    against_glibc();
    return (tot[0] | tot[1] | tot[2]) ? 1 : 0;
}
