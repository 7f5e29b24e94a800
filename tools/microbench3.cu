// microbench3.cu -- is the MUFU.RSQ64H seed what keeps the inverse-cube chain below FP64 peak?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ double seed_rsq(double s) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s)); return y; }
template <int MODE> __device__ __forceinline__ double seed(double s, double k) {
    if (MODE == 0) return s * k;                 // 1 DMUL instead of the MUFU
    if (MODE == 1) return seed_rsq(s);           // MUFU.RSQ64H (+ IMAD.MOV for the low word)
    if (MODE == 2) { double y = seed_rsq(s); return y * k; }  // MUFU + 1 DMUL (break the hi/lo register pairing)
    return s;
}
template <int MODE, int NT>
__global__ void kD(int iters, const double* in, double* out) {
    double s[NT], acc = 0.0;
    for (int k = 0; k < NT; ++k) s[k] = in[k] * (1.0 + threadIdx.x) * 1e-18;
    const double step = in[30] * 1e-20, kk = in[29];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const double y0 = seed<MODE>(s[k], kk);
            const double t = y0 * y0; const double e = fma(-s[k], t, 1.0); const double c0 = fma(-3.0e-18, y0, 1.0);
            const double p = fma(e, fma(1.875, e, 1.5), c0);
            acc += (y0 * t) * p; s[k] += step;
        }
    }
    if (acc == 1.2345) out[0] = acc;
}
template <int MODE, int NT>
void run(const char* name, double fp64_each, const double* in, double* out, int sms, double ghz, int bps) {
    const int iters = 20000, threads = 128;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kD<MODE, NT><<<sms * bps, threads>>>(iters / 10, in, out);
    CK(cudaEventRecord(e0));
    kD<MODE, NT><<<sms * bps, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double wi = (double)sms * bps * (threads / 32) * (double)iters * NT * fp64_each;
    const double cyc = ms * 1e-3 * ghz * 1e9;
    printf("%-44s x%d/SM: %7.3f ms  fp64 pipe %.1f%%\n", name, bps, ms, 50.0 * wi / cyc / sms);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double ghz = clk * 1e-6; const int sms = p.multiProcessorCount;
    double h[32]; for (int q = 0; q < 32; ++q) h[q] = 1.0 + 1e-7 * q; h[30] = 3; h[29] = 1.0000001;
    double *in, *out; CK(cudaMalloc(&in, sizeof(h))); CK(cudaMalloc(&out, 8)); CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int bps = 4; bps <= 8; bps *= 2) {
        if (bps == 4) { run<0, 6>("no MUFU: seed = s*k (10 fp64 each)", 10, in, out, sms, ghz, 4); run<1, 6>("MUFU seed (9 fp64 each)", 9, in, out, sms, ghz, 4); run<2, 6>("MUFU seed * k (10 fp64 each)", 10, in, out, sms, ghz, 4); }
        else { run<0, 6>("no MUFU: seed = s*k (10 fp64 each)", 10, in, out, sms, ghz, 8); run<1, 6>("MUFU seed (9 fp64 each)", 9, in, out, sms, ghz, 8); run<2, 6>("MUFU seed * k (10 fp64 each)", 10, in, out, sms, ghz, 8); }
    }
    return 0;
}
