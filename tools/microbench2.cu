// microbench2.cu -- which instructions steal FP64-pipe issue bandwidth on B200?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define FMA(a,b,c) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a) : "d"(b), "d"(c))
#define ADD(a,b)   asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a) : "d"(b))
#define MUL(a,b)   asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a) : "d"(b))
#define RSQ(y,s)   asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s))
#define IMAX(x,y)  asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define IADD(x,y)  asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define FFMA(a,b,c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c))

template <int MODE>
__global__ void k(int iters, const double* in, double* out) {
    double a[8], b = in[8], c = in[16], y[4] = {0, 0, 0, 0}, s[4];
    int x[4]; float f[4];
    for (int q = 0; q < 8; ++q) a[q] = in[q] + threadIdx.x;
    for (int q = 0; q < 4; ++q) { s[q] = in[q] * (threadIdx.x + 1.0); x[q] = threadIdx.x + q; f[q] = threadIdx.x + q; }
    const int xi = (int)in[30]; const float ff = (float)in[29];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); }                       // 8 DFMA
            if (MODE == 1) { for (int q = 0; q < 8; ++q) ADD(a[q], c); }                          // 8 DADD
            if (MODE == 2) { for (int q = 0; q < 8; ++q) MUL(a[q], b); }                          // 8 DMUL
            if (MODE == 3) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); RSQ(y[u], s[u]); }      // 8 DFMA + 1 MUFU.RSQ64H
            if (MODE == 4) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); RSQ(y[u], s[u]); RSQ(y[(u+1)&3], s[(u+1)&3]); } // 8 + 2 MUFU
            if (MODE == 5) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); IMAX(x[u], xi); }       // 8 DFMA + 1 ALU
            if (MODE == 6) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); IMAX(x[0], xi); IMAX(x[1], xi); IMAX(x[2], xi); IMAX(x[3], xi); } // 8 + 4 ALU
            if (MODE == 7) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); FFMA(f[0], ff, ff); FFMA(f[1], ff, ff); FFMA(f[2], ff, ff); FFMA(f[3], ff, ff); } // 8 + 4 FFMA
            if (MODE == 8) { for (int q = 0; q < 8; ++q) FMA(a[q], b, c); FFMA(f[0], ff, ff); FFMA(f[1], ff, ff); FFMA(f[2], ff, ff); FFMA(f[3], ff, ff);
                             FFMA(f[0], ff, ff); FFMA(f[1], ff, ff); FFMA(f[2], ff, ff); FFMA(f[3], ff, ff); } // 8 + 8 FFMA
            if (MODE == 9) { for (int q = 0; q < 8; ++q) { FMA(a[q], b, c); } for (int q = 0; q < 4; ++q) ADD(a[q], c); for (int q = 4; q < 8; ++q) MUL(a[q], b); } // 8 DFMA+4 DADD+4 DMUL
        }
    }
    double r = 0; for (int q = 0; q < 8; ++q) r += a[q];
    for (int q = 0; q < 4; ++q) r += y[q] + x[q] + f[q];
    if (r == 1.2345) out[0] = r;
}

template <int MODE>
void run(const char* name, double fp64_per_u, const double* in, double* out, int sms, double ghz, int threads = 128, int bps = 4) {
    const int iters = 10000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<sms * bps, threads>>>(iters / 10, in, out);
    CK(cudaEventRecord(e0));
    k<MODE><<<sms * bps, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double wi = (double)sms * bps * (threads / 32) * (double)iters * 4.0 * fp64_per_u;
    const double cyc = ms * 1e-3 * ghz * 1e9;
    printf("%-40s thr %d x %d: %7.3f ms  fp64 warp-instr/cyc/SM %.3f (%.1f%%)  cycles per u per SMSP %.2f\n", name, threads, bps, ms, wi / cyc / sms,
           50.0 * wi / cyc / sms, cyc / (iters * 4.0) / ((double)bps * (threads / 32) / 4.0));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double ghz = clk * 1e-6; const int sms = p.multiProcessorCount;
    double h[32]; for (int q = 0; q < 32; ++q) h[q] = 1.0 + 1e-7 * q; h[16] = 1e-9; h[30] = 3; h[29] = 1.0000001;
    double *in, *out; CK(cudaMalloc(&in, sizeof(h))); CK(cudaMalloc(&out, 8)); CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    run<0>("8 DFMA", 8, in, out, sms, ghz);
    run<1>("8 DADD", 8, in, out, sms, ghz);
    run<2>("8 DMUL", 8, in, out, sms, ghz);
    run<9>("8 DFMA + 4 DADD + 4 DMUL", 16, in, out, sms, ghz);
    run<3>("8 DFMA + 1 MUFU.RSQ64H", 8, in, out, sms, ghz);
    run<4>("8 DFMA + 2 MUFU.RSQ64H", 8, in, out, sms, ghz);
    run<5>("8 DFMA + 1 IMNMX", 8, in, out, sms, ghz);
    run<6>("8 DFMA + 4 IMNMX", 8, in, out, sms, ghz);
    run<7>("8 DFMA + 4 FFMA", 8, in, out, sms, ghz);
    run<8>("8 DFMA + 8 FFMA", 8, in, out, sms, ghz);
    run<3>("8 DFMA + 1 MUFU.RSQ64H  (8 CTA/SM)", 8, in, out, sms, ghz, 128, 8);
    run<4>("8 DFMA + 2 MUFU.RSQ64H  (8 CTA/SM)", 8, in, out, sms, ghz, 128, 8);
    return 0;
}
