// Microbenchmark: issue rate of ONE warp on an SM for instruction mixes like the MAS row step.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float set_ge(float a, float b) { float d; asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.5f + i; b[i] = 1.0f + i; }
    float acc = 8388608.0f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = a[i] + b[i]; }                                   // 8 independent FADD
            if (MODE == 1) { a[i] = fmaxf(a[i], b[i]) ; b[i] = b[i] + 1.0f; }          // FMNMX + FADD
            if (MODE == 2) { a[i] = fmaxf(a[i], b[(i + 1) & 7]); }                    // 8 independent FMNMX
            if (MODE == 3) { float m = fmaxf(a[i], a[(i + 7) & 7]); acc = fmaf(set_ge(a[(i + 7) & 7], a[i]), 2.0f, acc); b[i] = b[i] + m; }
            if (MODE == 4) { a[i] = fmaf(a[i], 1.0001f, b[i]); }                       // 8 independent FFMA
        }
        if (MODE == 3) { for (int i = 0; i < 8; ++i) { float t = a[i]; a[i] = b[i]; b[i] = t; } }
    }
    long long t1 = clock64();
    float s = acc;
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int instr_per_iter, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64 * 8);
    int iters = 4096;
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%d: %.2f cycles/iter, %.2f cycles/instr\n", name, warps, double(h) / iters, double(h) / iters / instr_per_iter);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8}) {
        if (w == 1) { run<0>("8 FADD", 8, 1); run<2>("8 FMNMX", 8, 1); run<4>("8 FFMA", 8, 1); run<1>("8 FMNMX + 8 FADD", 16, 1); run<3>("MAS-like 32 instr", 32, 1); }
        if (w == 4) { run<0>("8 FADD", 8, 4); run<2>("8 FMNMX", 8, 4); run<1>("8 FMNMX + 8 FADD", 16, 4); run<3>("MAS-like 32 instr", 32, 4); }
        if (w == 8) { run<0>("8 FADD", 8, 8); run<2>("8 FMNMX", 8, 8); run<1>("8 FMNMX + 8 FADD", 16, 8); run<3>("MAS-like 32 instr", 32, 8); }
    }
    return 0;
}
