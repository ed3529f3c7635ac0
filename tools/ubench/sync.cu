// Microbenchmarks of the synchronisation / async-copy primitives the MAS kernel leans on (one warp, cycles per op):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync sync.cu && ./sync
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define DEVINL __device__ __forceinline__
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(const float* src, float* out, long long* cyc, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ int flag[4];
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"((1 << 20) - 1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        flag[0] = 0;
    }
    for (int i = lane; i < 1024; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t b0 = smem_u32(&bar[0]);
    const uint32_t sm = smem_u32(smem);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {           // mbarrier.arrive (one lane)
            if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b0) : "memory");
        } else if (MODE == 1) {    // cp.async.mbarrier.arrive.noinc, nothing outstanding
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b0) : "memory");
        } else if (MODE == 2) {    // test_wait, result used at once
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b0), "r"(1) : "memory");
            acc += ok;
            if (acc == 0x7fffffff) break;
        } else if (MODE == 3) {    // try_wait on a completed phase, result used at once
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b0), "r"(1) : "memory");
            acc += ok;
            if (acc == 0x7fffffff) break;
        } else if (MODE == 4) {    // st.release.cta.shared
            asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(&flag[0])), "r"(it) : "memory");
        } else if (MODE == 5) {    // ld.acquire.cta.shared, result used at once
            int v;
            asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag[0])) : "memory");
            acc += v;
            if (acc == 0x7fffffff) break;
        } else if (MODE == 6) {    // 16 B cp.async, 2 per lane, + noinc arrive (the loader's inner pattern), L2-resident source
            const float* g = src + ((it & 63) * 1024) + lane * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm + lane * 16), "l"(g) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm + 512 + lane * 16), "l"(g + 128) : "memory");
            if ((it & 7) == 7) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b0) : "memory");
        } else if (MODE == 7) {    // bulk shared->global 4 KB + commit (one lane)
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (it & 63) * 1024), "r"(sm), "r"(4096) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (MODE == 8) {    // __syncwarp + elected arrive (stage release pattern)
            __syncwarp();
            if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b0) : "memory");
        } else if (MODE == 9) {    // test_wait issued, result consumed one iteration later
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b0), "r"(1) : "memory");
            if (acc == 0x7fffffff) break;
            acc += ok;
            // 40 independent FMAs of filler so that the latency can hide
            float f = float(it);
#pragma unroll
            for (int q = 0; q < 40; ++q) f = fmaf(f, 1.0001f, 0.5f);
            if (f == 123.0f) acc++;
        } else if (MODE == 10) {   // the same filler alone
            float f = float(it);
#pragma unroll
            for (int q = 0; q < 40; ++q) f = fmaf(f, 1.0001f, 0.5f);
            if (f == 123.0f) acc++;
        }
    }
    long long t1 = clock64();
    if (MODE == 6) asm volatile("cp.async.wait_all;" ::: "memory");
    if (MODE == 7 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    out[65536 + lane] = float(acc);
    if (lane == 0) cyc[0] = t1 - t0;
}

int main() {
    float *src, *out; long long* cyc;
    cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const char* names[] = {"mbarrier.arrive (1 lane)", "cp.async.mbarrier.arrive.noinc (idle)", "test_wait -> use", "try_wait (complete) -> use",
                           "st.release.cta.shared", "ld.acquire.cta.shared -> use", "2 x cp.async.cg 16 B (+ noinc arrive / 8)",
                           "bulk s2g 4 KB + commit", "__syncwarp + elected arrive", "test_wait + 40 FMA filler (use next iter)", "40 FMA filler alone"};
    const int iters = 4096;
    long long h;
#define RUN(M) for (int rep = 0; rep < 2; ++rep) k<M><<<1, 32, 16384>>>(src, out, cyc, iters); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-48s %.1f cycles/iter  %s\n", names[M], double(h) / iters, cudaGetErrorString(cudaGetLastError()));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    return 0;
}
