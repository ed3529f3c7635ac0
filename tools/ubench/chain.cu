// Microbenchmarks behind the MAS kernel's latency model (one warp unless stated):
//   A  dependent LOP3 chain                          -> ALU dependent-issue latency
//   B  backtrack row: P=R&~A; Rs=R>>1; R=P|(Rs&A1)   -> cycles/row, registers only
//   C  B + predicated STS per row,  D  C + LDS.128 of the windows per 2 rows
//   E  forward lane-skewed row step from shared memory (no TMA), 1/2/4 warps per CTA
//   F  mbarrier try_wait on a completed phase,  G  clock64 read
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain chain.cu && ./chain
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define DEVINL __device__ __forceinline__
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL float set_ge(float a, float b) { float d; asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
DEVINL uint32_t bt_step(uint32_t R, uint32_t A, uint32_t A1) {
    uint32_t P, Rs = R >> 1, out;
    asm("lop3.b32 %0, %1, %2, 0, 0x30;" : "=r"(P) : "r"(R), "r"(A));
    asm("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(out) : "r"(P), "r"(Rs), "r"(A1));
    return out;
}

__global__ void kA(uint32_t* out, long long* cyc, int iters) {
    uint32_t x = threadIdx.x, a = out[0], b = out[1];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
    long long t1 = clock64();
    out[threadIdx.x + 2] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
__global__ void kB(uint32_t* out, long long* cyc, int iters) {
    __shared__ __align__(16) uint32_t win[64];
    __shared__ uint32_t hist[32];
    const int lane = threadIdx.x;
    win[2 * lane] = out[lane] | 0x11111111u * (lane & 1);
    win[2 * lane + 1] = win[2 * lane] >> 1;
    __syncwarp();
    uint32_t A[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) A[i] = win[2 * i];
    uint32_t R = 0x80000000u, acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        R = 0x80000000u | (acc & 1);
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
            uint32_t a0, a1, a2, a3;
            if (MODE == 2) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(smem_u32(win + 2 * t))); }
            else { a0 = A[t]; a1 = A[t] >> 1; a2 = A[t + 1]; a3 = A[t + 1] >> 1; }
            if (MODE >= 1 && lane == 0) hist[t] = R;
            R = bt_step(R, a0, a1);
            if (MODE >= 1 && lane == 0) hist[t + 1] = R;
            R = bt_step(R, a2, a3);
        }
        acc += R;
        if (MODE >= 1) { __syncwarp(); acc += hist[lane]; }
    }
    long long t1 = clock64();
    out[threadIdx.x + 64] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C> DEVINL void lds_row(float (&x)[C], uint32_t saddr) {
    if (C >= 4) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(saddr));
        if (C == 8) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(x[4 % C]), "=f"(x[5 % C]), "=f"(x[6 % C]), "=f"(x[7 % C]) : "r"(saddr));
    } else {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x[0]), "=f"(x[1]) : "r"(saddr));
    }
}
DEVINL void sts_u8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
DEVINL void sts_f32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
DEVINL float lds_f32(uint32_t saddr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr)); return v; }
template <int C> DEVINL uint32_t dp_row(float (&q)[C], const float (&x)[C], float left) {
    float s[C];
#pragma unroll
    for (int c = C - 1; c >= 1; --c) { s[c] = set_ge(q[c - 1], q[c]); q[c] = x[c] + fmaxf(q[c - 1], q[c]); }
    s[0] = set_ge(left, q[0]);
    q[0] = x[0] + fmaxf(left, q[0]);
    if (C == 8) {
        const float t0 = fmaf(s[1], 2.0f, s[0]), t1 = fmaf(s[3 % C], 2.0f, s[2 % C]);
        const float t2 = fmaf(s[5 % C], 2.0f, s[4 % C]), t3 = fmaf(s[7 % C], 2.0f, s[6 % C]);
        const float u0 = fmaf(t1, 4.0f, t0), u1 = fmaf(t3, 4.0f, t2);
        return __float_as_uint(fmaf(u1, 16.0f, u0) + 8388608.0f);
    } else if (C == 4) {
        const float t0 = fmaf(s[1], 2.0f, s[0]), t1 = fmaf(s[3 % C], 2.0f, s[2 % C]);
        return __float_as_uint(fmaf(t1, 4.0f, t0) + 8388608.0f);
    } else {
        return __float_as_uint(fmaf(s[1], 2.0f, s[0]) + 8388608.0f);
    }
}

// E: the steady-state loop of the MAS strip warp.  VAR 0: as in the kernel; 1: no bits store; 2: no shuffle;
// 3: neighbour value through shared memory (STS + LDS) instead of SHFL
template <int C, int VAR>
__global__ void kE(float* out, long long* cyc, int steps, int pitch, int ring_rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    float* ring = reinterpret_cast<float*>(smem) + size_t(warp) * ring_rows * pitch;
    unsigned char* bits = smem + sizeof(float) * (size_t(nw) * ring_rows * pitch + 256) + size_t(warp) * 20 * 1024;
    float* xch = reinterpret_cast<float*>(bits + 18 * 1024);     // [2][33]
    for (int i = lane; i < ring_rows * pitch; i += 32) ring[i] = -1.0f - float((i * 37) % 101) * 0.01f;
    if (lane < 2) xch[lane * 33] = __int_as_float(0x7fffffff);
    __syncwarp();
    const uint32_t ring_sa = smem_u32(ring);
    const uint32_t pitchB = pitch * 4, ringB = ring_rows * pitchB, bpB = 32;
    uint32_t xoff = uint32_t((ring_rows - lane) % ring_rows) * pitchB + uint32_t((lane * C) % 128) * 4;
    uint32_t bits_sa = smem_u32(bits) + lane + 31 * 32 - lane * 32;
    const uint32_t xch_sa = smem_u32(xch) + lane * 4;
    float q[C], xc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = -float(lane * C + c);
    float left_cur = __int_as_float(0x7fffffff);
    lds_row<C>(xc, ring_sa + xoff);
    xoff += pitchB; if (xoff >= ringB) xoff -= ringB;
    long long t0 = clock64();
    for (int s0 = 0; s0 < steps; s0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float nxt = left_cur;
            if (VAR == 0 || VAR == 1) nxt = __shfl_up_sync(0xffffffffu, q[C - 1], 1);
            if (VAR == 3) { sts_f32(xch_sa + 4 + (k & 1) * 132, q[C - 1]); }
            float xn[C];
            lds_row<C>(xn, ring_sa + xoff);
            xoff += pitchB; if (xoff >= ringB) xoff -= ringB;
            if (VAR == 3) nxt = lds_f32(xch_sa + (k & 1) * 132);
            const uint32_t b = dp_row<C>(q, xc, left_cur);
            if (VAR != 1) sts_u8(bits_sa + uint32_t(k) * bpB, b);
            else if (b == 0x12345678u) sts_u8(bits_sa, b);
            left_cur = nxt;
            if (VAR != 3) left_cur = lane == 0 ? __int_as_float(0x7fffffff) : nxt;
#pragma unroll
            for (int c = 0; c < C; ++c) xc[c] = xn[c];
        }
        bits_sa += 8 * bpB;
        if ((s0 & 511) == 504) bits_sa -= 512 * bpB;
    }
    long long t1 = clock64();
    float sum = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) sum += q[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (lane == 0) cyc[warp] = t1 - t0;
}

__global__ void kF(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        acc += ok;
    }
    long long t1 = clock64();
    long long t2 = clock64();
    for (int it = 0; it < iters; ++it) acc += (uint32_t)clock64();
    long long t3 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
}

template <int C> void run_e(uint32_t* out, long long* cyc, int steps, int pitch) {
    long long h[32];
    for (int nw : {1, 2, 4, 8}) {
        const int ring_rows = 64;
        size_t smem = sizeof(float) * (size_t(nw) * ring_rows * pitch + 256) + size_t(nw) * 20 * 1024;
        if (smem > 227 * 1024) continue;
        double r[4];
        for (int var = 0; var < 4; ++var) {
            for (int rep = 0; rep < 2; ++rep) {
#define LAUNCH(V) { cudaFuncSetAttribute(kE<C, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kE<C, V><<<1, 32 * nw, smem>>>((float*)out, cyc, steps, pitch, ring_rows); }
                if (var == 0) LAUNCH(0) if (var == 1) LAUNCH(1) if (var == 2) LAUNCH(2) if (var == 3) LAUNCH(3)
            }
            cudaMemcpy(h, cyc, 8 * nw, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
            r[var] = double(mx) / steps;
        }
        cudaError_t e = cudaDeviceSynchronize();
        printf("E forward step C=%d pitch=%d warps=%d: %.1f cycles/step (no bits store %.1f, no shuffle %.1f, smem exchange %.1f) %s\n", C, pitch, nw,
               r[0], r[1], r[2], r[3], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMemset(out, 0x5a, 1 << 20); cudaMalloc(&cyc, 4096);
    long long h[32];
    const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) kA<<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("A dependent LOP3 chain: %.2f cycles/op\n", double(h[0]) / iters / 32);
    for (int rep = 0; rep < 2; ++rep) kB<0><<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("B backtrack row, registers only: %.2f cycles/row\n", double(h[0]) / iters / 32);
    for (int rep = 0; rep < 2; ++rep) kB<1><<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("C + predicated STS per row + hist readback: %.2f cycles/row\n", double(h[0]) / iters / 32);
    for (int rep = 0; rep < 2; ++rep) kB<2><<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("D + LDS.128 windows: %.2f cycles/row\n", double(h[0]) / iters / 32);
    for (int rep = 0; rep < 2; ++rep) kF<<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("F mbarrier.try_wait (completed phase): %.1f cycles;  G clock64: %.1f cycles\n", double(h[0]) / iters, double(h[1]) / iters);
    const int steps = 8192;
    run_e<8>(out, cyc, steps, 204); run_e<8>(out, cyc, steps, 132);
    run_e<4>(out, cyc, steps, 136); run_e<4>(out, cyc, steps, 72);
    run_e<2>(out, cyc, steps, 66);
    return 0;
}
