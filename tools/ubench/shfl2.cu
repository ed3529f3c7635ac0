// Micro-benchmark behind isp_mas_cluster.cu's row step (one warp, registers + a small shared ring):
//   A  one row per shuffle round: 2 rotating SHFL, 2 FSEL, 2 x (FSETP, FMNMX, FADD, VOTE), predicated STS.64
//   B  two rows per shuffle round: 4 rotating SHFL (lanes l-1 and l-2), the left neighbour's row-r cell recomputed in the lane
//      (halo), 4 x (FSETP, FMNMX, FADD, VOTE)
//   C  dependent SHFL -> FADD chain (shuffle latency)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o shfl2 shfl2.cu && ./shfl2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

#define DEVINL __device__ __forceinline__
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
DEVINL void sts_u64_if(uint32_t sa, uint32_t lo, uint32_t hi, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t}" ::"r"(sa), "r"(lo), "r"(hi), "r"(uint32_t(pred)) : "memory");
}

constexpr int kRows = 64;       // ring rows (reused)
constexpr uint32_t kPitch = 512;

template <int MODE>
__global__ void bench(long long* cyc, float* out, int chunks) {
    __shared__ __align__(16) float ring[kRows * 128 + 4];
    __shared__ __align__(16) uint32_t bits[kRows * 4];
    const int lane = threadIdx.x;
    for (int i = lane; i < kRows * 128 + 4; i += 32) ring[i] = -1.0f - float(i % 7) * 0.25f;
    __syncwarp();
    const uint32_t ra = smem_u32(ring + 4) + lane * 4, ba = smem_u32(bits);
    const int s1 = (lane + 31) & 31, s2 = (lane + 30) & 31;
    const bool lane0 = lane == 0, lane1 = lane == 1;
    float q0 = -CUDART_INF_F, q1 = -CUDART_INF_F;
    __shared__ __align__(16) unsigned long long slots[64];
    __shared__ __align__(16) uint4 big[64];
    const uint32_t sbig = smem_u32(big);
    const uint32_t sslot = smem_u32(slots);
    uint32_t cslot; uint64_t gslot;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(cslot) : "r"(sslot));
    asm volatile("{\n\t.reg .u64 a;\n\tcvt.u64.u32 a, %1;\n\tcvta.shared::cluster.u64 %0, a;\n\t}" : "=l"(gslot) : "r"(cslot));
    uint32_t acc = 0, ke = 0, ko = 0;
    const long long t0 = clock64();
    for (int ch = 0; ch < chunks; ++ch) {
        const uint32_t xa = ra + uint32_t(ch & 3) * 16 * kPitch;
        if (MODE == 0) {
            float x0[16], x1[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { x0[k] = lds_f32(xa + k * kPitch); x1[k] = lds_f32(xa + k * kPitch + 128); }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float r0 = __shfl_sync(0xffffffffu, q0, s1), r1 = __shfl_sync(0xffffffffu, q1, s1);
                const float l0 = lane0 ? -CUDART_INF_F : r0, l1 = lane0 ? r0 : r1;
                const bool a = l0 >= q0, b = l1 >= q1;
                q0 = x0[k] + fmaxf(l0, q0);
                q1 = x1[k] + fmaxf(l1, q1);
                sts_u64_if(ba + uint32_t(k) * 16, __ballot_sync(0xffffffffu, a), __ballot_sync(0xffffffffu, b), lane0);
            }
        } else if (MODE == 1) {
            float x0[16], x1[16], h0[8], h1[8];
#pragma unroll
            for (int k = 0; k < 16; ++k) { x0[k] = lds_f32(xa + k * kPitch); x1[k] = lds_f32(xa + k * kPitch + 128); }
#pragma unroll
            for (int k = 0; k < 8; ++k) { h0[k] = lds_f32(xa + 2 * k * kPitch - 4); h1[k] = lds_f32(xa + 2 * k * kPitch + 124); }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float a1 = __shfl_sync(0xffffffffu, q0, s1), a2 = __shfl_sync(0xffffffffu, q0, s2);
                const float b1 = __shfl_sync(0xffffffffu, q1, s1), b2 = __shfl_sync(0xffffffffu, q1, s2);
                const float bA = -CUDART_INF_F, bB = -CUDART_INF_F;
                const float A1 = lane0 ? bA : a1, A2 = lane1 ? bA : a2;
                const float C1 = lane0 ? a1 : b1, C2 = (lane0 || lane1) ? a2 : b2;
                const float H0 = lane0 ? bB : h0[k] + fmaxf(A2, A1);
                const float H1 = h1[k] + fmaxf(C2, C1);
                const bool sa = A1 >= q0, sb = C1 >= q1;
                const float n0 = x0[2 * k] + fmaxf(A1, q0), n1 = x1[2 * k] + fmaxf(C1, q1);
                const bool sc = H0 >= n0, sd = H1 >= n1;
                q0 = x0[2 * k + 1] + fmaxf(H0, n0);
                q1 = x1[2 * k + 1] + fmaxf(H1, n1);
                sts_u64_if(ba + uint32_t(2 * k) * 16, __ballot_sync(0xffffffffu, sa), __ballot_sync(0xffffffffu, sb), lane0);
                sts_u64_if(ba + uint32_t(2 * k + 1) * 16, __ballot_sync(0xffffffffu, sc), __ballot_sync(0xffffffffu, sd), lane0);
            }
        } else if (MODE >= 4) {
            // E/F/G: two rows per round on ADJACENT columns (2 shuffles), as isp_mas_cluster.cu's step2; F adds lane 31's relaxed
            // cluster-scope 64-bit stores through a generic pointer (one per row), G the same stores as plain st.shared::cluster
            float2 x[16]; float h[8];
#pragma unroll
            for (int k = 0; k < 16; ++k) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x[k].x), "=f"(x[k].y) : "r"(xa + lane * 4 + k * kPitch));
#pragma unroll
            for (int k = 0; k < 8; ++k) h[k] = lds_f32(xa + lane * 4 + 2 * k * kPitch - 4);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t1 = __shfl_sync(0xffffffffu, q1, s1), t0 = __shfl_sync(0xffffffffu, q0, s1);
                const float L1 = lane0 ? -CUDART_INF_F : t1;
                const float hh = h[k] + fmaxf(t0, L1);
                const float H = lane0 ? -CUDART_INF_F : hh;
                const bool a0 = L1 >= q0, a1 = q0 >= q1;
                const float n0 = x[2 * k].x + fmaxf(L1, q0), n1 = x[2 * k].y + fmaxf(q0, q1);
                const bool b0 = H >= n0, b1 = n0 >= n1;
                q0 = x[2 * k + 1].x + fmaxf(H, n0);
                q1 = x[2 * k + 1].y + fmaxf(n0, n1);
                if (MODE == 8) {
                    acc += __ballot_sync(0xffffffffu, a0) + __ballot_sync(0xffffffffu, a1) + __ballot_sync(0xffffffffu, b0) + __ballot_sync(0xffffffffu, b1);
                } else if (MODE == 9) {
                    const uint32_t ea = __ballot_sync(0xffffffffu, a0), oa = __ballot_sync(0xffffffffu, a1);
                    const uint32_t eb = __ballot_sync(0xffffffffu, b0), ob = __ballot_sync(0xffffffffu, b1);
                    if (lane == 2 * k) { ke = ea; ko = oa; }
                    if (lane == 2 * k + 1) { ke = eb; ko = ob; }
                } else {
                sts_u64_if(ba + uint32_t(2 * k) * 16, __ballot_sync(0xffffffffu, a0), __ballot_sync(0xffffffffu, a1), lane0);
                sts_u64_if(ba + uint32_t(2 * k + 1) * 16, __ballot_sync(0xffffffffu, b0), __ballot_sync(0xffffffffu, b1), lane0);
                }
                if (MODE == 5) {
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 w, {%1, %2};\n\t@p st.relaxed.cluster.b64 [%0], w;\n\t}"
                                 ::"l"(gslot + k * 16), "r"(__float_as_uint(n1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 w, {%1, %2};\n\t@p st.relaxed.cluster.b64 [%0], w;\n\t}"
                                 ::"l"(gslot + k * 16 + 8), "r"(__float_as_uint(q1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                }
                if (MODE == 6) {
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 w, {%1, %2};\n\t@p st.shared::cluster.b64 [%0], w;\n\t}"
                                 ::"r"(cslot + k * 16), "r"(__float_as_uint(n1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 w, {%1, %2};\n\t@p st.shared::cluster.b64 [%0], w;\n\t}"
                                 ::"r"(cslot + k * 16 + 8), "r"(__float_as_uint(q1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                }
                if (MODE == 10) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.relaxed.cluster.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                                 ::"l"(gslot + k * 16), "r"(__float_as_uint(n1)), "r"(k), "r"(__float_as_uint(q1)), "r"(k + 1), "r"(uint32_t(lane == 31)) : "memory");
                }
                if (MODE == 11) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                                 ::"r"(sslot + k * 16), "r"(__float_as_uint(n1)), "r"(k), "r"(__float_as_uint(q1)), "r"(k + 1), "r"(uint32_t(lane == 31)) : "memory");
                }
                if (MODE == 12) {       // every lane stores (its own slot): is it the predicate?
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(sbig + lane * 16 + (k & 1) * 512), "r"(__float_as_uint(n1)), "r"(k), "r"(__float_as_uint(q1)), "r"(k + 1) : "memory");
                }
                if (MODE == 7) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t}"
                                 ::"r"(sslot + k * 16), "r"(__float_as_uint(n1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t}"
                                 ::"r"(sslot + k * 16 + 8), "r"(__float_as_uint(q1)), "r"(k), "r"(uint32_t(lane == 31)) : "memory");
                }
            }
            if (MODE == 9) sts_u64_if(ba + uint32_t(lane & 15) * 16, ke, ko, lane < 16);
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 16; ++k) q0 = __shfl_sync(0xffffffffu, q0, s1) + 1.0f;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) q0 = __shfl_up_sync(0xffffffffu, q0, 1) + 1.0f;
        }
    }
    const long long t1 = clock64();
    out[lane] = q0 + q1 + float(acc + ke + ko);
    if (lane == 0) cyc[MODE] = t1 - t0;
}

int main() {
    long long* cyc; float* out;
    cudaMalloc(&cyc, 128); cudaMalloc(&out, 1024);
    const int chunks = 256;
    for (int rep = 0; rep < 2; ++rep) {
        bench<0><<<1, 32>>>(cyc, out, chunks);
        bench<1><<<1, 32>>>(cyc, out, chunks);
        bench<2><<<1, 32>>>(cyc, out, chunks);
        bench<3><<<1, 32>>>(cyc, out, chunks);
        bench<4><<<1, 32>>>(cyc, out, chunks);
        bench<5><<<1, 32>>>(cyc, out, chunks);
        bench<6><<<1, 32>>>(cyc, out, chunks);
        bench<7><<<1, 32>>>(cyc, out, chunks);
        bench<8><<<1, 32>>>(cyc, out, chunks);
        bench<9><<<1, 32>>>(cyc, out, chunks);
        bench<10><<<1, 32>>>(cyc, out, chunks);
        bench<11><<<1, 32>>>(cyc, out, chunks);
        bench<12><<<1, 32>>>(cyc, out, chunks);
        cudaDeviceSynchronize();
    }
    long long h[13];
    cudaMemcpy(h, cyc, 104, cudaMemcpyDeviceToHost);
    const double rows = chunks * 16.0;
    printf("A one row per round : %.1f cycles/row\n", h[0] / rows);
    printf("B two rows per round: %.1f cycles/row\n", h[1] / rows);
    printf("C SHFL.IDX + FADD   : %.1f cycles/link\n", h[2] / rows);
    printf("D SHFL.UP + FADD    : %.1f cycles/link\n", h[3] / rows);
    printf("E two rows, adjacent columns, 2 SHFL: %.1f cycles/row\n", h[4] / rows);
    printf("F E + st.relaxed.cluster.b64 (generic) per row by lane 31: %.1f cycles/row\n", h[5] / rows);
    printf("G E + st.shared::cluster.b64 per row by lane 31: %.1f cycles/row\n", h[6] / rows);
    printf("H E + st.shared.v2 per row by lane 31: %.1f cycles/row\n", h[7] / rows);
    printf("I E without the bit stores: %.1f cycles/row\n", h[8] / rows);
    printf("J E, bits kept in registers (lane k keeps row k), one store per chunk: %.1f cycles/row\n", h[9] / rows);
    printf("K E + one st.relaxed.cluster.v4.b32 (generic) per round by lane 31: %.1f cycles/row\n", h[10] / rows);
    printf("L E + one st.shared.v4.b32 per round by lane 31: %.1f cycles/row\n", h[11] / rows);
    printf("M E + one st.shared.v4.b32 per round by every lane: %.1f cycles/row\n", h[12] / rows);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
