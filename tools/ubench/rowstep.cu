// Microbenchmark: cycles per DP row of one strip warp for candidate MAS row-step codings,
// as a function of columns per lane (C) and of how many strip warps share an SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowstep rowstep.cu && ./rowstep
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float set_ge(float a, float b) {
    float d; asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}

// V0: FSETP + FSEL + FADD + predicated OR        (bits in an int)
// V1: FMNMX + FSET + FFMA(bits in a float) + FADD
// V2: FSETP + FSEL + FADD + ballot per column    (bits as C warp-wide words)
template <int C, int V>
__device__ __forceinline__ void row(float (&q)[C], const float (&x)[C], float left, uint32_t (&w)[C], uint32_t& bits) {
    if (V == 0) {
        uint32_t b = 0;
#pragma unroll
        for (int c = C - 1; c >= 0; --c) {
            const float a = c ? q[c - 1] : left, bb = q[c];
            const bool d = a >= bb;
            b |= d ? (1u << c) : 0u;
            q[c] = x[c] + (d ? a : bb);
        }
        bits = b;
    } else if (V == 1) {
        float acc = 8388608.0f;
#pragma unroll
        for (int c = C - 1; c >= 0; --c) {
            const float a = c ? q[c - 1] : left, bb = q[c];
            acc = fmaf(set_ge(a, bb), float(1u << c), acc);
            q[c] = x[c] + fmaxf(a, bb);
        }
        bits = __float_as_uint(acc);
    } else {
#pragma unroll
        for (int c = C - 1; c >= 0; --c) {
            const float a = c ? q[c - 1] : left, bb = q[c];
            const bool d = a >= bb;
            w[c] = __ballot_sync(0xffffffffu, d);
            q[c] = x[c] + (d ? a : bb);
        }
    }
}

template <int C, int V>
__global__ void k(float* out, long long* cyc, int rows) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    constexpr int RING = 64;
    const int pitch = nw * 32 * C;
    float* xs = sm;                                        // [RING][pitch]
    uint32_t* bits_s = reinterpret_cast<uint32_t*>(sm + RING * pitch);   // [rows][nw*C] words (wrapped to 256 rows)
    for (int i = threadIdx.x; i < RING * pitch; i += blockDim.x) xs[i] = -1.0f - float((i * 37) % 101) * 0.01f;
    __syncthreads();
    float q[C];
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = -float(lane * C + c);
    const float* xw = xs + warp * 32 * C + lane * C;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < rows; ++i) {
        float left = __shfl_up_sync(0xffffffffu, q[C - 1], 1);
        if (lane == 0) left = __int_as_float(0x7fffffff);
        float x[C];
        const float* xr = xw + (i & (RING - 1)) * pitch;
        if (C >= 4) {
#pragma unroll
            for (int v = 0; v < C / 4; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(xr + 4 * v);
                x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
            }
        } else if (C == 2) {
            const float2 t = *reinterpret_cast<const float2*>(xr);
            x[0] = t.x; x[1] = t.y;
        } else {
            x[0] = xr[0];
        }
        uint32_t w[C], bits = 0;
        row<C, V>(q, x, left, w, bits);
        uint32_t* brow = bits_s + (i & 255) * (nw * C) + warp * C;
        if (V == 2) {
            if (lane == 0) {
                if (C == 8) { *reinterpret_cast<uint4*>(brow) = make_uint4(w[0], w[1], w[2], w[3]); *reinterpret_cast<uint4*>(brow + 4) = make_uint4(w[4 % C], w[5 % C], w[6 % C], w[7 % C]); }
                else if (C == 4) *reinterpret_cast<uint4*>(brow) = make_uint4(w[0], w[1 % C], w[2 % C], w[3 % C]);
                else if (C == 2) *reinterpret_cast<uint2*>(brow) = make_uint2(w[0], w[1 % C]);
                else brow[0] = w[0];
            }
        } else {
            unsigned char* bb = reinterpret_cast<unsigned char*>(brow);
            if (C == 8) bb[lane] = (unsigned char)bits;
            else if (C == 4) { bits = (bits & 15u) | (__shfl_down_sync(0xffffffffu, bits, 1) << 4); if (!(lane & 1)) bb[lane >> 1] = (unsigned char)bits; }
            else if (C == 2) { bits = (bits & 3u) | (__shfl_down_sync(0xffffffffu, bits, 1) << 2); bits = (bits & 15u) | (__shfl_down_sync(0xffffffffu, bits, 2) << 4); if (!(lane & 3)) bb[lane >> 2] = (unsigned char)bits; }
            else { uint32_t b1 = __ballot_sync(0xffffffffu, bits & 1u); if (lane == 0) brow[0] = b1; }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += q[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(bits_s[threadIdx.x]);
    if (lane == 0) cyc[blockIdx.x * nw + warp] = t1 - t0;
}

template <int C, int V> void run(int nw) {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 4096);
    const int rows = 4096;
    size_t smem = sizeof(float) * 64 * nw * 32 * C + 4 * 256 * nw * C;
    cudaFuncSetAttribute(k<C, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) { k<C, V><<<1, 32 * nw, smem>>>(out, cyc, rows); cudaDeviceSynchronize(); }
    long long h[32]; cudaMemcpy(h, cyc, 8 * nw, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
    cudaError_t e = cudaGetLastError();
    printf("C=%d V=%d warps=%2d: %6.1f cycles/row  (%5.2f per cell-column)%s\n", C, V, nw, double(mx) / rows, double(mx) / rows / C,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

template <int C> void all() {
    for (int nw : {1, 4, 8, 16}) { run<C, 0>(nw); run<C, 1>(nw); run<C, 2>(nw); }
}

int main() {
    all<1>(); all<2>(); all<4>(); all<8>();
    return 0;
}
