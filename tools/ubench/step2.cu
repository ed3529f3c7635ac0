// Micro-benchmark of isp_mas2.cu's forward sweep in isolation: strip warps running strip_forward over a pre-filled ring
// (every flag already satisfied), optionally next to warps that poll the way the kernel's helpers do.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o step2 step2.cu -lcuda
//   ./step2
#include <cstdio>
#include <cstdarg>
#include "../../isp-tts_b200/csrc/isp_mas2.cu"

namespace isp {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t e, const char*) { return int(e); }
}

using namespace isp;
using namespace isp::mas2;

// mode bit 0: strips on the same SM sub-partition (warps 0, 4, 8, ..) instead of one per sub-partition
// mode bit 1: add pollers (one per sub-partition) that spin on a counter with nanosleep(20) + ld.acquire
__global__ void __launch_bounds__(512, 1) bench(int nstrips, int mode, int nch, long long* out, const float* gsrc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t sa = smem_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout: flags (256 B) | neg-inf page | per strip: V (kVBytes), W (nch+2 words x 32), ring 8 stages + mirror
    const uint32_t flags = sa, neginf = sa + 256;
    const int nstg = 8;
    const uint32_t per_strip = ((kVBytes + 127) & ~127) + uint32_t(nch + 2) * 128 + uint32_t(nstg + 1) * kR * 256;
    if (threadIdx.x < 16) reinterpret_cast<float*>(smem_raw + 256)[threadIdx.x] = -CUDART_INF_F;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) reinterpret_cast<int*>(smem_raw)[i] = 1 << 30;     // landed / prog: everything is there
        for (int i = 0; i < 64; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 512) + i, (1 << 20) - 1);
    }
    for (uint32_t i = threadIdx.x; i < uint32_t(nstrips) * per_strip / 4; i += blockDim.x) reinterpret_cast<float*>(smem_raw + 2048)[i] = -1.0f - float(i % 7);
    __syncthreads();
    const bool same = mode & 1;
    int strip = -1;
    if (same) { if ((warp & 3) == 0 && (warp >> 2) < nstrips) strip = warp >> 2; }
    else if (warp < nstrips) strip = warp;
    if (strip >= 0) {
        const uint32_t base = sa + 2048 + uint32_t(strip) * per_strip;
        StripCtx c;
        c.lane = lane; c.s = 0; c.nstg = nstg; c.nch = nch;
        c.pitchB = 256; c.ringB = uint32_t(nstg) * kR * 256;
        const uint32_t ring = base + ((kVBytes + 127) & ~127) + uint32_t(nch + 2) * 128;
        c.lane_ring = ring + lane * 8;
        c.full_s = sa + 512; c.empty_s = sa + 512 + 256; c.landed_s = flags; c.prog_sa = flags + 16 + 16 * strip;
        c.has_prev = false; c.has_next = false; c.w_ok = true;
        c.w_sa = base + ((kVBytes + 127) & ~127) + lane * 4; c.w_step = 128;
        c.v_rd = lane == 0 ? neginf : base + lane * kVLane;
        c.v_wr = base + (lane + 1) * kVLane;
        c.bnd_mine = base + 33 * kVLane; c.bnd_prev = c.bnd_mine;
        long long pc[2] = {0, 0};
        const long long t0 = clock64();
        strip_forward<true>(c, false, pc);
        const long long t1 = clock64();
        if (lane == 0) out[strip] = t1 - t0;
    } else if ((mode & 2) && warp >= 12) {
        // pollers: until strip 0 is through
        uint32_t spins = 0;
        while (ld_acquire_sa(flags + 16) < nch) { __nanosleep(20); if (++spins > (1u << 26)) break; }
    } else if ((mode & 4) && warp >= 12 && lane == 0) {
        // the loader's loop: two mbarrier tests and a sleep per round
        uint32_t spins = 0;
        while (ld_volatile_sa(flags + 16) < nch) {
            mbar_test_sa(sa + 512 + 8 * 40, 0); mbar_test_sa(sa + 512 + 8 * 41, 0);
            __nanosleep(100);
            if (++spins > (1u << 26)) break;
        }
    } else if ((mode & 8) && warp == 8 && lane == 0) {
        // a stream of bulk copies global -> shared next to the strips (4 KB every ~16 steps, like a TMA box per chunk)
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 512) + 50;
        uint32_t ph = 0, spins = 0;
        unsigned char* dst = smem_raw + 2048 + uint32_t(nstrips) * per_strip;     // scratch behind the strips
        mbar_init(bar, 1);
        fence_mbar_init();
        while (ld_volatile_sa(flags + 16) < nch) {
            mbar_arrive_expect_tx(bar, 4096);
            bulk_g2s(dst, gsrc, 4096, bar);
            while (!mbar_try_wait(bar, ph)) { if (++spins > (1u << 26)) return; }
            ph ^= 1;
            __nanosleep(200);
        }
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 64 * sizeof(long long));
    const int nch = 65;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    float* gsrc;
    cudaMalloc(&gsrc, 1 << 20);
    cudaMemset(gsrc, 0, 1 << 20);
    for (int mode : {0, 1, 4, 8, 12})
        for (int ns : {1, 4}) {
            if ((mode & 1) && ns == 1) continue;
            long long h[4] = {0, 0, 0, 0};
            for (int rep = 0; rep < 3; ++rep) {
                bench<<<1, 512, 200 * 1024>>>(ns, mode, nch, out, gsrc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            }
            printf("strips=%d mode=%d (1 same sub-partition, 2 acquire pollers, 4 mbarrier pollers, 8 bulk copies): %.1f cycles/step (strip 0), %.1f (last)\n", ns, mode,
                   double(h[0]) / (nch * 16), double(h[ns - 1]) / (nch * 16));
        }
    return 0;
}
