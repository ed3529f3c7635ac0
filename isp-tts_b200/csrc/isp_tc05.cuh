// tcgen05 / TMEM / TMA wrappers shared by the GEMM-shaped kernels (isp_gemm.cu).  Inline PTX, sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace isp {
namespace tc {

// ---- TMA (tiled tensor maps) ---------------------------------------------------------------------------------------
ISP_DEVINL void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
ISP_DEVINL void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
ISP_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
ISP_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
ISP_DEVINL void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- TMEM ----------------------------------------------------------------------------------------------------------
ISP_DEVINL void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ISP_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ISP_DEVINL void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ISP_DEVINL void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

ISP_DEVINL void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// ---- MMA -----------------------------------------------------------------------------------------------------------
template <bool TF32>
ISP_DEVINL void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// arrives on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
ISP_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor, 128 B swizzle, Blackwell version bit.
//   K-major operand  (rows of 128 B = one row of the tile each, 8-row groups 1024 B apart): lbo unused, sbo = 1024.
//   MN-major operand (rows of 128 B = 64 bf16 / 32 tf32 consecutive M- or N-indices of ONE k; 8 k-rows form a 1024 B atom):
//                    lbo = distance between two 128 B-wide chunks along M/N, sbo = distance between 8-row groups along K.
//   MN-major TF32 operands only exist with the "128 B swizzle, 32 B atom" layout (layout type 1; TMA:
//   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32 B chunks are permuted within a 128 B row by (row % 4), the atom is 4 k-rows
//   (512 B), so sbo = 512 for a dense box.
ISP_DEVINL uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3ffff) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(layout_type) << 61;
    return d;
}

}  // namespace tc
}  // namespace isp
