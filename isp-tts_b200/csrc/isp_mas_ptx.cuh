// PTX helpers shared by the MAS kernels (isp_mas.cu, isp_mas2.cu): tiled TMA loads, mbarrier forms on
// shared-space addresses, predicated single-instruction stores, bulk shared->global copies.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace isp {

// ---- small PTX helpers ---------------------------------------------------------------
ISP_DEVINL void tma_load_box(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
        : "memory");
}
ISP_DEVINL int ld_volatile_sa(uint32_t saddr) {
    int v;
    asm volatile("ld.volatile.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
// predicated forms: a divergent `if (lane == ...)` around one instruction costs BSSY/BSYNC and a branch.  The counter
// stores are plain st.shared inside `asm volatile` (the compiler keeps their place; st.volatile would add a MEMBAR)
ISP_DEVINL void mbar_arrive_if_sa(uint32_t bar, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 st;\n\tsetp.ne.u32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 st, [%0];\n\t}"
                 ::"r"(bar), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void st_volatile_if_sa(uint32_t saddr, int v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared::cta.s32 [%0], %1;\n\t}"
                 ::"r"(saddr), "r"(v), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void sts_f32_if(uint32_t saddr, float v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}"
                 ::"r"(saddr), "f"(v), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void stg_u16_if(int16_t* gp, int v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u16 [%0], %1;\n\t}" ::"l"(gp), "h"(short(v)), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void stg_s64_if(int64_t* gp, int64_t v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.s64 [%0], %1;\n\t}" ::"l"(gp), "l"(v), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void sts_u64(uint32_t saddr, uint32_t lo, uint32_t hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(lo), "r"(hi) : "memory");
}
ISP_DEVINL void st_volatile_sa(uint32_t saddr, int v) {
    asm volatile("st.shared::cta.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
ISP_DEVINL void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
ISP_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
ISP_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
ISP_DEVINL void cp_async4(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(gsrc) : "memory");
}
ISP_DEVINL void cp_async_arrive_noinc_sa(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
ISP_DEVINL void mbar_arrive_sa(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
ISP_DEVINL void mbar_expect_tx_sa(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
ISP_DEVINL uint32_t mbar_test_sa(uint32_t bar, uint32_t parity) {     // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
ISP_DEVINL uint32_t mbar_try_sa(uint32_t bar, uint32_t parity) {      // may sleep in hardware
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// the loaders' wait: sleeps in hardware (up to ~1 us per try) instead of spinning on an issue port a strip warp needs
ISP_DEVINL void mbar_wait_idle_sa(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000u) : "memory");
        if (++spins > (1u << 24)) __trap();
    } while (!ok);
}
// spin with a watchdog: a protocol bug must surface as a launch failure, not as a hung GPU
ISP_DEVINL void mbar_wait_sa(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_sa(bar, parity)) { if (++spins > (1u << 24)) __trap(); }
}
ISP_DEVINL float set_ge(float a, float b) {   // 1.0f if a >= b (false on NaN), else 0.0f: one FSET
    float d;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
ISP_DEVINL float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
ISP_DEVINL void sts_f32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
ISP_DEVINL void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
ISP_DEVINL uint4 lds_v4(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
ISP_DEVINL void st_release_sa(uint32_t saddr, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
ISP_DEVINL int ld_acquire_sa(uint32_t saddr) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}

// one backtrack row on a one-hot position: stay where A is 0, move one column down where A is 1.
// Written as two LOP3 levels so that the dependent chain is 2 ALU ops per row, not 3.
ISP_DEVINL uint32_t bt_step(uint32_t R, uint32_t A, uint32_t A1) {
    uint32_t P, Rs = R >> 1, out;
    asm("lop3.b32 %0, %1, %2, 0, 0x30;" : "=r"(P) : "r"(R), "r"(A));              // R & ~A
    asm("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(out) : "r"(P), "r"(Rs), "r"(A1));   // P | (Rs & A1)
    return out;
}

}  // namespace isp
