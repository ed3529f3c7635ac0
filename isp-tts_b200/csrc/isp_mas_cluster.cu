// Monotonic Alignment Search for long / wide utterances on sm_100a: one thread-block CLUSTER per utterance.
//
// Reference semantics (paths relative to the reference root):
//   tts/modules/aligner/mas.py:8-26        mas_width1 (DP, tie rule, backtrack)
//   tts/modules/aligner/cuda_mas.py:11-46  cuda_b_mas (the GPU kernel this replaces; 256 columns per launch configuration)
//   tts/models/acoustic/modules/alignment.py:275  durations = attn_hard.sum(dim=1)
//
// Where the strip kernels (isp_mas2.cu: <= 256 tokens, isp_mas.cu: <= 640) keep an utterance on ONE SM, a long-form batch
// (BASELINE configs[3]: 16 utterances of 512 tokens x 4096 frames) leaves most of the 148 SMs idle and makes one SM carry
// 8 strips.  Here the token axis is cut into CTAs of 128 columns, 2..8 CTAs form a cluster, and everything that crosses a CTA
// boundary goes through distributed shared memory:
//
//   * Forward.  A strip warp owns 64 columns, lane l the columns 2l and 2l + 1; one exchange with the left neighbour (two
//     rotating shuffles) advances TWO rows, the neighbour's cell in between being recomputed in the lane.  Strip g + 1 runs one
//     chunk of 32 rows behind strip g (a wavefront over strips); the one value that crosses a strip boundary per row --
//     Q[i][last column] -- is written by the producer's lane 31 straight into the CONSUMER's shared memory (st.relaxed.cluster
//     through a generic pointer, the same code for a neighbour warp and a neighbour CTA) as a 64-bit word {value, row tag}: the
//     tag makes the slot its own flag, so the chain warps never execute a fence.  Logits arrive as tiled TMA boxes of 32 rows x
//     128 columns, a ring of 128 rows; a loader warp turns the mbarriers into a plain counter (an mbarrier test costs a chain
//     warp ~100 cycles).
//   * Backpointers: 1 bit per cell in the owning CTA's shared memory (16 B per row and CTA) -- never in HBM.  The strips leave
//     them as {even columns, odd columns} words; the mapper interleaves them into column order, one row per lane.
//   * Backtrack, parallel over blocks of 32 rows.  j <- j - bit[i][j] is T1 dependent lookups; instead a mapper warp per CTA
//     composes, per block, the map "column at the block's last row -> column above its first row" while the sweep is still
//     running, bit-sliced: plane k of the map holds bit k of the target column for all 128 source columns, and a row is
//     plane' = select(bits, plane << 1, plane) -- one funnel shift and one LOP3 per 32 columns.  Only the low 6 bits of the
//     target are kept (a block moves a column by at most 32, so the source column disambiguates).  The shift crosses CTA
//     boundaries: the mapper of CTA c hands the 32 values of its last column to CTA c + 1 per block (DSMEM, tagged slots again).
//     After the sweep T1/32 serial hops remain, handed from CTA to CTA as the path moves left; then every block walks its
//     own 32 rows (one thread per block, bits read through ld.shared::cluster from whichever CTA owns the column).
//   * Outputs: each CTA zero-fills a slice of the utterance's dense int16 rows under the sweep, writes that slice's ones after
//     the walk, and the durations of its own 128 tokens from the path's change points.
//
// Bit-exactness: one fp32 add per cell on top of an exact max, the reference's `>=` (ties and -inf >= -inf take the
// diagonal), column 0 never moves (mas.py:16); everything after the bits is integer arithmetic.

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_mas_ptx.cuh"

#ifndef ISP_MASC_BNDSLEEP
#define ISP_MASC_BNDSLEEP 20    // ns between two looks at the strip on the left's slots
#endif
#ifndef ISP_MASC_LDSLEEP
#define ISP_MASC_LDSLEEP 40     // ns between two rounds of the loader when it had nothing to do
#endif
#ifndef ISP_MASC_MAPPOLL
#define ISP_MASC_MAPPOLL 200    // ns between two looks of the mapper at its strips' progress
#endif
#ifndef ISP_MASC_FILLPACE
#define ISP_MASC_FILLPACE 100   // ns between two 16 B zero-fill stores of a filler thread
#endif
#ifndef ISP_MASC_PF
#define ISP_MASC_PF 2           // A/B: the next chunk's logit loads 0 in front of each round's shuffles, 1 after the round's arithmetic, 2 none (burst at the chunk's top)
#endif
#ifndef ISP_MASC_BITS_ROW
#define ISP_MASC_BITS_ROW 1     // A/B: 1 = lane 0 stores each row's backpointer words, 0 = lane k keeps row k in registers until the chunk's end
#endif
#ifndef ISP_MASC_PUT_END
#define ISP_MASC_PUT_END 1      // A/B: 1 = the chunk's boundary stores in one burst after its rounds, 0 = one per round
#endif

namespace isp {
namespace masc {

constexpr int kStrip = 64;              // columns per strip warp
constexpr int kColsCta = 128;           // two strips per CTA
#ifndef ISP_MASC_KCH
#define ISP_MASC_KCH 32
#endif
constexpr int kCh = ISP_MASC_KCH;                 // rows per chunk = rows per mbarrier
constexpr int kBnd = 256;               // slots of a strip's boundary-in ring (8 B each)
constexpr int kPlanes = 6;              // low bits of a column index kept in a block map
constexpr int kBlk = 32;                // rows per backtrack block
constexpr int kThreads = 256;           // warps: 0, 1 strips; 2 loader; 3 mapper + hops; 4..7 zero fill
constexpr int kMaxCluster = 8;
constexpr int kMaxStages = 128 / kCh;           // ring of at most 128 rows
constexpr uint32_t kRowBytes = kColsCta * 4;

// control block (32-bit words from off_ctl)
constexpr uint32_t kCtlLanded = 0;      // rows of logits in the ring (loader -> strips)
constexpr uint32_t kCtlProg = 1;        // [2] rows completed by the CTA's strips (-> loader, mapper)
constexpr uint32_t kCtlCons = 3;        // [2] rows the consumer of strip s's boundary values has completed (written by the consumer)
constexpr uint32_t kCtlMapIn = 5;       // blocks whose carry-in words have arrived from the CTA on the left
constexpr uint32_t kCtlHopFlag = 6;     // 1: take the hop chain over (kCtlHopBlk, kCtlHopJ); 2: the chain is complete
constexpr uint32_t kCtlHopBlk = 7;
constexpr uint32_t kCtlHopJ = 8;
constexpr uint32_t kCtlFull = 16;       // [kMaxStages] mbarriers (8 B each)
constexpr uint32_t kCtlBytes = 256;

struct Params {
    const float* logp;
    int64_t sB, sT1;
    const int64_t* text_len;
    const int64_t* mel_len;
    int B, T1max, T2max;
    int16_t* hard;
    int64_t* dur;
    int16_t* path;          // (B, T1max): the caller's, or scratch in the workspace
    int path_is_output;     // the caller wants -1 past mel_len
    int* status;
    int nc;                 // CTAs per cluster
    int ring_rows;          // power of two, multiple of kCh
    int bulk;               // 1: rows are 16 B aligned -> tiled TMA boxes; 0: the loader warp copies through registers
    uint32_t off_bnd, off_ent, off_carry, off_maps, off_bits, off_ring;
    int dbg;                // debug ("masc.dbg"): 1 = a strip without a right neighbour still publishes its boundary values (timing)
    long long* trace;       // debug ("masc.trace"): (B * nc, 16) timestamps and wait cycles per CTA, or nullptr
};

// ---- cluster-scope PTX ------------------------------------------------------------------------------------------
ISP_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
ISP_DEVINL uint32_t mapa(uint32_t sa, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(sa), "r"(rank));
    return r;
}
ISP_DEVINL void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ISP_DEVINL void st_cluster_u32(uint32_t ca, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ca), "r"(v) : "memory"); }
ISP_DEVINL void st_cluster_u16(uint32_t ca, int v) { asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(ca), "h"(short(v)) : "memory"); }
ISP_DEVINL uint32_t ld_cluster_u32(uint32_t ca) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ca) : "memory");
    return v;
}
ISP_DEVINL int ld_cluster_u16(uint32_t ca) {
    unsigned short v;
    asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(ca) : "memory");
    return int(v);
}
ISP_DEVINL void st_release_cluster(uint32_t ca, int v) {
    asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(ca), "r"(v) : "memory");
}
ISP_DEVINL int ld_acquire_cluster_sa(uint32_t sa) {        // a word of THIS CTA's shared memory that another CTA releases
    int v;
    asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
ISP_DEVINL void st_relaxed_cluster(uint32_t ca, int v) {
    asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(ca), "r"(v) : "memory");
}
ISP_DEVINL int ld_relaxed_cluster_sa(uint32_t sa) {
    int v;
    asm volatile("ld.relaxed.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
// boundary slot {value, row tag}: one 64-bit relaxed store by the producer's lane 31, one 64-bit relaxed load by the consumer
// (through a GENERIC pointer computed once: a shared::cluster operand makes the compiler rebuild the 64-bit address -- an S2R of
// the shared window -- in front of every store)
ISP_DEVINL uint64_t cluster_generic(uint32_t ca) {
    uint64_t g;
    asm volatile("{\n\t.reg .u64 a;\n\tcvt.u64.u32 a, %1;\n\tcvta.shared::cluster.u64 %0, a;\n\t}" : "=l"(g) : "r"(ca));
    return g;
}
ISP_DEVINL void st_slot_if(uint64_t ga, float v, int tag, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 w, {%1, %2};\n\t"
                 "@p st.relaxed.cluster.b64 [%0], w;\n\t}"
                 ::"l"(ga), "r"(__float_as_uint(v)), "r"(tag), "r"(uint32_t(pred)) : "memory");
}
// two adjacent slots in one 16 B access (each 8 B half is its own flag, so a torn pair is harmless)
ISP_DEVINL void st_slot2_if(uint64_t ga, float v0, int tag0, float v1, int tag1, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 w0, w1;\n\tsetp.ne.u32 p, %5, 0;\n\tmov.b64 w0, {%1, %2};\n\tmov.b64 w1, {%3, %4};\n\t"
                 "@p st.relaxed.cluster.v2.b64 [%0], {w0, w1};\n\t}"
                 ::"l"(ga), "r"(__float_as_uint(v0)), "r"(tag0), "r"(__float_as_uint(v1)), "r"(tag1), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void ld_slot2(uint32_t sa, float& v0, int& tag0, float& v1, int& tag1) {
    uint32_t a, b, c, d;
    asm volatile("{\n\t.reg .b64 w0, w1;\n\tld.relaxed.cluster.shared::cta.v2.b64 {w0, w1}, [%4];\n\tmov.b64 {%0, %1}, w0;\n\tmov.b64 {%2, %3}, w1;\n\t}"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sa) : "memory");
    v0 = __uint_as_float(a); tag0 = int(b); v1 = __uint_as_float(c); tag1 = int(d);
}
ISP_DEVINL void st_generic_u16_if(uint64_t ga, int v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.u16 [%0], %1;\n\t}" ::"l"(ga), "h"(short(v)), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void ld_slot(uint32_t sa, float& v, int& tag) {
    uint32_t lo, hi;
    asm volatile("{\n\t.reg .b64 w;\n\tld.relaxed.cluster.shared::cta.b64 w, [%2];\n\tmov.b64 {%0, %1}, w;\n\t}"
                 : "=r"(lo), "=r"(hi) : "r"(sa) : "memory");
    v = __uint_as_float(lo);
    tag = int(hi);
}
ISP_DEVINL void sts_u64_if(uint32_t sa, uint32_t lo, uint32_t hi, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t}"
                 ::"r"(sa), "r"(lo), "r"(hi), "r"(uint32_t(pred)) : "memory");
}
ISP_DEVINL void sts_v4(uint32_t sa, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
ISP_DEVINL uint32_t lds_u32_(uint32_t sa) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
ISP_DEVINL float2 lds_f32x2_(uint32_t sa) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa));
    return v;
}
// bits 0..15 of e to the even positions, of o to the odd ones
ISP_DEVINL uint32_t spread16(uint32_t x) {
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
ISP_DEVINL uint32_t interleave16(uint32_t e, uint32_t o) { return spread16(e) | (spread16(o) << 1); }
ISP_DEVINL long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
ISP_DEVINL void sts_u16_(uint32_t sa, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa), "h"(short(v)) : "memory"); }
ISP_DEVINL int lds_u16_(uint32_t sa) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(sa) : "memory");
    return int(v);
}
ISP_DEVINL void spin_check(uint32_t& spins) { if (++spins > (1u << 25)) __trap(); }   // a protocol bug must not hang the GPU

// =================================== the strip warp: forward DP ===================================================
// Lane l owns the strip's columns 2l and 2l + 1.  A round advances TWO rows on one exchange: the lane takes its left
// neighbour's two values of row r - 1 (two rotating shuffles), recomputes the neighbour's cell Q[r][2l-1] itself (the halo:
// same operands, same fp32 operations, so the same bits), and then has everything rows r and r + 1 need.
// What bounds a strip is one warp's dependent chain and whatever shares the SM's memory-instruction queue with the two
// shuffles of a round, and the generated code is very sensitive to its schedule.  The one that ships was chosen by measurement
// among seven builds (cfg4 268 -> 217 us): a chunk's logits in ONE burst of loads at its top (no
// prefetch inside the rounds), the backpointer words stored per row by lane 0, the chunk's boundary values assembled as
// {value, tag} pairs in register quads of their own and stored in one burst after its rounds, 32 rows per chunk.
// HAS_PREV: a strip on the left feeds this one's column 0; HAS_NEXT: this strip feeds one on the right.
template <bool HAS_PREV, bool HAS_NEXT>
ISP_DEVINL void sweep(const Params& p, uint32_t smem_sa, uint32_t ctl, int s, int lane, int n, uint32_t rank, long long* tr) {
    long long w_land = 0, w_bnd = 0;      // trace: cycles spent waiting for logits / for the strip on the left
    const uint32_t ring = smem_sa + p.off_ring + uint32_t(s * kStrip + 2 * lane) * 4u;
    const uint32_t rmask = uint32_t(p.ring_rows - 1);
    const uint32_t bits_sa = smem_sa + p.off_bits + uint32_t(s) * 8u;                     // row r at + 16 r
    const uint32_t bnd_in = smem_sa + p.off_bnd + uint32_t(s) * uint32_t(kBnd * 8);
    // the next strip's boundary-in ring: strip 1 of this CTA, or strip 0 of the CTA on the right
    const uint64_t bnd_out = HAS_NEXT ? cluster_generic(mapa(smem_sa + p.off_bnd + uint32_t(s == 0 ? kBnd * 8 : 0), s == 0 ? rank : rank + 1)) : 0ull;
    const int src = (lane + 31) & 31;
    const bool lane0 = lane == 0, lane31 = lane == 31;
    auto put4 = [&](uint32_t off, const uint4& v) __attribute__((always_inline)) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.relaxed.cluster.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                     ::"l"(bnd_out + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(uint32_t(lane31)) : "memory");
    };
    // where this strip reports the rows it has consumed: the producer's kCtlCons word
    const uint32_t cons_out = HAS_PREV ? mapa(ctl + 4u * (kCtlCons + (s == 1 ? 0u : 1u)), s == 1 ? rank : rank - 1) : 0u;
    const uint32_t landed_sa = ctl + 4u * kCtlLanded, cons_in = ctl + 4u * (kCtlCons + uint32_t(s));
    const uint32_t emask = HAS_PREV ? 0xffffffffu : 0xfffffffeu;                         // global column 0 never moves (mas.py:16)
    float q0 = -CUDART_INF_F, q1 = -CUDART_INF_F;
    uint32_t ke = 0, ko = 0;        // (ISP_MASC_BITS_ROW == 0: lane k keeps the backpointer words of row k of the current chunk)

    // Q[r][j] = x[r][j] + max(Q[r-1][j-1], Q[r-1][j]), bit = Q[r-1][j-1] >= Q[r-1][j]   (mas.py:14, :17)
    // rows r = R + 2k and r + 1; bA = Q[r-1][-1], bB = Q[r][-1] of this strip (the strip on the left, or the virtual Q[-1][-1] = 0 / -inf)
    // (ISP_MASC_PF / ISP_MASC_BITS_ROW / ISP_MASC_PUT_END select the schedules that were measured against each other; the
    // defaults are what ships.  pf: address of the even row of the next chunk's logits for this round, 0 = none -- always 0 with
    // ISP_MASC_PF == 2.)
    uint4 bq[kCh / 2];
    auto step2 = [&](int R, int k, float2 xa, float2 xb, float ha, float bA, float bB, uint32_t slot, uint32_t pf, float2& na, float2& nb,
                     float& nh) __attribute__((always_inline)) {
#if ISP_MASC_PF == 0
        if (pf) { na = lds_f32x2_(pf); nb = lds_f32x2_(pf + kRowBytes); nh = lds_f32(pf - 4u); }
#endif
        const float t1 = __shfl_sync(0xffffffffu, q1, src);
        const float t0 = __shfl_sync(0xffffffffu, q0, src);

        const float L1 = lane0 ? bA : t1;                    // Q[r-1][2l-1]
        const float hh = __fadd_rn(ha, fmaxf(t0, L1));       // Q[r][2l-1], as its owner computes it
        const float H = lane0 ? bB : hh;
        const bool a0 = L1 >= q0, a1 = q0 >= q1;
        const float n0 = __fadd_rn(xa.x, fmaxf(L1, q0));
        const float n1 = __fadd_rn(xa.y, fmaxf(q0, q1));
        const bool b0 = H >= n0, b1 = n0 >= n1;
        q0 = __fadd_rn(xb.x, fmaxf(H, n0));
        q1 = __fadd_rn(xb.y, fmaxf(n0, n1));
        const uint32_t ea = __ballot_sync(0xffffffffu, a0) & emask, oa = __ballot_sync(0xffffffffu, a1);
        const uint32_t eb = __ballot_sync(0xffffffffu, b0) & emask, ob = __ballot_sync(0xffffffffu, b1);
#if ISP_MASC_BITS_ROW
        sts_u64_if(bits_sa + uint32_t(R + 2 * k) * 16u, ea, oa, lane0);
        sts_u64_if(bits_sa + uint32_t(R + 2 * k + 1) * 16u, eb, ob, lane0);
#else
        const bool mine_a = lane == 2 * k, mine_b = lane == 2 * k + 1;
        ke = mine_a ? ea : ke; ko = mine_a ? oa : ko;
        ke = mine_b ? eb : ke; ko = mine_b ? ob : ko;
#endif
#if ISP_MASC_PF == 1
        if (pf) { na = lds_f32x2_(pf); nb = lds_f32x2_(pf + kRowBytes); nh = lds_f32(pf - 4u); }      // after the round's arithmetic
#endif
#if ISP_MASC_PUT_END
        if (HAS_NEXT) bq[k] = make_uint4(__float_as_uint(n1), uint32_t(R + 2 * k + 1), __float_as_uint(q1), uint32_t(R + 2 * k + 2));
#else
        if (HAS_NEXT) st_slot2_if(bnd_out + slot, n1, R + 2 * k + 1, q1, R + 2 * k + 2, lane31);
#endif
    };
    auto step1 = [&](int r, float2 xa, float bA, uint64_t slot) __attribute__((always_inline)) {
        const float t1 = __shfl_sync(0xffffffffu, q1, src);
        const float L1 = lane0 ? bA : t1;
        const bool a0 = L1 >= q0, a1 = q0 >= q1;
        const float n0 = __fadd_rn(xa.x, fmaxf(L1, q0));
        q1 = __fadd_rn(xa.y, fmaxf(q0, q1));
        q0 = n0;
        const uint32_t ea = __ballot_sync(0xffffffffu, a0) & emask, oa = __ballot_sync(0xffffffffu, a1);
        sts_u64_if(bits_sa + uint32_t(r) * 16u, ea, oa, lane0);
        if (HAS_NEXT) st_slot_if(slot, q1, r + 1, lane31);
    };
    auto wait_landed = [&](int need) __attribute__((always_inline)) {
        if (ld_volatile_sa(landed_sa) < need) {
            uint32_t spins = 0;
            const long long t0 = tr ? clock64() : 0;
            while (ld_volatile_sa(landed_sa) < need) { __nanosleep(20); spin_check(spins); }
            if (tr) w_land += clock64() - t0;
        }
    };
    // one chunk of (up to) kCh rows from R: xc / hc hold its logits if `have`; xn / hn receive the next chunk's (ISP_MASC_PF != 2)
    auto chunk = [&](int R, float2 (&xc)[kCh], float (&hc)[kCh / 2], bool have, float2 (&xn)[kCh], float (&hn)[kCh / 2]) __attribute__((always_inline)) -> bool {
        const int rows = min(kCh, n - R);
        const uint32_t xa = ring + (uint32_t(R) & rmask) * kRowBytes;
        const uint64_t slot0 = bnd_out + uint64_t(R & (kBnd - 1)) * 8u;           // (R is a multiple of 16: a chunk never wraps)
        bool have_next = false;

        if (HAS_NEXT) {
            // this chunk overwrites the slots of rows R - kBnd ..: the consumer must be past them
            const int need = R + rows + 1 - kBnd;
            uint32_t spins = 0;
            if (need > 0 && p.dbg != 1) while (ld_relaxed_cluster_sa(cons_in) < need) { __nanosleep(20); spin_check(spins); }
        }
        if (rows == kCh) {
            if (!have) {
                wait_landed(R + kCh);
#pragma unroll
                for (int k = 0; k < kCh; ++k) xc[k] = lds_f32x2_(xa + uint32_t(k) * kRowBytes);
#pragma unroll
                for (int k = 0; k < kCh / 2; ++k) hc[k] = lds_f32(xa + uint32_t(2 * k) * kRowBytes - 4u);
            }
            // the next chunk's logits, if they are there already, are loaded round by round under this chunk's steps
#if ISP_MASC_PF == 2
            have_next = false;                                   // no prefetch: every chunk loads its logits in one burst at its top
#else
            have_next = n - (R + kCh) >= kCh && ld_volatile_sa(landed_sa) >= R + 2 * kCh;
#endif
            const uint32_t xb = ring + (uint32_t(R + kCh) & rmask) * kRowBytes;
            float bv[kCh];
            if (HAS_PREV) {
                // Q[r-1][last column of the strip on the left] for r = R .. R + 15: slots R - 1 .. R + 14, read as the 16 B pairs the
                // producer wrote (slots 2i, 2i + 1)
                uint32_t spins = 0;
                const uint32_t base = uint32_t((R - 2) & (kBnd - 1)) * 8u;
                for (;;) {
                    bool ok = true;
                    float v0, v1; int g0, g1;
                    ld_slot2(bnd_in + base, v0, g0, v1, g1);                                   // slots R - 2, R - 1
                    bv[0] = R == 0 ? -CUDART_INF_F : v1;
                    ok = R == 0 || g1 == R;
#pragma unroll
                    for (int i = 0; i < kCh / 2; ++i) {
                        ld_slot2(bnd_in + uint32_t((R + 2 * i) & (kBnd - 1)) * 8u, v0, g0, v1, g1);   // slots R + 2i, R + 2i + 1
                        bv[2 * i + 1] = v0; ok = ok && g0 == R + 2 * i + 1;
                        if (2 * i + 2 < kCh) { bv[2 * i + 2] = v1; ok = ok && g1 == R + 2 * i + 2; }
                    }
                    if (ok) break;
                    const long long t0 = tr ? clock64() : 0;
                    __nanosleep(ISP_MASC_BNDSLEEP); spin_check(spins);
                    if (tr) w_bnd += clock64() - t0;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kCh; ++k) bv[k] = (R + k == 0) ? 0.0f : -CUDART_INF_F;     // Q[-1][-1] = 0 stands in for mas.py:11-12
            }
#pragma unroll
            for (int k = 0; k < kCh / 2; ++k)
                step2(R, k, xc[2 * k], xc[2 * k + 1], hc[k], bv[2 * k], bv[2 * k + 1], uint32_t(R & (kBnd - 1)) * 8u + uint32_t(2 * k) * 8u,
                      have_next ? xb + uint32_t(2 * k) * kRowBytes : 0u, xn[2 * k], xn[2 * k + 1], hn[k]);

#if ISP_MASC_PUT_END
            if (HAS_NEXT) {
#pragma unroll
                for (int k = 0; k < kCh / 2; ++k) put4(uint32_t(R & (kBnd - 1)) * 8u + uint32_t(2 * k) * 8u, bq[k]);
            }
#endif
#if !ISP_MASC_BITS_ROW
            sts_u64_if(bits_sa + uint32_t(R + (lane & (kCh - 1))) * 16u, ke, ko, lane < kCh);     // raw: even columns, odd columns (the mapper interleaves)
#endif
        } else {
            wait_landed(R + rows);
            for (int k = 0; k < rows; ++k) {
                const int r = R + k;
                float bv = (!HAS_PREV && r == 0) ? 0.0f : -CUDART_INF_F;
                if (HAS_PREV && r > 0) {
                    int tag;
                    uint32_t spins = 0;
                    for (;;) { ld_slot(bnd_in + uint32_t((r - 1) & (kBnd - 1)) * 8u, bv, tag); if (tag == r) break; __nanosleep(20); spin_check(spins); }
                }
                step1(r, lds_f32x2_(xa + uint32_t(k) * kRowBytes), bv, slot0 + uint64_t(k) * 8u);
            }
        }
        __syncwarp();
        // the stores of the chunk's bits precede this one in program order in every lane (shared memory keeps a thread's stores in
        // order, and the warp has converged)
        st_volatile_if_sa(ctl + 4u * (kCtlProg + uint32_t(s)), R + rows, lane0);
        if (HAS_PREV && lane0) st_relaxed_cluster(cons_out, R + rows);
        __syncwarp();
        return have_next;
    };

    float2 xA[kCh], xB[kCh];        // two chunks of logits (this lane's two columns), used in turn
    float hA[kCh / 2], hB[kCh / 2]; // and the left neighbour's column on the even rows
    bool have = false;
    for (int R = 0; R < n; R += 2 * kCh) {
        have = chunk(R, xA, hA, have, xB, hB);
        if (R + kCh >= n) break;
        have = chunk(R + kCh, xB, hB, have, xA, hA);
    }
    if (tr && lane0) { tr[2 + s] = gtimer(); tr[10 + 2 * s] = w_land; tr[11 + 2 * s] = w_bnd; }
}

// =================================== the kernel ======================================================================
__global__ void __launch_bounds__(kThreads, 1)
mas_cluster_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem_sa = smem_u32(smem_raw);
    const uint32_t ctl = smem_sa;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t rank = cluster_ctarank();
    const int nc = p.nc;
    const int b = blockIdx.x / nc;

    const long long n64 = p.mel_len[b], m64 = p.text_len[b];
    if (rank == 0 && tid == 0 && (n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max)) atomicAdd(p.status, 1);
    const int n = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));
    const int m = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));
    const int nblk = (n + kBlk - 1) / kBlk;
    const int col0 = int(rank) * kColsCta;                         // first global column of this CTA
    const bool act0 = col0 < m, act1 = col0 + kStrip < m;          // which of the two strips carry valid tokens
    const int stages = p.ring_rows / kCh;
    long long* tr = p.trace ? p.trace + size_t(blockIdx.x) * 32 : nullptr;
    if (tr && tid == 0) tr[0] = gtimer();

    // ---- set-up: control words, boundary tags, barriers ----
    for (uint32_t i = tid; i < (kCtlBytes + 2u * kBnd * 8u) / 4u; i += kThreads) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    for (uint32_t i = tid; i < uint32_t(nblk) * kPlanes * 2u; i += kThreads) reinterpret_cast<uint32_t*>(smem_raw + p.off_carry)[i] = 0u;
    __syncthreads();
    if (tid == 0) {
        for (int st = 0; st < stages; ++st) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 4 * kCtlFull) + st, 1);
        fence_mbar_init();
    }
    cluster_sync();                                                // nobody writes into a neighbour that is not set up yet
    if (tr && tid == 0) tr[1] = gtimer();

    const int rows_per = (p.T1max + nc - 1) / nc;                  // this CTA's slice of the dense output
    const int r_lo = min(p.T1max, int(rank) * rows_per), r_hi = min(p.T1max, r_lo + rows_per);

    if (warp < 2) {
        // ------------------------------------------ strips ------------------------------------------
        const int s = warp;
        const int gcol = col0 + s * kStrip;
        if (gcol < m) {
            const bool has_prev = gcol > 0, has_next = gcol + kStrip < m || (p.dbg == 1 && s == 0);
            if (has_prev) { if (has_next) sweep<true, true>(p, smem_sa, ctl, s, lane, n, rank, tr); else sweep<true, false>(p, smem_sa, ctl, s, lane, n, rank, tr); }
            else          { if (has_next) sweep<false, true>(p, smem_sa, ctl, s, lane, n, rank, tr); else sweep<false, false>(p, smem_sa, ctl, s, lane, n, rank, tr); }
        }
    } else if (warp == 2) {
        // ------------------------------------------ loader ------------------------------------------
        if (act0) {
            const int ncols = min(kColsCta, p.T2max - col0);       // columns of this CTA that exist in the tensor
            const float* src = p.logp + size_t(b) * p.sB + col0;
            const uint32_t ring = smem_sa + p.off_ring;
            const uint32_t rmask = uint32_t(p.ring_rows - 1);
            const uint32_t prog0 = ctl + 4u * kCtlProg, prog1 = prog0 + 4u;
            const int nch = (n + kCh - 1) / kCh;
            auto consumed = [&]() __attribute__((always_inline)) {
                int c = ld_volatile_sa(prog0);
                if (act1) c = min(c, ld_volatile_sa(prog1));
                return c;
            };
            if (p.bulk) {
                uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 4 * kCtlFull);
                const uint64_t policy = policy_evict_first();
                // issue and completion are decoupled: a loader that blocks on the oldest chunk's barrier cannot refill the slots the
                // strips free meanwhile, and the ring drains to one chunk per TMA round trip
                int ci = 0, cc = 0;                                // next chunk to issue / to complete
                uint32_t spins = 0;
                while (cc < nch) {
                    bool did = false;
                    while (ci < nch && ci < cc + stages && consumed() >= (ci + 1) * kCh - p.ring_rows) {
                        const int R = ci * kCh;
                        uint64_t* bar = full + (ci % stages);
                        // one box of kCh rows x 128 columns (columns past T2max and rows past T1max arrive as zeros): per-row bulk
                        // copies cost the TMA unit ~50 ns each, more than a strip takes to consume the row
                        if (lane == 0) {
                            mbar_arrive_expect_tx(bar, uint32_t(kCh) * kRowBytes);
                            tma_load_box(ring + (uint32_t(R) & rmask) * kRowBytes, &tmap, col0, R, b, smem_u32(bar), policy);
                        }
                        ++ci;
                        did = true;
                    }
                    __syncwarp();
                    while (cc < ci && mbar_test_sa(smem_u32(full + (cc % stages)), uint32_t(cc / stages) & 1u)) {
                        ++cc;
                        did = true;
                        if (lane == 0) st_volatile_sa(ctl + 4u * kCtlLanded, min(n, cc * kCh));
                    }
                    __syncwarp();
                    if (!did) { __nanosleep(ISP_MASC_LDSLEEP); spin_check(spins); }
                }
            } else {
                // rows that are not 16 B aligned (strided or odd T2max): through registers, one chunk at a time
                for (int cc = 0; cc < nch; ++cc) {
                    const int R = cc * kCh, rows = min(kCh, n - R);
                    uint32_t spins = 0;
                    while (consumed() < R + rows - p.ring_rows) { __nanosleep(40); spin_check(spins); }
                    for (int k = 0; k < rows; ++k) {
                        const float* xr = src + size_t(R + k) * p.sT1;
                        const uint32_t dst = ring + (uint32_t(R + k) & rmask) * kRowBytes;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = lane + 32 * q;
                            if (c < ncols) sts_f32(dst + uint32_t(c) * 4u, __ldcs(xr + c));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) st_volatile_sa(ctl + 4u * kCtlLanded, R + rows);
                    __syncwarp();
                }
            }
        }
        if (tr && lane == 0) tr[9] = gtimer();
    } else if (warp == 3) {
        // ------------------------------------------ mapper, then the hop chain ------------------------------------------
        const int pl = min(lane, kPlanes - 1);
        const uint32_t maps_sa = smem_sa + p.off_maps, carry_sa = smem_sa + p.off_carry, bits_sa = smem_sa + p.off_bits;
        if (act0) {
            const bool feed_next = int(rank) + 1 < nc && col0 + kColsCta < m;
            // the carry words travel as {value, block tag} in one 64-bit relaxed store each: like the strips' boundary slots they are
            // their own flags (a release / acquire pair per block costs the mapper a fence, ~1 us, and it fell behind the sweep)
            const uint64_t carry_next = feed_next ? cluster_generic(mapa(carry_sa, rank + 1)) : 0ull;
            const uint32_t m23 = act1 ? 0xffffffffu : 0u;          // an inactive strip's words were never written
            // identity, low 6 bits of the column: bit c of word w is bit `pl` of 32 w + c
            const uint32_t idp = pl == 0 ? 0xaaaaaaaau : pl == 1 ? 0xccccccccu : pl == 2 ? 0xf0f0f0f0u : pl == 3 ? 0xff00ff00u : 0xffff0000u;
            uint32_t idw[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) idw[w] = pl < 5 ? idp : ((w & 1) ? 0xffffffffu : 0u);
            for (int blk = 0; blk < nblk; ++blk) {
                const int rows_end = min(n, kBlk * (blk + 1));
                uint32_t spins = 0;
                while (ld_volatile_sa(ctl + 4u * kCtlProg) < rows_end || (act1 && ld_volatile_sa(ctl + 4u * (kCtlProg + 1)) < rows_end)) { __nanosleep(ISP_MASC_MAPPOLL); spin_check(spins); }
                uint32_t cin = 0;
                if (rank > 0) {
                    for (;;) {
                        float cv; int tag;
                        ld_slot(carry_sa + uint32_t(blk * kPlanes + pl) * 8u, cv, tag);
                        cin = __float_as_uint(cv);
                        if (__all_sync(0xffffffffu, tag == blk + 1)) break;
                        __nanosleep(100); spin_check(spins);
                    }
                }
                {
                    // the strips leave a row as {even columns, odd columns} per strip: interleave into column order, one row per lane
                    const int r = kBlk * blk + lane;
                    if (r < n) {
                        const uint4 raw = lds_v4(bits_sa + uint32_t(r) * 16u);
                        sts_v4(bits_sa + uint32_t(r) * 16u, make_uint4(interleave16(raw.x, raw.y), interleave16(raw.x >> 16, raw.y >> 16),
                                                                        interleave16(raw.z, raw.w), interleave16(raw.z >> 16, raw.w >> 16)));
                    }
                    __syncwarp();
                }
                uint32_t W0 = idw[0], W1 = idw[1], W2 = idw[2], W3 = idw[3], cout = 0;
                for (int t0 = 0; t0 < kBlk; t0 += 8) {             // eight rows at a time: the loads first, so that their latencies overlap
                    uint4 Bq[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) Bq[u] = lds_v4(bits_sa + uint32_t(kBlk * blk + t0 + u) * 16u);   // (rows >= n: stale words, masked below)
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = t0 + u, r = kBlk * blk + t;
                        const uint32_t live = (r >= 1 && r < n) ? 0xffffffffu : 0u;     // row 0 has no predecessor, rows >= n do not exist
                        const uint32_t Bx = Bq[u].x & live, By = Bq[u].y & live, Bz = Bq[u].z & live & m23, Bw = Bq[u].w & live & m23;
                        cout |= (W3 >> 31) << t;                   // the map's value at this CTA's last column BEFORE this row
                        const uint32_t n0 = (W0 << 1) | ((cin >> t) & 1u);
                        const uint32_t n1 = __funnelshift_l(W0, W1, 1), n2 = __funnelshift_l(W1, W2, 1), n3 = __funnelshift_l(W2, W3, 1);
                        W0 = (Bx & n0) | (~Bx & W0);
                        W1 = (By & n1) | (~By & W1);
                        W2 = (Bz & n2) | (~Bz & W2);
                        W3 = (Bw & n3) | (~Bw & W3);
                    }
                }
                if (lane < kPlanes) sts_v4(maps_sa + uint32_t(blk * kPlanes + lane) * 16u, make_uint4(W0, W1, W2, W3));
                if (feed_next) st_slot_if(carry_next + uint64_t(blk * kPlanes + pl) * 8u, __uint_as_float(cout), blk + 1, lane < kPlanes);
                __syncwarp();
            }
        }
        if (tr && lane == 0) tr[4] = gtimer();
        // ---- hops: one lookup per block, from the last block down; the chain moves to the CTA on the left with the path ----
        const int rank_top = (m - 1) / kColsCta;
        int blk = -1, j = 0;
        bool go = false;
        if (int(rank) == rank_top) { blk = nblk - 1; j = m - 1; go = true; }
        else {
            uint32_t spins = 0;
            int f;
            while ((f = ld_acquire_cluster_sa(ctl + 4u * kCtlHopFlag)) == 0) { __nanosleep(200); spin_check(spins); }
            if (f == 1) { blk = ld_volatile_sa(ctl + 4u * kCtlHopBlk); j = ld_volatile_sa(ctl + 4u * kCtlHopJ); go = true; }
        }
        if (go) {
            // the entries (the path's column at each block's last row) go to THIS CTA's array inside the loop -- a store into another
            // CTA followed by a shared load costs the chain ~200 cycles per hop -- and to CTA 0's, lanes in parallel, after it
            const uint32_t ent_sa = smem_sa + p.off_ent;
            const int blk_first = blk;
            while (blk >= 0 && j >= col0) {
                if (lane == 0) sts_u16_(ent_sa + uint32_t(blk) * 2u, j);
                const int cl = j - col0;
                const uint32_t wv = lds_u32_(maps_sa + uint32_t(blk * kPlanes + pl) * 16u + uint32_t(cl >> 5) * 4u);
                const uint32_t low6 = __ballot_sync(0xffffffffu, ((wv >> (cl & 31)) & 1u) != 0u) & 63u;
                j -= int((uint32_t(j) - low6) & 63u);                                    // the target is within 32 columns below j
                --blk;
            }
            if (blk < 0) {
                if (lane < nc) st_release_cluster(mapa(ctl + 4u * kCtlHopFlag, uint32_t(lane)), 2);
            } else if (lane == 0) {
                st_cluster_u32(mapa(ctl + 4u * kCtlHopBlk, rank - 1), uint32_t(blk));
                st_cluster_u32(mapa(ctl + 4u * kCtlHopJ, rank - 1), uint32_t(j));
                st_release_cluster(mapa(ctl + 4u * kCtlHopFlag, rank - 1), 1);
            }
            __syncwarp();
            if (rank != 0) {
                const uint32_t ent0 = mapa(ent_sa, 0);
                for (int bb = blk_first - lane; bb > blk; bb -= 32) st_cluster_u16(ent0 + uint32_t(bb) * 2u, lds_u16_(ent_sa + uint32_t(bb) * 2u));
            }
            __syncwarp();
        }
        if (tr && lane == 0) tr[5] = gtimer();
    } else if (p.hard != nullptr) {
        // ------------------------------------------ zero fill of this CTA's rows of the dense output ------------------------------------------
        const int ft = tid - 4 * 32, nft = kThreads - 4 * 32;
        int16_t* base = p.hard + (size_t(b) * p.T1max + r_lo) * p.T2max;
        const size_t total = size_t(r_hi - r_lo) * p.T2max;
        size_t head = (16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15;
        head >>= 1; if (head > total) head = total;
        for (size_t i = ft; i < head; i += nft) base[i] = 0;
        const size_t n16 = (total - head) >> 3;
        uint4* v = reinterpret_cast<uint4*>(base + head);
        // paced: an unpaced fill of every cluster at once saturates HBM for its duration and starves the sweeps' loaders
        const unsigned pace = n > 512 ? unsigned(ISP_MASC_FILLPACE) : 0u;
        for (size_t i = ft; i < n16; i += nft) { st_cs_v4(v + i, make_uint4(0u, 0u, 0u, 0u)); if (pace) __nanosleep(pace); }
        for (size_t i = head + (n16 << 3) + ft; i < total; i += nft) base[i] = 0;
        if (tr && ft == 0) tr[14] = gtimer();
    }
    __syncwarp();
    cluster_sync();        // every block's entry column is in CTA 0; every CTA's bits are complete; the zero fills are issued
    if (tr && tid == 0) tr[6] = gtimer();

    // ---- walk: one thread per block of 32 rows, bits from whichever CTA owns the column ----
    int16_t* path = p.path + size_t(b) * p.T1max;
    {
        const uint32_t ent0 = mapa(smem_sa + p.off_ent, 0);
        const uint32_t bits_sa = smem_sa + p.off_bits;
        // a block is walked by the CTA that owns its entry column: its bits are local until the path crosses to the left
        for (int blk = tid; blk < nblk; blk += kThreads) {
            int j = ld_cluster_u16(ent0 + uint32_t(blk) * 2u);
            if (j / kColsCta != int(rank)) continue;
            for (int t = kBlk - 1; t >= 0; --t) {
                const int r = kBlk * blk + t;
                if (r >= n) continue;
                path[r] = int16_t(j);
                if (r >= 1) {
                    const int cl = j & (kColsCta - 1);
                    const uint32_t wa = bits_sa + uint32_t(r) * 16u + uint32_t(cl >> 5) * 4u;
                    const uint32_t wv = j >= col0 ? lds_u32_(wa) : ld_cluster_u32(mapa(wa, uint32_t(j / kColsCta)));
                    j -= int((wv >> (cl & 31)) & 1u);
                }
            }
        }
    }
    cluster_sync();        // the whole path is in global memory; nobody reads a neighbour's shared memory after this
    if (tr && tid == 0) tr[7] = gtimer();

    // ---- outputs: the ones of this CTA's rows, -1 past the utterance, the durations of this CTA's tokens ----
    if (p.hard != nullptr)
        for (int r = r_lo + tid; r < min(r_hi, n); r += kThreads) p.hard[(size_t(b) * p.T1max + r) * p.T2max + path[r]] = 1;
    if (p.path_is_output)
        for (int r = max(r_lo, n) + tid; r < r_hi; r += kThreads) path[r] = -1;
    if (p.dur != nullptr && col0 < p.T2max) {
        int* first = reinterpret_cast<int*>(smem_raw + p.off_ring);     // the ring is free now
        int* last = first + kColsCta;
        if (tid < kColsCta) { first[tid] = 0; last[tid] = 0; }
        __syncthreads();
        // the path is monotone: a token's run is [first frame with it, last frame with it]
        for (int r = tid; r < n; r += kThreads) {
            const int jj = path[r], jl = jj - col0;
            if (jl >= 0 && jl < kColsCta) {
                if (r == 0 || path[r - 1] != jj) first[jl] = r;
                if (r == n - 1 || path[r + 1] != jj) last[jl] = r + 1;
            }
        }
        __syncthreads();
        if (tid < kColsCta && col0 + tid < p.T2max) p.dur[size_t(b) * p.T2max + col0 + tid] = int64_t(last[tid] - first[tid]);
    }
    if (tr && tid == 0) tr[8] = gtimer();
}

// ---------------------------------------------------------------------------------------------------------------------
struct Layout { uint32_t off_bnd, off_ent, off_carry, off_maps, off_bits, off_ring, total; int ring_rows; };

static bool make_layout(int T1max, Layout* L) {
    const uint32_t nblk = uint32_t(T1max + kBlk - 1) / kBlk;
    uint32_t off = kCtlBytes;
    L->off_bnd = off;   off += 2u * kBnd * 8u;
    L->off_ent = off;   off += (nblk * 2u + 15u) & ~15u;
    L->off_carry = off; off += (nblk * kPlanes * 8u + 15u) & ~15u;
    L->off_maps = off;  off += nblk * kPlanes * 16u;
    L->off_bits = off;  off += uint32_t(T1max) * 16u;
    off = (off + 127u) & ~127u;
    L->off_ring = off;
    for (int rows = kMaxStages * kCh; rows >= 2 * kCh; rows >>= 1) {
        if (off + uint32_t(rows) * kRowBytes <= 227u * 1024u) { L->ring_rows = rows; L->total = off + uint32_t(rows) * kRowBytes; return true; }
    }
    return false;
}

}  // namespace masc

bool mas_cluster_supported(int B, int T1max, int T2max) {
    if (B <= 0 || T1max <= 0 || T2max <= 0 || T2max > masc::kColsCta * masc::kMaxCluster) return false;
    masc::Layout L;
    return masc::make_layout(T1max, &L);
}

static int g_opt_trace = 0, g_opt_dbg = 0;
int mas_cluster_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "masc.trace")) { *prev = g_opt_trace; g_opt_trace = value; return 0; }
    if (!strcmp(key, "masc.dbg")) { *prev = g_opt_dbg; g_opt_dbg = value; return 0; }
    return -1;
}

// status word, the path's scratch, the debug trace (16 words per CTA)
size_t mas_cluster_workspace_bytes(int B, int T1max, int) { return 256 + ((size_t(B) * T1max * 2 + 15) & ~size_t(15)) + size_t(B) * masc::kMaxCluster * 32 * 8; }

// Do all B clusters of this shape run at once?  (A second wave doubles the kernel: the dispatcher then prefers the single-CTA
// kernel.)  cudaOccupancyMaxActiveClusters knows the GPC packing; the answer is cached per (cluster size, shared memory).
bool mas_cluster_one_wave(int B, int T1max, int T2max) {
    masc::Layout L;
    if (!mas_cluster_supported(B, T1max, T2max) || !masc::make_layout(T1max, &L)) return false;
    const int nc = (T2max + masc::kColsCta - 1) / masc::kColsCta;
    static thread_local struct { int nc; uint32_t smem; int dev; int clusters; } memo[8] = {};
    static thread_local int used = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    for (int i = 0; i < used; ++i)
        if (memo[i].nc == nc && memo[i].smem == L.total && memo[i].dev == dev) return B <= memo[i].clusters;
    if (cudaFuncSetAttribute(masc::mas_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total)) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(nc) * 148u, 1, 1);
    cfg.blockDim = dim3(masc::kThreads, 1, 1);
    cfg.dynamicSmemBytes = L.total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(nc);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, masc::mas_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    if (used < 8) { memo[used].nc = nc; memo[used].smem = L.total; memo[used].dev = dev; memo[used].clusters = clusters; ++used; }
    return B <= clusters;
}

typedef CUresult (*PFN_encodeTiledC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (T2max, T1max, B) fp32, box 128 columns x kCh rows: what one CTA of a cluster takes per chunk
static bool make_cluster_map(CUtensorMap* map, const float* logp, int64_t sB, int64_t sT1, int B, int T1max, int T2max) {
    static PFN_encodeTiledC enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return false;
        enc = reinterpret_cast<PFN_encodeTiledC>(ptr);
    }
    if ((reinterpret_cast<uintptr_t>(logp) & 15) || (sT1 & 3) || (sB & 3)) return false;
    cuuint64_t dims[3] = {cuuint64_t(T2max), cuuint64_t(T1max), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(sT1) * 4, cuuint64_t(B > 1 ? sB : sT1 * T1max) * 4};
    cuuint32_t box[3] = {cuuint32_t(masc::kColsCta), cuuint32_t(masc::kCh), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(logp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int mas_cluster_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len, int B, int T1max,
                        int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, cudaStream_t stream) {
    masc::Layout L;
    if (!mas_cluster_supported(B, T1max, T2max) || !masc::make_layout(T1max, &L)) {
        set_error("isp_mas_forward: the cluster kernel does not cover T1max=%d T2max=%d", T1max, T2max);
        return ISP_ERR_UNSUPPORTED;
    }
    masc::Params p;
    p.logp = logp; p.sB = sB; p.sT1 = sT1;
    p.text_len = text_len; p.mel_len = mel_len;
    p.B = B; p.T1max = T1max; p.T2max = T2max;
    p.hard = attn_hard; p.dur = durations;
    p.status = reinterpret_cast<int*>(ws);
    p.path = path ? path : reinterpret_cast<int16_t*>(reinterpret_cast<char*>(ws) + 256);
    p.path_is_output = path ? 1 : 0;
    p.nc = (T2max + masc::kColsCta - 1) / masc::kColsCta;
    p.ring_rows = L.ring_rows;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    p.bulk = make_cluster_map(&tmap, logp, sB, sT1, B, T1max, T2max) ? 1 : 0;     // else (rows not 16 B aligned) the loader copies through registers
    p.dbg = g_opt_dbg;
    p.trace = g_opt_trace ? reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + 256 + ((size_t(B) * T1max * 2 + 15) & ~size_t(15))) : nullptr;
    p.off_bnd = L.off_bnd; p.off_ent = L.off_ent; p.off_carry = L.off_carry; p.off_maps = L.off_maps; p.off_bits = L.off_bits; p.off_ring = L.off_ring;

    cudaError_t e = cudaMemsetAsync(ws, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");
    e = cudaFuncSetAttribute(masc::mas_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas_cluster_kernel)");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(B) * unsigned(p.nc), 1, 1);
    cfg.blockDim = dim3(masc::kThreads, 1, 1);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(p.nc);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, masc::mas_cluster_kernel, tmap, p);
    if (e != cudaSuccess) return cuda_fail(e, "mas_cluster_kernel launch");
    return 0;
}

}  // namespace isp
