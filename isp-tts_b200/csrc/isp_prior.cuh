// The diagonal prior of batch_diagonal_prior (tts/models/acoustic/modules/alignment.py:18-37) as the kernels evaluate it: one
// definition for the forward kernel (isp_loglik.cu) and for the backward kernel that re-derives the prior's cells from the row
// sums the forward saved (isp_loglik_bwd.cu).  The 1e-4 threshold is a comparison on these values, so both sides must compute
// them with the same instructions in the same order:
//     g_j = fdiv_rn(j, T2_b) * kPriorScale          u_i = fdiv_rn(i, T1_b) * kPriorScale          d = g_j - u_i
//     raw = ex2.approx(-d * d)                      ( = exp(-(j / T2_b - i / T1_b)^2 / (2 * 0.1^2)) )
//     pr  = raw * (1 / (row sum of raw + 1e-5));    pr = pr >= 1e-4 ? pr : 0;      P = pr + 1e-6
#pragma once

#include "common.cuh"

namespace isp {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLogPriorFloor = -13.815510557964274f;   // log(1e-6)
constexpr float kPriorEps = 1e-6f;
constexpr float kPriorThreshold = 1e-4f;                 // alignment.py:18
constexpr float kNegInvTwoGammaSq = -50.0f;              // -1 / (2 * 0.1^2)
constexpr float kPriorScale = 8.493218002880191f;        // sqrt(50 log2 e): exp(-50 x^2) = 2^(-(kPriorScale x)^2)

ISP_DEVINL float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
ISP_DEVINL float prior_grid(int idx, float len) { return __fdiv_rn(float(idx), len) * kPriorScale; }
// P = thresholded prior + 1e-6 of one cell (ok: the cell is a valid token of a valid frame)
ISP_DEVINL float prior_cell_P(float g, float u, float inv_psum, bool ok) {
    const float d = g - u;
    float pr = fast_ex2(-d * d) * inv_psum;
    pr = (ok && pr >= kPriorThreshold) ? pr : 0.0f;
    return pr + kPriorEps;
}

}  // namespace isp
