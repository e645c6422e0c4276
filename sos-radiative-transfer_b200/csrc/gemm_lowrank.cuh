// Source contraction for LOW-RANK operands: J = coef * (I . Us) . Vt with A = Us Vt, rank r <= 16.
//
// The azimuth-averaged Rayleigh phase matrix of the reference (SOS_Aer_phase_func.py:79-133) is
//   P_raw[m, n] = a + b mu_m^2 mu_n^2 + c (1 - mu_m^2)(1 - mu_n^2)      (the cos(phi) cross term cancels between the half rings)
// followed by a per-column scale (:131): rank 2 -- and the isotropic one (:68-76) rank 1 -- so the contraction operand
// A[k][m] = w_k/4 P[m][N-1-k] of SOS_Aer_I1_In.py:73 is rank 2 (1) too, exactly up to rounding (singular values
// 1, 0.100, < 1e-16 at M = 501).  Every row outside the aerosol layer uses the molecular (Rayleigh) operand
// (SOS_Aer_main_specular.py:323): 746 of the 800 default rows.  For those rows the N x N contraction collapses to two
// skinny products, 2 r N multiply-adds per row instead of N^2: the rows become HBM bound (read I once, write J once)
// and the dense DMMA kernel is left with the aerosol rows.  Agreement with the dense kernels ~3e-15 relative.
//
// The factors come from sos_build_lowrank_mu2 (closed form, below; the host enables them when the fitted operand
// reproduces the dense one to rounding) and are registered with sos_plan_set_lowrank; groups of such operands get
// class 3 in the fold-mode tile plan (no dense tiles) and are processed here: one warp per LR_ROWS rows of an 8-row
// segment, fragments in registers.  Inside sos_solve the fused order kernel (strip.cuh) goes further and never
// materialises J for those rows.
#pragma once
#include "gemm_f64.cuh"

namespace sosgemm {

struct LowRankParams {
  const double* I;
  double* J;
  const double* Ut[SOS_MAX_PHASE];  // [RP][ldr]: (U diag(s))^T, zero rows beyond the rank
  const double* Vt[SOS_MAX_PHASE];  // [RP][ldr]
  const TilePlan* plan;
  const int* active_list;
  const int* seg_row0;    // class-0 segments: first row, valid rows
  const int* seg_valid0;
  int nseg0;
  int L, N, ld, ldr;
  const sos_scenario* scen;
};

constexpr int LR_ROWS = 4;      // rows per warp (2 rows per warp and twice the warps was measured: 0.643 vs 0.615 ms at S = 96)
constexpr int LR_THREADS = 256;

template <int RP>
__global__ void __launch_bounds__(LR_THREADS) jn_lowrank_kernel(const LowRankParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TilePlan* plan = p.plan;
  constexpr int PER_SEG = SEG_ROWS / LR_ROWS;
  const int per_scen = p.nseg0 * PER_SEG;  // units (LR_ROWS rows of a segment) per scenario
  int total = 0;
  for (int g = 0; g < plan->n_groups; ++g)
    if (plan->group_cls[g] == 3) total += plan->group_nactive[g] * per_scen;
  const int nwarps = gridDim.x * (LR_THREADS / 32);
  for (int u = blockIdx.x * (LR_THREADS / 32) + warp; u < total; u += nwarps) {
    int g = 0, base = 0;
    for (; g < plan->n_groups; ++g) {
      if (plan->group_cls[g] != 3) continue;
      const int cnt = plan->group_nactive[g] * per_scen;
      if (u < base + cnt) break;
      base += cnt;
    }
    const int v = u - base;
    const int rank = v / per_scen;
    const int w = v - rank * per_scen;
    const int seg = w / PER_SEG, half = w - seg * PER_SEG;
    const int valid = min(LR_ROWS, p.seg_valid0[seg] - LR_ROWS * half);
    if (valid <= 0) continue;
    const int s = p.active_list[plan->group_list_off[g] + rank];
    const size_t row0 = static_cast<size_t>(s) * p.L + p.seg_row0[seg] + LR_ROWS * half;
    const int op = plan->group_phaseA[g];
    const double coef = p.scen[s].coef_atm;
    const double* __restrict__ Ut = p.Ut[op];
    const double* __restrict__ Vt = p.Vt[op];
    const double* __restrict__ Irow = p.I + row0 * p.ld;
    double* __restrict__ Jrow = p.J + row0 * p.ld;

    // T[r][k] = sum_m I[row0 + r][m] Us[m][k]
    double acc[LR_ROWS][RP];
#pragma unroll
    for (int r = 0; r < LR_ROWS; ++r)
#pragma unroll
      for (int k = 0; k < RP; ++k) acc[r][k] = 0.0;
#pragma unroll 4
    for (int m = lane; m < p.N; m += 32) {  // (16 row loads in flight per warp: the kernel is HBM-latency bound)
      double x[LR_ROWS];
#pragma unroll
      for (int r = 0; r < LR_ROWS; ++r) x[r] = (r < valid) ? Irow[static_cast<size_t>(r) * p.ld + m] : 0.0;
#pragma unroll
      for (int k = 0; k < RP; ++k) {
        const double uk = Ut[static_cast<size_t>(k) * p.ldr + m];
#pragma unroll
        for (int r = 0; r < LR_ROWS; ++r) acc[r][k] = fma(x[r], uk, acc[r][k]);
      }
    }
#pragma unroll
    for (int r = 0; r < LR_ROWS; ++r)
#pragma unroll
      for (int k = 0; k < RP; ++k) {
        double t = acc[r][k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        acc[r][k] = t;  // (xor butterfly: every lane ends with the same sum, summed in the same order)
      }
    // J[row0 + r][m] = coef * sum_k T[r][k] Vt[k][m]
#pragma unroll 2
    for (int m = lane; m < p.N; m += 32) {
      double vk[RP];
#pragma unroll
      for (int k = 0; k < RP; ++k) vk[k] = Vt[static_cast<size_t>(k) * p.ldr + m];
#pragma unroll
      for (int r = 0; r < LR_ROWS; ++r) {
        if (r < valid) {
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < RP; ++k) sum = fma(acc[r][k], vk[k], sum);
          Jrow[static_cast<size_t>(r) * p.ld + m] = coef * sum;
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Closed-form factors of the molecular operands (no SVD anywhere on the product path).
//
// Summed over the two half rings the azimuth integrand of the Rayleigh builder is a polynomial in cos^2(Theta):
//   f(cc + x) + f(cc - x),  cc = mu_m mu_n,  x = sqrt((1 - mu_m^2)(1 - mu_n^2)) cos(phi)
//   = 0.75 (2 + 2 mu_m^2 mu_n^2 + 2 (1 - mu_m^2)(1 - mu_n^2) cos^2(phi))
// so every column of P -- and every ROW k of the contraction operand A[k][m] = w_k/4 P[m][N-1-k], whatever the
// per-column normalisation did -- is affine in mu_m^2:   A[k][m] = alpha_k + beta_k mu_m^2   (isotropic: beta = 0).
// The factors are read off two columns of the operand itself (mu^2 = 1 and mu = 0):
//   Vt = [1; mu^2],  Ut = [alpha; beta],  alpha_k = A[k][M-1],  beta_k = A[k][0] - A[k][M-1]
// and the caller enables the low-rank path only if the residual max|A - Ut^T Vt| / max|A| returned here is at rounding
// level.  Any operand with that structure qualifies, whichever builder made it; everything else stays dense.
// ---------------------------------------------------------------------------------------------
__global__ void lowrank_mu2_fit_kernel(const double* __restrict__ A, int lda, int N, int M, const double* __restrict__ mu,
                                       double* __restrict__ Ut, double* __restrict__ Vt, int ldr, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ldr) return;
  for (int r = 2; r < rows; ++r) { Ut[static_cast<size_t>(r) * ldr + i] = 0.0; Vt[static_cast<size_t>(r) * ldr + i] = 0.0; }
  if (i < N) {
    const double a0 = A[static_cast<size_t>(i) * lda + (M - 1)];   // mu = 0-
    const double a1 = A[static_cast<size_t>(i) * lda];             // mu = -1
    const double q0 = mu[M - 1] * mu[M - 1], q1 = mu[0] * mu[0];   // (0 and 1 on the reference grid; kept general)
    const double beta = (a1 - a0) / (q1 - q0);
    Ut[i] = a0 - beta * q0;
    Ut[ldr + i] = beta;
    Vt[i] = 1.0;
    Vt[ldr + i] = mu[i] * mu[i];
  } else {
    Ut[i] = Ut[ldr + i] = Vt[i] = Vt[ldr + i] = 0.0;
  }
}

// stats[0] = max|A - Ut^T Vt|, stats[1] = max|A|, stats[2] = max|beta| (bit patterns of non-negative doubles: atomicMax on u64)
__global__ void lowrank_mu2_residual_kernel(const double* __restrict__ A, int lda, int N, const double* __restrict__ Ut,
                                            const double* __restrict__ Vt, int ldr, unsigned long long* stats) {
  const int k = blockIdx.y;
  double res = 0.0, amax = 0.0;
  const double al = Ut[k], be = Ut[ldr + k];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < N; m += gridDim.x * blockDim.x) {
    const double a = A[static_cast<size_t>(k) * lda + m];
    const double fit = fma(be, Vt[ldr + m], al * Vt[m]);
    const double d = fabs(a - fit);
    res = (d > res || isnan(d)) ? d : res;
    amax = fmax(amax, fabs(a));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double r2 = __shfl_xor_sync(0xffffffffu, res, o);
    res = (r2 > res || isnan(r2)) ? r2 : res;
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (isnan(res)) res = INFINITY;
    atomicMax(&stats[0], static_cast<unsigned long long>(__double_as_longlong(res)));
    atomicMax(&stats[1], static_cast<unsigned long long>(__double_as_longlong(amax)));
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&stats[2], static_cast<unsigned long long>(__double_as_longlong(fabs(be))));
  }
}

}  // namespace sosgemm
