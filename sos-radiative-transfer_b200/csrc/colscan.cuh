// One scattering order for a BATCH of scenarios in one pass over memory: the layer sweeps of In_NumInt
// (SOS_Aer_I1_In.py:77-130; inlined three-region form SOS_Aer_main_specular.py:327-449, Lambert surface
// SOS_Aer_main_lambertian.py:399,401), the accumulation I += I_n (:454-456), the convergence ratios (:309) and -- on
// rows whose contraction operand is the rank <= 2 molecular one (gemm_lowrank.cuh) -- the source function itself
// (:315-323).
//
// sweep.cuh cuts the layer axis into chunks (local aggregates, carry chain, apply, zone fix-up: J is read twice,
// 40 B per element) because a single solve has nothing else to parallelise over.  A batch has: S scenarios x N columns
// are thousands of independent recurrences, so here the layer axis is NOT cut.
//
//   column_scan_kernel<UP>   one WARP streams 64 adjacent columns (two per lane: 16-byte accesses, 512 B per warp and
//                            row) of one scenario through all L rows with the running D / U in registers.  There is
//                            no cross-warp dependency and no barrier: every lane prefetches the rows it will need
//                            (I, and J where it is read) with cp.async into a shared-memory ring that only it reads
//                            back, 20 rows ahead, and writes I (and I_n where it is needed) straight from registers.
//                            Down pass and up pass are two launches; the specular / Lambert coupling is a read of the
//                            finished surface row in between.
//   zone_rows_kernel<UP>     one warp per (scenario, row) finishes the columns the reference treats specially next to
//                            mu = 0 (windowed / Taylor columns SOS_Aer_In_limit.py:70-109, the extrapolation :113-141,
//                            I_n[t, 0+] = J, the find-first blend SOS_Aer_I1_In.py:101-108) from the raw values the
//                            column kernel left in I_n for those columns, corrects I, and reduces the ratios.
//
// Generated source.  On rows that use the molecular operand alone (every row outside the aerosol layer,
// SOS_Aer_main_specular.py:323), A = Us Vt with Vt = [1; mu^2] (gemm_lowrank.cuh), so
//     J[t, m] = c0[t] + c1[t] mu_m^2,     c_r[t] = coef * sum_k I_{n-1}[t, k] Us[k][r]
// and the two numbers c_r[t] per row are all the next order needs from this one.  The column kernel therefore emits
// the partial projections of its own columns for every finished row (one slot per warp; butterfly transpose-reduce, a
// fixed tree: deterministic), zone_rows_kernel<UP> adds the zone columns and sums the slots in a fixed order, and the
// next order REBUILDS J from c instead of reading it: on those rows neither J nor I_n touches memory.  Per element and
// order that leaves the read-modify-write of I, 16 B, on 746 of the 800 default rows; the aerosol rows keep
// J 8 + I_n 8 + I 16 = 32 B (the minimum when J is materialised) and the DMMA contraction.
//
// Windowed columns (0.001 <= |mu| < 0.01, SOS_Aer_In_limit.py:96-107: trapezoid over tau' >= tau_t - 5|mu| inside the
// region).  With D the same recurrence restarted at the region's first row, the window integral is
//     D_t - exp((tau_t - tau_k0)/mu) D_k0,      k0 = first row of the window
// (the intervals above k0, decayed to t, cancel exactly), so the column costs one recurrence step per row in the column
// kernel (its D goes to a small history array) and one exp and one look-up in the zone kernel instead of a fresh sum
// over up to 5|mu|/dtau rows.  k0 depends only on tau and mu and is tabulated on the host with the reference's own
// rounding (tau_t - 5*abs(mu) as two operations).
//
// Region boundaries of the up pass: the first row of a region is read by the region above AFTER its
// blend, and the step across the boundary is pure attenuation (SOS_Aer_main_specular.py:413,433).  The blend of those
// (two) rows is done inside the column kernel by the CTA that owns the first 128 upward columns (two warps, the only
// place where warps of a CTA meet); a blend that reaches further raises SOS_STATUS_STRIP_FALLBACK and the solve is
// repeated with the chunked kernels.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "sweep.cuh"

namespace soscol {

using sossweep::exp_small;
using sossweep::kExpTaylor;

constexpr int CPL = 2;                  // columns per lane
constexpr int WCOLS = 32 * CPL;         // columns per warp
constexpr int WARPS = 2;                // warps per CTA (they meet only in the blend at region boundaries)
constexpr int BCOLS = WCOLS * WARPS;    // columns per CTA
constexpr int THREADS = 32 * WARPS;
constexpr int RG = 4;                   // rows per group (one cp.async group, unrolled together)
constexpr int RING = 6;                 // groups in the ring: RING - 1 groups (20 rows, 10 KB per warp) in flight
constexpr int PF = RING - 1;
constexpr int AUX = 16;                 // doubles per ring slot of per-row scalars: dtau[RG], c[RG][2], padding
constexpr int MAX_SMALL = 16;           // windowed / Taylor columns (|mu| < 0.01, without mu = 0-) a plan may have
constexpr int ZONE_ROWS = 8;            // rows (warps) per CTA of zone_rows_kernel

constexpr int WARP_SMEM = 2 * RING * RG * 32 * 16 + RING * AUX * 8;
constexpr int CTA_SMEM = WARPS * WARP_SMEM;

struct ColParams {
  GridDev g;
  const double* J;         // dense rows: the contraction's output
  double* In;
  double* I;
  double* saved;           // I_saved of SOS_Aer_main_specular.py:458 (or nullptr)
  const double* dt;        // [S][Lp]: down: tau[t] - tau[t-1] (0 at t = 0), up: tau[t+1] - tau[t] (0 at t = L-1); 0 on padding
  const double* cj;        // [S][Lp][2] coefficients of this order's generated source
  int Lp;
  double* proj;            // [S][L][nslots][2] partial projections of I_n onto the molecular factors
  int nslots, slot0;       // slot of (block b, warp w) = slot0 + b * WARPS + w
  int nblocks;             // CTAs per scenario
  int col_first;           // first column of block 0 (down: 0; up: the first even column > mu = 0+ ... see ue)
  int zlo, zu_end;         // zone columns: down m >= zlo, up m < zu_end (raw I_n stored, left out of the projections)
  int gen;                 // molecular rows rebuild J from cj
  int store_all;           // store I_n on every row (debugging aid)
  const double* Ut[SOS_MAX_PHASE];
  int rank[SOS_MAX_PHASE];
  int ldr;
  double* dhist;           // [S][L][MAX_SMALL] region-restarted D of the windowed columns
  const double* lam;       // [S] Lambert seed of the up pass (written by zone_rows_kernel<false>)
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int bytes = valid ? 16 : 0;  // 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// exp(x) for |x| <= 2^-10 (x^5/5! < 7e-18 relative): what almost every scan step of a thin atmosphere needs
constexpr double kTinyArg = 0.0009765625;
__device__ __forceinline__ double exp_tiny(double x) {
  double p = kExpTaylor[4];
  p = fma(p, x, kExpTaylor[5]);
  p = fma(p, x, kExpTaylor[6]);
  p = fma(p, x, kExpTaylor[7]);
  return fma(p, x, kExpTaylor[7]);
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// v[0..7] of every lane -> the sum over the warp of v[lane >> 2], in every lane (fixed tree: deterministic)
__device__ __forceinline__ double transpose_reduce8(double (&v)[8], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double keep = h16 ? v[i + 4] : v[i], send = h16 ? v[i] : v[i + 4];
    v[i] = keep + shfl_xor_d(send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double keep = h8 ? v[i + 2] : v[i], send = h8 ? v[i] : v[i + 2];
    v[i] = keep + shfl_xor_d(send, 8);
  }
  {
    const double keep = h4 ? v[1] : v[0], send = h4 ? v[0] : v[1];
    v[0] = keep + shfl_xor_d(send, 4);
  }
  v[0] += shfl_xor_d(v[0], 2);
  v[0] += shfl_xor_d(v[0], 1);
  return v[0];
}

template <bool UP>
__global__ void __launch_bounds__(THREADS, 4) column_scan_kernel(const __grid_constant__ ColParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ double s_row[BCOLS + 4];  // raw I_n[t, M + i] of a carry row (first up block only)
  __shared__ int s_istar;
  const GridDev& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ai = blockIdx.x / p.nblocks, blk = blockIdx.x - ai * p.nblocks;
  if (ai >= *g.n_active) return;
  const int s = g.active_flat[ai];
  const sos_scenario sc = g.scen[s];
  const int L = g.L, M = g.M, N = g.N;
  const size_t ld = g.ld;
  const int op = sc.phase_atm;
  const bool gen = p.gen && p.rank[op] > 0;
  const bool three = g.nreg == 3;
  const int a0 = three ? g.rstart[1] : L;  // rows [a0, a1) keep the dense contraction (aerosol layer)
  const int a1 = three ? g.rstart[2] : L;
  const int rs1 = three ? g.rstart[1] : -8, rs2 = three ? g.rstart[2] : -8;

  // ---- the two columns of this lane ----
  const int col0 = p.col_first + blk * BCOLS + warp * WCOLS + lane * CPL;
  const bool live = UP ? (col0 < N) : (col0 < M - 1);  // the pair holds at least one column of this pass
  bool stdc[CPL], win[CPL];
  double muc[CPL], imu[CPL], q[CPL], us0[CPL], us1[CPL];
  int csm[CPL];
  double imumax = 0.0;
  bool lane_zone = false, lane_win = false;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int m = col0 + c;
    const bool valid = UP ? (m > M && m < N) : (m < M - 1);
    const double mu = valid ? g.mu[m] : (UP ? 1.0 : -1.0);
    const bool small = !UP && valid && fabs(mu) < SOS_MU_THRESHOLD;
    stdc[c] = valid && !small;
    win[c] = small && fabs(mu) >= SOS_MU_VERY_SMALL;
    const bool zone = UP ? (m < p.zu_end) : (m >= p.zlo);
    muc[c] = mu;
    // x = dtau * imu is the (negative) exponent of the attenuation in either direction; columns that have no recurrence
    // in this pass (Taylor columns, mu = 0, beyond the grid) get imu = 0: x = 0, a = 1, b = 0, their X stays 0
    imu[c] = (stdc[c] || win[c]) ? (UP ? -1.0 / mu : 1.0 / mu) : 0.0;
    q[c] = mu * mu;
    const bool projected = gen && stdc[c] && !zone;
    us0[c] = projected ? p.Ut[op][m] : 0.0;
    us1[c] = projected ? p.Ut[op][p.ldr + m] : 0.0;
    csm[c] = m - g.first_small;
    imumax = fmax(imumax, fabs(imu[c]));
    lane_zone |= valid && zone;
    lane_win |= win[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) imumax = fmax(imumax, shfl_xor_d(imumax, o));
  const bool warp_win = __any_sync(0xffffffffu, lane_win);
  const bool blend_cta = UP && three && blk == 0;  // owns the columns next to mu = 0+: blends the carry rows

  // ---- rings ----
  uint8_t* wsm = smem_raw + static_cast<size_t>(warp) * WARP_SMEM;
  double2* Iring = reinterpret_cast<double2*>(wsm) + lane;
  double2* Jring = reinterpret_cast<double2*>(wsm + RING * RG * 32 * 16) + lane;
  double* aux = reinterpret_cast<double*>(wsm + 2 * RING * RG * 32 * 16);
  const int ngr = (L + RG - 1) / RG;
  const double* __restrict__ dts = p.dt + static_cast<size_t>(s) * p.Lp;
  const double* __restrict__ cjs = p.cj + static_cast<size_t>(s) * p.Lp * 2;
  const size_t ldb = ld * sizeof(double);
  const size_t lane_off = live ? (static_cast<size_t>(s) * L * ld + col0) * sizeof(double) : 0;
  const char* gI = reinterpret_cast<const char*>(p.I) + lane_off;
  const char* gJ = reinterpret_cast<const char*>(p.J) + lane_off;

  // loads of one group of rows (t0 .. t0 + RG - 1) into ring slot `slot`
  auto issue = [&](int t0, int slot) {
    const char* srcI = gI + static_cast<size_t>(t0) * ldb;
    const char* srcJ = gJ + static_cast<size_t>(t0) * ldb;
    double2* dI = Iring + slot * (RG * 32);
    double2* dJ = Jring + slot * (RG * 32);
    const bool anyJ = !gen || (t0 < a1 && t0 + RG > a0);
    if (t0 + RG <= L) {
#pragma unroll
      for (int r = 0; r < RG; ++r) cp_async16(dI + r * 32, srcI + r * ldb, live);
      if (anyJ) {
#pragma unroll
        for (int r = 0; r < RG; ++r)
          if (!gen || (t0 + r >= a0 && t0 + r < a1)) cp_async16(dJ + r * 32, srcJ + r * ldb, live);
      }
    } else {
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const bool ok = live && t0 + r < L;
        cp_async16(dI + r * 32, ok ? srcI + r * ldb : gI, ok);
        if (!gen || (t0 + r >= a0 && t0 + r < a1)) cp_async16(dJ + r * 32, ok ? srcJ + r * ldb : gJ, ok);
      }
    }
    if (lane < RG) cp_async8(&aux[slot * AUX + lane], dts + t0 + lane);
    else if (gen && lane < 2 * RG) cp_async16(&aux[slot * AUX + RG + 2 * (lane - RG)], cjs + 2 * (t0 + lane - RG), true);
  };

  // ---- running state ----
  double X[CPL] = {0.0, 0.0};   // D (down) or U (up) of the two columns
  double jp[CPL] = {0.0, 0.0};  // J of the row processed before
  if (UP) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int m = col0 + c;
      if (stdc[c]) {
        if (g.surface == SOS_SURFACE_SPECULAR) X[c] = sc.grd_alb * p.In[(static_cast<size_t>(s) * L + (L - 1)) * ld + (N - 1 - m)];
        else if (g.surface == SOS_SURFACE_LAMBERT) X[c] = p.lam[s];
      }
    }
  }
  double* const projs = p.proj + (static_cast<size_t>(s) * L * p.nslots + p.slot0 + blk * WARPS + warp) * 2;
  const int tstep = UP ? -RG : RG;
  int pf_t0 = UP ? (ngr - 1) * RG : 0;  // next group to prefetch
  int pf_slot = 0, pf_left = ngr;
#pragma unroll 1
  for (int i = 0; i < PF; ++i) {
    if (pf_left > 0) { issue(pf_t0, pf_slot); pf_t0 += tstep; pf_slot = (pf_slot + 1 == RING) ? 0 : pf_slot + 1; --pf_left; }
    cp_async_commit();
  }

  int t0 = UP ? (ngr - 1) * RG : 0;
  int slot = 0;
#pragma unroll 1
  for (int gi = 0; gi < ngr; ++gi, t0 += tstep, slot = (slot + 1 == RING) ? 0 : slot + 1) {
    cp_async_wait<PF - 1>();
    __syncwarp();  // the group's per-row scalars (copied by lanes 0..7) are visible; everyone is done with the slot refilled next
    if (pf_left > 0) { issue(pf_t0, pf_slot); pf_t0 += tstep; pf_slot = (pf_slot + 1 == RING) ? 0 : pf_slot + 1; --pf_left; }
    cp_async_commit();

    double dt[RG], c0[RG], c1[RG];
    {
      const double2* ax = reinterpret_cast<const double2*>(aux + slot * AUX);
      const double2 d01 = ax[0], d23 = ax[1];
      dt[0] = d01.x; dt[1] = d01.y; dt[2] = d23.x; dt[3] = d23.y;
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        double2 cc = make_double2(0.0, 0.0);
        if (gen) cc = ax[2 + r];
        c0[r] = cc.x;
        c1[r] = cc.y;
      }
    }
    const double dtmax = fmax(fmax(dt[0], dt[1]), fmax(dt[2], dt[3]));
    const bool tiny = dtmax * imumax <= kTinyArg;  // warp-uniform
    const bool full = t0 + RG <= L;
    const bool dense_all = !gen || (t0 >= a0 && t0 + RG <= a1);
    const bool gen_all = gen && (t0 + RG <= a0 || t0 >= a1);
    // rows that need the generic path: carry / gap rows of the up pass, region starts of windowed columns, the row whose
    // I_n every lane stores (row 0 going up, the surface row going down)
    const bool boundary = (rs1 >= t0 && rs1 <= t0 + RG) || (rs2 >= t0 && rs2 <= t0 + RG);
    const bool forced = UP ? (t0 == 0) : (t0 + RG >= L);
    const bool fast = full && !((UP || warp_win) && boundary) && (dense_all || (gen_all && !forced));
    const double2* Is = Iring + slot * (RG * 32);
    const double2* Js = Jring + slot * (RG * 32);
    char* oI = const_cast<char*>(gI) + static_cast<size_t>(t0) * ldb;
    const size_t d_in = reinterpret_cast<const char*>(p.In) - reinterpret_cast<const char*>(p.I);
    double pr0[RG], pr1[RG];

    if (fast) {
      // ---------- whole group, one kind of row, no boundary: straight-line code ----------
      auto run = [&](auto dense_tag, auto tiny_tag) {
        constexpr bool DENSE = decltype(dense_tag)::value, TINY = decltype(tiny_tag)::value;
        double2 iv[RG];
        double jj[RG][CPL], a[RG][CPL], hx[RG][CPL];
#pragma unroll
        for (int r = 0; r < RG; ++r) {
          iv[r] = Is[r * 32];
          if (DENSE) {
            const double2 jv = Js[r * 32];
            jj[r][0] = jv.x;
            jj[r][1] = jv.y;
          } else {
            jj[r][0] = fma(c1[r], q[0], c0[r]);
            jj[r][1] = fma(c1[r], q[1], c0[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const double x = dt[r] * imu[c];
            a[r][c] = TINY ? exp_tiny(x) : exp_small(x);
            hx[r][c] = 0.5 * x;
          }
        double val[RG][CPL];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
          const int r = UP ? RG - 1 - k : k;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const double b = hx[r][c] * fma(jp[c], a[r][c], jj[r][c]);
            X[c] = fma(X[c], a[r][c], -b);
            jp[c] = jj[r][c];
            val[r][c] = X[c];
          }
          if (!UP && warp_win) {  // windowed columns: the recurrence goes to the history array, I_n comes from the zone kernel
#pragma unroll
            for (int c = 0; c < CPL; ++c)
              if (win[c]) {
                p.dhist[(static_cast<size_t>(s) * L + t0 + r) * MAX_SMALL + csm[c]] = X[c];
                val[r][c] = 0.0;
              }
          }
        }
        if (live) {
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            char* o = oI + r * ldb;
            *reinterpret_cast<double2*>(o) = make_double2(iv[r].x + val[r][0], iv[r].y + val[r][1]);
            if (DENSE || lane_zone || p.store_all) *reinterpret_cast<double2*>(o + d_in) = make_double2(val[r][0], val[r][1]);
          }
          if (p.saved) {
            char* oS = reinterpret_cast<char*>(p.saved) + lane_off + static_cast<size_t>(t0) * ldb;
#pragma unroll
            for (int r = 0; r < RG; ++r) *reinterpret_cast<double2*>(oS + r * ldb) = make_double2(val[r][0], val[r][1]);
          }
        }
        if (!DENSE) {
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            pr0[r] = fma(val[r][1], us0[1], val[r][0] * us0[0]);
            pr1[r] = fma(val[r][1], us1[1], val[r][0] * us1[0]);
          }
        }
      };
      if (dense_all) {
        if (tiny) run(std::true_type{}, std::true_type{}); else run(std::true_type{}, std::false_type{});
      } else {
        if (tiny) run(std::false_type{}, std::true_type{}); else run(std::false_type{}, std::false_type{});
      }
    } else {
      // ---------- generic rows ----------
      auto row_step = [&](int r) {
        const int t = t0 + r;
        const bool dense = !gen || (t >= a0 && t < a1);
        const double2 iv = Is[r * 32];
        double2 jv = make_double2(0.0, 0.0);
        if (dense) jv = Js[r * 32];
        const double jj[CPL] = {dense ? jv.x : fma(c1[r], q[0], c0[r]), dense ? jv.y : fma(c1[r], q[1], c0[r])};
        double val[CPL];
        const bool gap = UP && (t + 1 == rs1 || t + 1 == rs2);
        const bool carry = UP && (t == rs1 || t == rs2);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const double x = dt[r] * imu[c];
          if (gap) {
            // the step across a region boundary is pure attenuation of the (blended) carry row
            X[c] = X[c] * exp(-dt[r] / muc[c]);
          } else {
            const double a = exp_small(x);
            X[c] = X[c] * a - (0.5 * x) * (jp[c] * a + jj[c]);
          }
          if (!UP && win[c] && (t == rs1 || t == rs2)) X[c] = 0.0;  // windows restart with their region
          val[c] = stdc[c] ? X[c] : 0.0;
          jp[c] = jj[c];
        }
        if (!UP && warp_win && t < L) {
#pragma unroll
          for (int c = 0; c < CPL; ++c)
            if (win[c]) p.dhist[(static_cast<size_t>(s) * L + t) * MAX_SMALL + csm[c]] = X[c];
        }
        pr0[r] = fma(val[1], us0[1], val[0] * us0[0]);
        pr1[r] = fma(val[1], us1[1], val[0] * us1[0]);
        if (live && t < L) {
          char* o = oI + r * ldb;
          *reinterpret_cast<double2*>(o) = make_double2(iv.x + val[0], iv.y + val[1]);
          const bool store_in = dense || p.store_all || lane_zone || t == (UP ? 0 : L - 1);
          if (store_in) *reinterpret_cast<double2*>(o + d_in) = make_double2(val[0], val[1]);
          if (p.saved) *reinterpret_cast<double2*>(reinterpret_cast<char*>(p.saved) + lane_off + static_cast<size_t>(t) * ldb) = make_double2(val[0], val[1]);
        }
        if (carry && blend_cta) {
          // the region above reads this row after its blend (SOS_Aer_I1_In.py:101-108)
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const int i = col0 + c - M;
            if (i > 0 && i < BCOLS + 4 && col0 + c < N) s_row[i] = X[c];
          }
          if (threadIdx.x == 0) s_row[0] = dense ? p.J[static_cast<size_t>(s) * L * ld + static_cast<size_t>(t) * ld + M] : fma(c1[r], g.mu[M] * g.mu[M], c0[r]);
          __syncthreads();
          if (warp == 0) {
            const int lim = min(p.zu_end, N) - M;  // positions 0 .. lim-1 are in s_row
            int istar = -1;
            for (int base = 1; base + 2 <= lim - 1 && istar < 0; base += 32) {
              const int i = base + lane;
              bool hit = false;
              if (i + 2 <= lim - 1) {
                const double x0 = s_row[i], x1 = s_row[i + 1], x2 = s_row[i + 2];
                hit = !(fabs((x0 - x1) - (x1 - x2)) > SOS_BLEND_THRESHOLD);
              }
              const unsigned mask = __ballot_sync(0xffffffffu, hit);
              if (mask) istar = base + __ffs(mask) - 1 + 1;
            }
            if (lane == 0) {
              s_istar = istar;
              if (istar < 0) atomicOr(&g.state[s].status, (p.zu_end >= N) ? SOS_STATUS_BLEND_OVERRUN : SOS_STATUS_STRIP_FALLBACK);
            }
          }
          __syncthreads();
          const int istar = s_istar;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const int i = col0 + c - M;
            if (istar > 0 && i > 0 && i < istar) {
              const double w = muc[c] / g.mu[M + istar];
              X[c] = (1.0 - w) * s_row[0] + w * s_row[istar];
            }
          }
          __syncthreads();  // s_row is reused at the next boundary
        }
      };
      if (UP) {
#pragma unroll
        for (int r = RG - 1; r >= 0; --r) row_step(r);
      } else {
#pragma unroll
        for (int r = 0; r < RG; ++r) row_step(r);
      }
    }

    if (gen && !dense_all) {
      double v[8] = {pr0[0], pr0[1], pr0[2], pr0[3], pr1[0], pr1[1], pr1[2], pr1[3]};
      const double tot = transpose_reduce8(v, lane);
      const int idx = lane >> 2, r = idx & 3, comp = idx >> 2;
      if ((lane & 3) == 0 && t0 + r < L) projs[static_cast<size_t>(t0 + r) * p.nslots * 2 + comp] = tot;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// rows: the columns next to mu = 0, the ratios, the next order's source coefficients
// ---------------------------------------------------------------------------------------------
struct ZoneParams {
  GridDev g;
  const double* J;
  double* In;
  double* I;
  double* saved;
  const double* cj_in;   // [S][Lp][2] this order's generated-source coefficients
  double* cj_out;        // ... the next order's (written by the up kernel)
  int Lp;
  double* proj;          // [S][L][nslots][2]
  int nslots, zone_slot; // the down kernel leaves its zone columns' projections in zone_slot
  int zlo, zu_end;
  int gen;
  const double* Ut[SOS_MAX_PHASE];
  int rank[SOS_MAX_PHASE];
  int ldr;
  const double* dhist;
  const int* k0tab;      // [S][L][nsc] first row of every window (host-built)
  int nsc;               // small columns first_small .. M-2
  double* lam;           // [S] Lambert seeds
  int zone_buf;          // doubles per warp of shared memory (>= M - zlo)
};

template <bool UP>
__global__ void __launch_bounds__(32 * ZONE_ROWS) zone_rows_kernel(const __grid_constant__ ZoneParams p) {
  extern __shared__ double sm_zone[];
  const GridDev& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y, t = blockIdx.x * ZONE_ROWS + warp;
  const int L = g.L, M = g.M, N = g.N, ld = g.ld;
  if (t >= L || !g.state[s].active) return;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const size_t fbase = static_cast<size_t>(s) * L * ld;
  const double* __restrict__ Js = p.J + fbase;
  double* __restrict__ Is = p.In + fbase;
  double* __restrict__ Ia = p.I + fbase;
  double* __restrict__ Sv = p.saved ? p.saved + fbase : nullptr;
  const sos_scenario sc = g.scen[s];
  const int op = sc.phase_atm;
  const bool gen = p.gen && p.rank[op] > 0;
  const bool three = g.nreg == 3;
  const int a0 = three ? g.rstart[1] : L, a1 = three ? g.rstart[2] : L;
  const double* __restrict__ cjs = p.cj_in + static_cast<size_t>(s) * p.Lp * 2;
  const double* __restrict__ Ut = p.Ut[op];
  int region = 0;
  while (region + 1 < g.nreg && t >= g.rstart[region + 1]) ++region;
  const size_t roff = static_cast<size_t>(t) * ld;
  auto source = [&](int tt, int m) -> double {  // J[tt, m]: rebuilt on the molecular rows, read on the dense ones
    if (gen && (tt < a0 || tt >= a1)) {
      const double mu = g.mu[m];
      return fma(cjs[2 * tt + 1], mu * mu, cjs[2 * tt]);
    }
    return Js[static_cast<size_t>(tt) * ld + m];
  };
  double p0 = 0.0, p1 = 0.0;  // projections of this row's zone columns (final values)

  if (!UP) {
    const int zlo = p.zlo;
    double* row = sm_zone + static_cast<size_t>(warp) * p.zone_buf - zlo;  // row[m] valid for m in [zlo, M)
    for (int m = zlo + lane; m < M; m += 32) {
      const bool std_col = (m < M - 1) && fabs(g.mu[m]) >= SOS_MU_THRESHOLD;
      row[m] = std_col ? Is[roff + m] : 0.0;  // raw (stored and accumulated by the column kernel)
    }
    __syncwarp();
    const int idxw = sc.extrap_width[region];
    const int r0 = g.rstart[region];
    // non-standard columns that survive the extrapolation
    const int hi = min(M - 1, M - idxw);
    {
      const int m = g.first_small + lane;
      if (lane < p.nsc && m < hi) {
        const double mu = g.mu[m];
        const double jt = source(t, m);
        double v;
        if (fabs(mu) < SOS_MU_VERY_SMALL) {  // Taylor: -J + mu dJ/dtau (SOS_Aer_In_limit.py:79-93)
          const double slope = (t > r0) ? (jt - source(t - 1, m)) / (tau[t] - tau[t - 1]) : 0.0;
          v = -jt + mu * slope;
        } else {  // window (:96-107) from the region-restarted recurrence
          const double* __restrict__ dh = p.dhist + static_cast<size_t>(s) * L * MAX_SMALL;
          const int k0 = p.k0tab[(static_cast<size_t>(s) * L + t) * p.nsc + lane];
          const double Dt = dh[static_cast<size_t>(t) * MAX_SMALL + lane];
          const double Dk = dh[static_cast<size_t>(k0) * MAX_SMALL + lane];
          v = Dt - exp((tau[t] - tau[k0]) / mu) * Dk;
          if (!isfinite(v)) v = -jt;  // (:104-105)
        }
        row[m] = v;
        Is[roff + m] = v;
        if (Sv) Sv[roff + m] = v;
        Ia[roff + m] += v;
      }
    }
    __syncwarp();
    if (idxw > 0) {
      // sources (columns < M - idx) and targets (columns >= M - idx) never overlap
      const int wclass = sossweep::width_class(g, idxw);
      const int ns = g.wns[wclass];
      const int src0 = (idxw < 2) ? (M - idxw - 2) : (M - idxw - ns);
      const double* __restrict__ W = g.W + g.woff[wclass];
      for (int i = lane; i < idxw; i += 32) {
        const int m = M - 1 - i;
        double v = 0.0;
        for (int k = 0; k < ns; ++k) v += W[i * ns + k] * row[src0 + k];
        const bool std_col = (m < M - 1) && fabs(g.mu[m]) >= SOS_MU_THRESHOLD;
        const double raw = row[m];
        row[m] = v;
        Is[roff + m] = v;
        if (Sv) Sv[roff + m] = v;
        Ia[roff + m] += std_col ? (v - raw) : v;  // standard targets were accumulated raw
      }
    } else if (lane == 0) {
      row[M - 1] = 0.0;
      Is[roff + M - 1] = 0.0;  // mu = 0- stays 0 when nothing is extrapolated
      if (Sv) Sv[roff + M - 1] = 0.0;
    }
    __syncwarp();
    if (gen) {
      for (int m = zlo + lane; m < M; m += 32) {
        p0 = fma(row[m], Ut[m], p0);
        p1 = fma(row[m], Ut[p.ldr + m], p1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { p0 += shfl_xor_d(p0, o); p1 += shfl_xor_d(p1, o); }
      if (lane == 0)
        *reinterpret_cast<double2*>(p.proj + ((static_cast<size_t>(s) * L + t) * p.nslots + p.zone_slot) * 2) = make_double2(p0, p1);
    }
    if (t == L - 1) {
      // the surface row is final: ratio of the downward half (SOS_Aer_main_specular.py:309), Lambert seed
      __threadfence_block();
      __syncwarp();
      double rmax = -INFINITY, part = 0.0;
      bool nonfinite = false;
      for (int m = lane; m < M; m += 32) {
        const double v = Is[roff + m];
        const double r = v / Ia[roff + m];
        if (isnan(r)) nonfinite = true; else rmax = fmax(rmax, r);
        if (g.surface == SOS_SURFACE_LAMBERT && m <= M - 2) {
          // -2 rho trapz(I mu, mu) over columns M-2 .. 0 as a weighted sum (SOS_Aer_main_lambertian.py:399,401)
          double wgt = 0.0;
          if (m >= 1) wgt += (g.mu[m - 1] - g.mu[m]) * 0.5;
          if (m + 1 <= M - 2) wgt += (g.mu[m] - g.mu[m + 1]) * 0.5;
          part += wgt * v * g.mu[m];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rmax = fmax(rmax, shfl_xor_d(rmax, o));
        part += shfl_xor_d(part, o);
      }
      nonfinite = __any_sync(0xffffffffu, nonfinite);
      if (lane == 0) {
        g.state[s].ratio_surf = rmax;
        if (nonfinite || rmax == INFINITY) atomicOr(&g.state[s].status, SOS_STATUS_NONFINITE);
        p.lam[s] = -2.0 * sc.grd_alb * part;
      }
    }
    return;
  }

  // ---------------- up ----------------
  const int zu_end = min(p.zu_end, N);
  const double v0 = source(t, M);  // I_n[t, mu = 0+] = J[t, mu = 0+] (SOS_Aer_I1_In.py:100)
  if (lane == 0) {
    Is[roff + M] = v0;
    if (Sv) Sv[roff + M] = v0;
    Ia[roff + M] += v0;
  }
  int istar = -1;
  for (int base = M + 1; base + 2 <= zu_end - 1 && istar < 0; base += 32) {
    const int i = base + lane;
    bool hit = false;
    if (i + 2 <= zu_end - 1) {
      const double a = Is[roff + i], b = Is[roff + i + 1], cc = Is[roff + i + 2];
      hit = !(fabs((a - b) - (b - cc)) > SOS_BLEND_THRESHOLD);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) istar = base + __ffs(mask) - 1 + 1;
  }
  if (istar < 0) {
    if (lane == 0) atomicOr(&g.state[s].status, (p.zu_end >= N) ? SOS_STATUS_BLEND_OVERRUN : SOS_STATUS_STRIP_FALLBACK);
  } else {
    const double v1 = Is[roff + istar];
    const double mus = g.mu[istar];
    __syncwarp();  // every lane has read what it needs before anything is overwritten
    for (int m = M + 1 + lane; m < istar; m += 32) {
      const double w = g.mu[m] / mus;
      const double val = (1.0 - w) * v0 + w * v1;
      const double old = Is[roff + m];
      Is[roff + m] = val;
      if (Sv) Sv[roff + m] = val;
      Ia[roff + m] += (val - old);
    }
  }
  __threadfence_block();
  __syncwarp();
  if (gen) {
    // the next order's source coefficients of this row: zone columns (final values) + the column kernels' slots
    for (int m = M + lane; m < zu_end; m += 32) {
      const double v = Is[roff + m];
      p0 = fma(v, Ut[m], p0);
      p1 = fma(v, Ut[p.ldr + m], p1);
    }
    const double* __restrict__ pr = p.proj + (static_cast<size_t>(s) * L + t) * p.nslots * 2;
    for (int j = lane; j < p.nslots; j += 32) { p0 += pr[2 * j]; p1 += pr[2 * j + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { p0 += shfl_xor_d(p0, o); p1 += shfl_xor_d(p1, o); }
    if (lane == 0)
      *reinterpret_cast<double2*>(p.cj_out + (static_cast<size_t>(s) * p.Lp + t) * 2) = make_double2(sc.coef_atm * p0, sc.coef_atm * p1);
  }
  if (t == 0) {
    double rmax = -INFINITY;
    bool nonfinite = false;
    for (int m = M + lane; m < N; m += 32) {
      const double r = Is[roff + m] / Ia[roff + m];
      if (isnan(r)) nonfinite = true; else rmax = fmax(rmax, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, shfl_xor_d(rmax, o));
    nonfinite = __any_sync(0xffffffffu, nonfinite);
    if (lane == 0) {
      g.state[s].ratio_toa = rmax;
      if (nonfinite || rmax == INFINITY) atomicOr(&g.state[s].status, SOS_STATUS_NONFINITE);
    }
  }
}

// Source coefficients of a whole field (the first order, before the loop): cj[s][t][r] = coef * sum_k I[t, k] Us[k][r]
__global__ void __launch_bounds__(256) project_rows_kernel(const GridDev g, const double* __restrict__ I, const double* const* Ut_tab,
                                                           const int* rank_tab, int ldr, int Lp, double* __restrict__ cj) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y, t = blockIdx.x * 8 + warp;
  if (t >= g.L) return;
  const int op = g.scen[s].phase_atm;
  if (rank_tab[op] == 0) return;
  const double* __restrict__ Ut = Ut_tab[op];
  const double* __restrict__ row = I + (static_cast<size_t>(s) * g.L + t) * g.ld;
  double p0 = 0.0, p1 = 0.0;
  for (int m = lane; m < g.N; m += 32) {
    const double x = row[m];
    p0 = fma(x, Ut[m], p0);
    p1 = fma(x, Ut[ldr + m], p1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    p0 += shfl_xor_d(p0, o);
    p1 += shfl_xor_d(p1, o);
  }
  const double coef = g.scen[s].coef_atm;
  if (lane == 0) *reinterpret_cast<double2*>(cj + (static_cast<size_t>(s) * Lp + t) * 2) = make_double2(coef * p0, coef * p1);
}

}  // namespace soscol
