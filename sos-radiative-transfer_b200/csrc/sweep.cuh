// Layer integration of one scattering order: In_NumInt (SOS_Aer_I1_In.py:77-130) and its inlined
// three-region form (SOS_Aer_main_specular.py:327-449; Lambert surface SOS_Aer_main_lambertian.py:399,401).
//
// The reference evaluates every (layer, mu) value as a trapezoid over the whole slice above/below it
// (O(L^2 N)).  With a_t = exp(-dtau_t/|mu|) this is the linear recurrence
//     down:  D_t = D_{t-1} a_t - dtau_t/(2 mu) (J_{t-1} a_t + J_t)
//     up:    U_t = U_{t+1} a_t + s_t dtau_t/(2 mu) (J_t + J_{t+1} a_t)      (s_t = 0 on carry gaps)
// which is evaluated here as a chunked scan in three launches:
//   1. sweep_local   every (scenario, chunk, mu column) runs its chunk from a zero carry and keeps only
//                    the chunk aggregate = the last local value (the products of a_t telescope to
//                    exp((tau_t - tau_start)/mu), so no separate "a" aggregate is needed); reads J once.
//   2. sweep_carry   one CTA per scenario chains the chunk carries (down), finishes the surface row
//                    (mu->0 columns + extrapolation), applies the specular/Lambert coupling, chains the
//                    up carries and re-seeds them from the *blended* boundary rows.
//   3. sweep_apply   the same (scenario, chunk, column) threads rerun the recurrence from the TRUE carry,
//                    write I_n and accumulate I += I_n (SOS_Aer_main_specular.py:454-456) in one pass;
//      sweep_zone    one small CTA per (scenario, layer) finishes the ~160 columns next to mu = 0:
//                    windowed / Taylor columns (SOS_Aer_In_limit.py:70-109), the polynomial
//                    extrapolation (:113-141, a fixed linear map W), the find-first second-difference
//                    blend (SOS_Aer_I1_In.py:101-108) and the convergence ratios of :309.
//   Traffic: J 8 (pass 1) + J 8 + I_n 8 + I 16 (pass 3) = 40 B per element (32 B is the algorithmic minimum).
#pragma once
#include "common.cuh"

namespace sossweep {

// exp(x) for the attenuation factors a_t = exp(-dtau/|mu|): almost every argument is tiny (dtau ~ 1e-4 ..
// 3e-3 per layer), where a degree-8 Taylor polynomial is exact to < 1e-19 relative (|x| <= 2^-5:
// x^9/9! < 1e-19) and costs 8 DFMA instead of the ~30 instructions of the general routine.  The two
// scan passes are FP64-pipe bound, not HBM bound, without this.
// (the coefficients sit in constant memory so that every DFMA takes its constant-bank operand directly: as 64-bit
// immediates they cost two UMOVs per DFMA, a third of all instructions of the issue-bound local pass)
__constant__ double kExpTaylor[8] = {1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
__device__ __forceinline__ double exp_small(double x) {
  if (fabs(x) <= 0.03125) {
    double p = kExpTaylor[0];
    p = fma(p, x, kExpTaylor[1]);
    p = fma(p, x, kExpTaylor[2]);
    p = fma(p, x, kExpTaylor[3]);
    p = fma(p, x, kExpTaylor[4]);
    p = fma(p, x, kExpTaylor[5]);
    p = fma(p, x, kExpTaylor[6]);
    p = fma(p, x, kExpTaylor[7]);
    return fma(p, x, kExpTaylor[7]);
  }
  return exp(x);
}

constexpr int LOCAL_THREADS = 128;
constexpr int LOCAL_UNROLL = 4;  // (8 rows in flight was measured: lower occupancy, 0.756 vs 0.744 ms per order at S = 96)
constexpr int ROW_THREADS = 256;
constexpr int CARRY_THREADS = 1024;

// ------------------------------------------------------------------------------------------
// 1. chunk-local recurrences
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LOCAL_THREADS)
sweep_local_kernel(const GridDev g, const double* __restrict__ J, double* __restrict__ aggD, double* __restrict__ aggU) {
  const int s = blockIdx.z;
  if (!g.state[s].active) return;
  const int c = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.N || m < g.col0 || m >= g.col1) return;
  const int t0 = g.chunk_start[c], t1 = g.chunk_start[c + 1];
  const int L = g.L, M = g.M, ld = g.ld;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const double* __restrict__ Js = J + static_cast<size_t>(s) * L * ld;
  const double mu = g.mu[m];
  const double imu = 1.0 / mu;  // one division per thread; the scan steps multiply
  const size_t agg = (static_cast<size_t>(s) * g.nchunks + c) * g.N + m;

  if (m < M - 1) {
    if (fabs(mu) < SOS_MU_THRESHOLD) return;  // windowed / Taylor columns are row-local (sweep_zone_kernel)
    double D = 0.0;
    int t = t0;
    double Jp;
    if (t == 0) {
      Jp = Js[m];
      t = 1;
    } else {
      Jp = Js[static_cast<size_t>(t - 1) * ld + m];
    }
    double tp = tau[t - 1 < 0 ? 0 : t - 1];
    // LOCAL_UNROLL independent loads / exps in flight, one dependent DFMA chain
    for (; t + LOCAL_UNROLL - 1 < t1; t += LOCAL_UNROLL) {
      double tc[LOCAL_UNROLL], jv[LOCAL_UNROLL], a[LOCAL_UNROLL], b[LOCAL_UNROLL];
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) {
        tc[u] = tau[t + u];
        jv[u] = Js[static_cast<size_t>(t + u) * ld + m];
      }
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) {
        const double d = tc[u] - (u ? tc[u - 1] : tp);
        a[u] = exp_small(d * imu);
        b[u] = (d * 0.5) * ((u ? jv[u - 1] : Jp) * a[u] + jv[u]) * imu;
      }
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) D = D * a[u] - b[u];
      Jp = jv[LOCAL_UNROLL - 1];
      tp = tc[LOCAL_UNROLL - 1];
    }
    for (; t < t1; ++t) {
      const double tc = tau[t];
      const double jc = Js[static_cast<size_t>(t) * ld + m];
      const double d = tc - tp;
      const double a = exp_small(d * imu);
      D = D * a - (d * 0.5) * (Jp * a + jc) * imu;
      Jp = jc;
      tp = tc;
    }
    aggD[agg] = D;
  } else if (m > M) {
    double U = 0.0;
    int t = t1 - 1;
    double Jn, tn;
    if (t == L - 1) {
      Jn = Js[static_cast<size_t>(t) * ld + m];
      tn = tau[t];
      --t;
    } else {
      Jn = Js[static_cast<size_t>(t + 1) * ld + m];
      tn = tau[t + 1];
      // chunk ends at a region boundary: the slice stops one row short of the carry row
      // (SOS_Aer_main_specular.py:413,433) -> pure attenuation, no source on this step
      if (g.chunk_region[c + 1] != g.chunk_region[c]) {
        Jn = Js[static_cast<size_t>(t) * ld + m];
        tn = tau[t];
        --t;
      }
    }
    for (; t - (LOCAL_UNROLL - 1) >= t0; t -= LOCAL_UNROLL) {
      double tc[LOCAL_UNROLL], jv[LOCAL_UNROLL], a[LOCAL_UNROLL], b[LOCAL_UNROLL];
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) {
        tc[u] = tau[t - u];
        jv[u] = Js[static_cast<size_t>(t - u) * ld + m];
      }
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) {
        const double d = (u ? tc[u - 1] : tn) - tc[u];
        a[u] = exp_small(-d * imu);
        b[u] = (d * 0.5) * (jv[u] + (u ? jv[u - 1] : Jn) * a[u]) * imu;
      }
#pragma unroll
      for (int u = 0; u < LOCAL_UNROLL; ++u) U = U * a[u] + b[u];
      Jn = jv[LOCAL_UNROLL - 1];
      tn = tc[LOCAL_UNROLL - 1];
    }
    for (; t >= t0; --t) {
      const double tc = tau[t];
      const double jc = Js[static_cast<size_t>(t) * ld + m];
      const double d = tn - tc;
      const double a = exp_small(-d * imu);
      U = U * a + (d * 0.5) * (jc + Jn * a) * imu;
      Jn = jc;
      tn = tc;
    }
    aggU[agg] = U;
  }
}

// ------------------------------------------------------------------------------------------
// row-level helpers of sweep_carry_kernel (one CTA works on one row in smem)
// ------------------------------------------------------------------------------------------

// |mu| < MU_THRESHOLD downward column m at layer t (improved_asymptotic_downward_radiance,
// SOS_Aer_In_limit.py:70-109).  One warp per call; every lane returns the value.
__device__ __forceinline__ double asymptotic_column(const GridDev& g, const double* __restrict__ Js,
                                                    const double* __restrict__ tau, int t, int r0, int m) {
  const int lane = threadIdx.x & 31;
  const int ld = g.ld;
  const double mu = g.mu[m];
  const double jt = Js[static_cast<size_t>(t) * ld + m];
  if (fabs(mu) < SOS_MU_VERY_SMALL) {  // Taylor: -J + mu dJ/dtau (:79-93)
    double slope = 0.0;
    if (t > r0) slope = (jt - Js[static_cast<size_t>(t - 1) * ld + m]) / (tau[t] - tau[t - 1]);
    return -jt + mu * slope;
  }
  // windowed trapezoid over tau' >= tau_t - 5|mu| inside the region slice (:96-107)
  const double tt = tau[t];
  const double lim = tt - 5.0 * fabs(mu);
  int lo = r0, hi = t;  // first k in [r0, t] with tau[k] >= lim (tau is non-decreasing)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (tau[mid] >= lim) hi = mid; else lo = mid + 1;
  }
  const int k0 = lo;
  double sum = 0.0;
  bool bad = false;
  // interval k..k+1 handled by lane (k-k0)%32; fixed order -> deterministic
  for (int k = k0 + lane; k < t; k += 32) {
    const double f0 = Js[static_cast<size_t>(k) * ld + m] * exp((tt - tau[k]) / mu);
    const double f1 = Js[static_cast<size_t>(k + 1) * ld + m] * exp((tt - tau[k + 1]) / mu);
    bad |= !isfinite(f0) || !isfinite(f1);
    sum += (tau[k + 1] - tau[k]) * (f1 + f0) * 0.5;
  }
  if (t == k0) bad |= !isfinite(jt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  bad = __any_sync(0xffffffffu, bad);
  if (bad) return -jt;  // (:104-105)
  return -sum / mu;
}

// Complete the downward half of a row held in smem `row[0..M-1]`:
// windowed/Taylor columns that survive the extrapolation, then the extrapolation itself.
// Must be called by all threads of the CTA; contains __syncthreads.
__device__ __forceinline__ void finish_down_row(const GridDev& g, double* row, const double* __restrict__ Js,
                                                const double* __restrict__ tau, int t, int region, int idx_width,
                                                int wclass) {
  const int M = g.M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int r0 = g.rstart[region];
  // columns first_small .. M-2 are non-standard; those >= M - idx are overwritten below -> skip them
  const int hi = min(M - 1, M - idx_width);
  for (int m = g.first_small + warp; m < hi; m += nwarps) {
    const double v = asymptotic_column(g, Js, tau, t, r0, m);
    if (lane == 0) row[m] = v;
  }
  __syncthreads();
  if (idx_width > 0) {
    // sources (columns < M - idx) and targets (columns >= M - idx) never overlap
    const int ns = g.wns[wclass];
    const int src0 = (idx_width < 2) ? (M - idx_width - 2) : (M - idx_width - ns);
    const double* __restrict__ W = g.W + g.woff[wclass];
    for (int i = threadIdx.x; i < idx_width; i += blockDim.x) {
      double v = 0.0;
      for (int j = 0; j < ns; ++j) v += W[i * ns + j] * row[src0 + j];
      row[M - 1 - i] = v;
    }
  }
  __syncthreads();
}

// Blend of the upward half towards mu = 0+ (SOS_Aer_I1_In.py:101-108).  row[M] must already hold J[t,M].
// `found` is a shared int.  All threads of the CTA must call; returns false on search overrun (Q11).
__device__ __forceinline__ bool blend_up_row(const GridDev& g, double* row, int* found) {
  const int M = g.M, N = g.col1;  // the search never leaves the owned columns
  if (threadIdx.x == 0) *found = 0x7fffffff;
  __syncthreads();
  for (int base = M + 1; base + 2 <= N - 1; base += blockDim.x) {
    const int i = base + threadIdx.x;
    if (i + 2 <= N - 1) {
      const double d = fabs((row[i] - row[i + 1]) - (row[i + 1] - row[i + 2]));
      if (!(d > SOS_BLEND_THRESHOLD)) atomicMin(found, i);
    }
    __syncthreads();
    const int f = *found;
    __syncthreads();
    if (f != 0x7fffffff) break;
  }
  const int f = *found;
  if (f == 0x7fffffff) return false;
  const int istar = f + 1;
  const double v0 = row[M], v1 = row[istar], mus = g.mu[istar];
  __syncthreads();
  for (int m = M + 1 + threadIdx.x; m < istar; m += blockDim.x) {
    const double w = g.mu[m] / mus;
    row[m] = (1.0 - w) * v0 + w * v1;
  }
  __syncthreads();
  return true;
}

__device__ __forceinline__ int width_class(const GridDev& g, int idx_width) {
  // widths are one of the four classes of sos_extrap_layout (or 0)
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (g.widx[k] == idx_width) c = k;
  return c;
}

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in all threads
__device__ __forceinline__ double block_sum(double v, double* scratch /*[32]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < nwarps; ++w) tot += scratch[w];
  __syncthreads();
  return tot;
}

// ------------------------------------------------------------------------------------------
// 2. carry chain (one CTA per scenario)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CARRY_THREADS)
sweep_carry_kernel(const GridDev g, const double* __restrict__ J, const double* __restrict__ aggD,
                   const double* __restrict__ aggU, double* __restrict__ carryD, double* __restrict__ carryU) {
  extern __shared__ double sm_row[];  // [N] + scratch[32]
  __shared__ int found;
  const int s = blockIdx.x;
  if (!g.state[s].active) return;
  const int L = g.L, M = g.M, N = g.N, ld = g.ld, nch = g.nchunks;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const double* __restrict__ Js = J + static_cast<size_t>(s) * L * ld;
  const sos_scenario sc = g.scen[s];
  double* row = sm_row;
  double* scratch = sm_row + N;
  const size_t base = static_cast<size_t>(s) * nch * N;

  // ---- down: chain the standard columns through the chunks ----
  // (the loads of the chunk aggregates and the decay factors do not depend on the carry: fetch them
  //  eight chunks at a time so that only the FMA chain is serial)
  for (int m = threadIdx.x; m < M - 1; m += blockDim.x) {
    const double mu = g.mu[m];
    double cd = 0.0;
    if (fabs(mu) >= SOS_MU_THRESHOLD && m >= g.col0 && m < g.col1) {
      for (int c0 = 0; c0 < nch; c0 += 8) {
        double ag[8], ex[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u;
          ag[u] = 0.0;
          ex[u] = 0.0;
          if (c < nch) {
            ag[u] = aggD[base + static_cast<size_t>(c) * N + m];
            const int t0 = g.chunk_start[c], t1 = g.chunk_start[c + 1];
            if (t0 > 0) ex[u] = exp((tau[t1 - 1] - tau[t0 - 1]) / mu);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u;
          if (c < nch) {
            carryD[base + static_cast<size_t>(c) * N + m] = cd;
            cd = ag[u] + cd * ex[u];
          }
        }
      }
    }
    row[m] = cd;  // true D at the surface row for standard columns
  }
  if (threadIdx.x == 0) row[M - 1] = 0.0;
  __syncthreads();

  // ---- surface row: finish the downward half, then couple ----
  const int last_region = g.nreg - 1;
  if (g.surface != SOS_SURFACE_NONE) {
    const int idxw = sc.extrap_width[last_region];
    finish_down_row(g, row, Js, tau, L - 1, last_region, idxw, width_class(g, idxw));
  }
  double lambert = 0.0;
  if (g.surface == SOS_SURFACE_LAMBERT) {
    // -2 rho trapz(I[L-1,c] mu[c], mu[c]) over c = M-2 .. 0 (descending abscissa, column M-1 excluded)
    double part = 0.0;
    for (int j = threadIdx.x; j < M - 2; j += blockDim.x) {
      const int c0 = M - 2 - j, c1 = c0 - 1;
      part += (g.mu[c1] - g.mu[c0]) * (row[c1] * g.mu[c1] + row[c0] * g.mu[c0]) * 0.5;
    }
    lambert = -2.0 * sc.grd_alb * block_sum(part, scratch);
  }
  // seeds into the upper half of the row buffer
  for (int m = M + 1 + threadIdx.x; m < N; m += blockDim.x) {
    double seed = 0.0;
    if (g.surface == SOS_SURFACE_SPECULAR) seed = sc.grd_alb * row[N - 1 - m];
    else if (g.surface == SOS_SURFACE_LAMBERT) seed = lambert;
    row[m] = seed;
  }
  __syncthreads();

  // ---- up: chain from the surface, re-seeding from blended rows at region boundaries ----
  int c = nch - 1;
  while (c >= 0) {
    // batch [ce, c]: at most 8 chunks, ending at the first chunk that starts a region
    int ce = c, nb = 1;
    while (nb < 8 && ce > 0 && g.chunk_region[ce - 1] == g.chunk_region[ce]) { --ce; ++nb; }
    for (int m = M + 1 + threadIdx.x; m < N; m += blockDim.x) {
      if (m < g.col0 || m >= g.col1) continue;
      const double mu = g.mu[m];
      double ag[8], ex[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cc = c - u;
        ag[u] = 0.0;
        ex[u] = 0.0;
        if (cc >= ce) {
          const int t0 = g.chunk_start[cc], t1 = g.chunk_start[cc + 1];
          ag[u] = aggU[base + static_cast<size_t>(cc) * N + m];
          ex[u] = exp(-(tau[(t1 == L) ? L - 1 : t1] - tau[t0]) / mu);
        }
      }
      double cu = row[m];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cc = c - u;
        if (cc >= ce) {
          carryU[base + static_cast<size_t>(cc) * N + m] = cu;
          cu = ag[u] + cu * ex[u];  // raw U at the first row of chunk cc
        }
      }
      row[m] = cu;
    }
    if (ce > 0 && g.chunk_region[ce - 1] != g.chunk_region[ce]) {
      // row chunk_start[ce] is the carry row of the region above and is read after its blend (A.7)
      if (threadIdx.x == 0) row[M] = Js[static_cast<size_t>(g.chunk_start[ce]) * ld + M];
      __syncthreads();
      if (!blend_up_row(g, row, &found) && threadIdx.x == 0) atomicOr(&g.state[s].status, SOS_STATUS_BLEND_OVERRUN);
    }
    c = ce - 1;
  }
}

// ------------------------------------------------------------------------------------------
// 3. apply the carries (chunk x column threads) and finish the mu -> 0 zone (row CTAs)
// ------------------------------------------------------------------------------------------

// first downward column of the row-wise zone: all non-standard columns, the extrapolation targets and
// their sources (largest width of the scenario)
__device__ __forceinline__ int zone_lo(const GridDev& g, const sos_scenario& sc) {
  int w = 0;
  for (int k = 0; k < g.nreg; ++k) w = max(w, sc.extrap_width[k]);
  const int ns = (w <= 0) ? 0 : (w < 2 ? 2 : min(5, w));
  return max(0, min(g.first_small, g.M - w - ns));
}

// Second sweep pass: the same recurrences as sweep_local_kernel, now started from the TRUE carry of the
// chunk, writing the final I_n and accumulating I += I_n in the same pass (J 8 + I_n 8 + I 16 bytes per
// element).  The few columns next to mu = 0 that the reference post-processes are corrected afterwards by
// sweep_zone_kernel (it replaces the raw value in I_n and adds the difference to I).
__global__ void __launch_bounds__(LOCAL_THREADS)
sweep_apply_kernel(const GridDev g, const double* __restrict__ J, double* __restrict__ In,
                   const double* __restrict__ carryD, const double* __restrict__ carryU,
                   double* __restrict__ I, double* __restrict__ saved) {
  const int s = blockIdx.z;
  if (!g.state[s].active) return;
  const int c = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.N || m < g.col0 || m >= g.col1) return;
  const int t0 = g.chunk_start[c], t1 = g.chunk_start[c + 1];
  const int L = g.L, M = g.M, ld = g.ld;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const size_t fbase = static_cast<size_t>(s) * L * ld;
  const double* __restrict__ Js = J + fbase;
  double* __restrict__ Is = In + fbase;
  double* __restrict__ Ia = I ? I + fbase : nullptr;
  double* __restrict__ Sv = saved ? saved + fbase : nullptr;
  const double mu = g.mu[m];
  const double imu = 1.0 / mu;  // one division per thread; the scan steps multiply
  const size_t agg = (static_cast<size_t>(s) * g.nchunks + c) * g.N + m;

// store I_n (and I_saved); accumulate with the I value that was prefetched together with J
#define SOS_EMIT(T_, V_, IOLD_)                                      \
  do {                                                               \
    const size_t o_ = static_cast<size_t>(T_) * ld + m;              \
    const double v_ = (V_);                                          \
    Is[o_] = v_;                                                     \
    if (Sv) Sv[o_] = v_;                                             \
    if (Ia) Ia[o_] = (IOLD_) + v_;                                   \
  } while (0)
#define SOS_IOLD(T_) (Ia ? Ia[static_cast<size_t>(T_) * ld + m] : 0.0)

  if (m < M - 1) {
    if (fabs(mu) < SOS_MU_THRESHOLD) return;
    double D = (t0 > 0) ? carryD[agg] : 0.0;
    int t = t0;
    double Jp;
    if (t == 0) {
      Jp = Js[m];
      SOS_EMIT(0, 0.0, SOS_IOLD(0));
      t = 1;
    } else {
      Jp = Js[static_cast<size_t>(t - 1) * ld + m];
    }
    double tp = tau[t - 1 < 0 ? 0 : t - 1];
    for (; t + 3 < t1; t += 4) {
      const double tc0 = tau[t], tc1 = tau[t + 1], tc2 = tau[t + 2], tc3 = tau[t + 3];
      const double j0 = Js[static_cast<size_t>(t) * ld + m];
      const double j1 = Js[static_cast<size_t>(t + 1) * ld + m];
      const double j2 = Js[static_cast<size_t>(t + 2) * ld + m];
      const double j3 = Js[static_cast<size_t>(t + 3) * ld + m];
      const double i0 = SOS_IOLD(t), i1 = SOS_IOLD(t + 1), i2 = SOS_IOLD(t + 2), i3 = SOS_IOLD(t + 3);
      const double d0 = tc0 - tp, d1 = tc1 - tc0, d2 = tc2 - tc1, d3 = tc3 - tc2;
      const double a0 = exp_small(d0 * imu), a1 = exp_small(d1 * imu), a2 = exp_small(d2 * imu), a3 = exp_small(d3 * imu);
      const double b0 = (d0 * 0.5) * (Jp * a0 + j0) * imu;
      const double b1 = (d1 * 0.5) * (j0 * a1 + j1) * imu;
      const double b2 = (d2 * 0.5) * (j1 * a2 + j2) * imu;
      const double b3 = (d3 * 0.5) * (j2 * a3 + j3) * imu;
      D = D * a0 - b0; SOS_EMIT(t, D, i0);
      D = D * a1 - b1; SOS_EMIT(t + 1, D, i1);
      D = D * a2 - b2; SOS_EMIT(t + 2, D, i2);
      D = D * a3 - b3; SOS_EMIT(t + 3, D, i3);
      Jp = j3;
      tp = tc3;
    }
    for (; t < t1; ++t) {
      const double tc = tau[t];
      const double jc = Js[static_cast<size_t>(t) * ld + m];
      const double io = SOS_IOLD(t);
      const double d = tc - tp;
      const double a = exp_small(d * imu);
      D = D * a - (d * 0.5) * (Jp * a + jc) * imu;
      SOS_EMIT(t, D, io);
      Jp = jc;
      tp = tc;
    }
  } else if (m > M) {
    double U = carryU[agg];  // value at the carry row (t1, or the surface seed for the last chunk)
    int t = t1 - 1;
    double Jn, tn;
    if (t == L - 1) {
      Jn = Js[static_cast<size_t>(t) * ld + m];
      tn = tau[t];
      SOS_EMIT(t, U, SOS_IOLD(t));  // zero-length integral: the surface row is the seed itself
      --t;
    } else {
      Jn = Js[static_cast<size_t>(t + 1) * ld + m];
      tn = tau[t + 1];
      if (g.chunk_region[c + 1] != g.chunk_region[c]) {  // carry gap: pure attenuation on this step
        const double tc = tau[t];
        U = U * exp(-(tn - tc) / mu);
        SOS_EMIT(t, U, SOS_IOLD(t));
        Jn = Js[static_cast<size_t>(t) * ld + m];
        tn = tc;
        --t;
      }
    }
    for (; t - 3 >= t0; t -= 4) {
      const double tc0 = tau[t], tc1 = tau[t - 1], tc2 = tau[t - 2], tc3 = tau[t - 3];
      const double j0 = Js[static_cast<size_t>(t) * ld + m];
      const double j1 = Js[static_cast<size_t>(t - 1) * ld + m];
      const double j2 = Js[static_cast<size_t>(t - 2) * ld + m];
      const double j3 = Js[static_cast<size_t>(t - 3) * ld + m];
      const double i0 = SOS_IOLD(t), i1 = SOS_IOLD(t - 1), i2 = SOS_IOLD(t - 2), i3 = SOS_IOLD(t - 3);
      const double d0 = tn - tc0, d1 = tc0 - tc1, d2 = tc1 - tc2, d3 = tc2 - tc3;
      const double a0 = exp_small(-d0 * imu), a1 = exp_small(-d1 * imu), a2 = exp_small(-d2 * imu), a3 = exp_small(-d3 * imu);
      const double b0 = (d0 * 0.5) * (j0 + Jn * a0) * imu;
      const double b1 = (d1 * 0.5) * (j1 + j0 * a1) * imu;
      const double b2 = (d2 * 0.5) * (j2 + j1 * a2) * imu;
      const double b3 = (d3 * 0.5) * (j3 + j2 * a3) * imu;
      U = U * a0 + b0; SOS_EMIT(t, U, i0);
      U = U * a1 + b1; SOS_EMIT(t - 1, U, i1);
      U = U * a2 + b2; SOS_EMIT(t - 2, U, i2);
      U = U * a3 + b3; SOS_EMIT(t - 3, U, i3);
      Jn = j3;
      tn = tc3;
    }
    for (; t >= t0; --t) {
      const double tc = tau[t];
      const double jc = Js[static_cast<size_t>(t) * ld + m];
      const double io = SOS_IOLD(t);
      const double d = tn - tc;
      const double a = exp_small(-d * imu);
      U = U * a + (d * 0.5) * (jc + Jn * a) * imu;
      SOS_EMIT(t, U, io);
      Jn = jc;
      tn = tc;
    }
  }
#undef SOS_IOLD
#undef SOS_EMIT
}

// Row-wise post-processing of the columns next to mu = 0: windowed / Taylor columns
// (SOS_Aer_In_limit.py:70-109), the extrapolation W (:113-141), I_n[t, mu=0+] = J (SOS_Aer_I1_In.py:100),
// the find-first second-difference blend (:101-108) and, on the TOA / surface rows, the convergence
// ratios of SOS_Aer_main_specular.py:309.  ONE WARP per (layer, scenario): it reads the ~40 raw values
// it needs, rewrites only the columns that change and corrects I by (new - raw) for those the apply
// pass had already accumulated.
constexpr int ZONE_ROWS = 8;

__global__ void __launch_bounds__(32 * ZONE_ROWS)  // (forcing 6 or 8 CTAs per SM spills and is slower: 0.755 vs 0.744 ms per order)
sweep_zone_kernel(const GridDev g, const double* __restrict__ J, double* __restrict__ In,
                  double* __restrict__ I, double* __restrict__ saved, int zone_buf) {
  extern __shared__ double sm_zone[];  // ZONE_ROWS x zone_buf: raw downward values of columns [zl, M)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y, t = blockIdx.x * ZONE_ROWS + warp;
  const int L = g.L, M = g.M, ld = g.ld;
  if (t >= L || !g.state[s].active) return;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const size_t fbase = static_cast<size_t>(s) * L * ld;
  const double* __restrict__ Js = J + fbase;
  double* __restrict__ Is = In + fbase;
  double* __restrict__ Ia = I ? I + fbase : nullptr;
  double* __restrict__ Sv = saved ? saved + fbase : nullptr;
  const sos_scenario sc = g.scen[s];
  const int region = g.chunk_region[g.row_chunk[t]];
  const size_t roff = static_cast<size_t>(t) * ld;
  const int c_lo = g.col0, c_hi = g.col1;
  const bool own_down_zone = (c_lo < M && c_hi >= M);
  const bool own_up_zone = (c_lo <= M && c_hi > M + 1);

  if (own_down_zone) {
    const int zl = zone_lo(g, sc);
    double* row = sm_zone + static_cast<size_t>(warp) * zone_buf - zl;  // row[m] valid for m in [zl, M)
    for (int m = zl + lane; m < M; m += 32) {
      const bool std_col = (m < M - 1) && fabs(g.mu[m]) >= SOS_MU_THRESHOLD;
      row[m] = std_col ? Is[roff + m] : 0.0;  // raw (already stored and accumulated by the apply pass)
    }
    __syncwarp();
    const int idxw = sc.extrap_width[region];
    const int r0 = g.rstart[region];
    // non-standard columns that survive the extrapolation: computed here, never touched by the apply pass
    const int hi = min(M - 1, M - idxw);
    for (int m = g.first_small; m < hi; ++m) {
      const double v = asymptotic_column(g, Js, tau, t, r0, m);
      if (lane == 0) {
        row[m] = v;
        Is[roff + m] = v;
        if (Sv) Sv[roff + m] = v;
        if (Ia) Ia[roff + m] += v;
      }
    }
    __syncwarp();
    if (idxw > 0) {
      const int wclass = width_class(g, idxw);
      const int ns = g.wns[wclass];
      const int src0 = (idxw < 2) ? (M - idxw - 2) : (M - idxw - ns);
      const double* __restrict__ W = g.W + g.woff[wclass];
      for (int i = lane; i < idxw; i += 32) {
        const int m = M - 1 - i;  // sources (< M - idx) and targets (>= M - idx) never overlap
        double v = 0.0;
        for (int k = 0; k < ns; ++k) v += W[i * ns + k] * row[src0 + k];
        const bool std_col = (m < M - 1) && fabs(g.mu[m]) >= SOS_MU_THRESHOLD;
        Is[roff + m] = v;
        if (Sv) Sv[roff + m] = v;
        if (Ia) Ia[roff + m] += std_col ? (v - row[m]) : v;  // standard targets were accumulated raw
      }
    } else if (lane == 0) {
      Is[roff + M - 1] = 0.0;  // mu = 0- stays 0 when nothing is extrapolated
      if (Sv) Sv[roff + M - 1] = 0.0;
    }
    __syncwarp();
  }

  if (own_up_zone) {
    const double v0 = Js[roff + M];  // I_n[t, mu = 0+] = J[t, mu = 0+]
    if (lane == 0) {
      Is[roff + M] = v0;
      if (Sv) Sv[roff + M] = v0;
      if (Ia) Ia[roff + M] += v0;
    }
    // find-first over the raw values written by the apply pass
    const int lim = c_hi;  // the search never leaves the owned columns
    int istar = -1;
    for (int base = M + 1; base + 2 <= lim - 1 && istar < 0; base += 32) {
      const int i = base + lane;
      bool hit = false;
      if (i + 2 <= lim - 1) {
        const double a = Is[roff + i], b = Is[roff + i + 1], cc = Is[roff + i + 2];
        hit = !(fabs((a - b) - (b - cc)) > SOS_BLEND_THRESHOLD);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) istar = base + __ffs(mask) - 1 + 1;
    }
    if (istar < 0) {
      if (lane == 0) atomicOr(&g.state[s].status, SOS_STATUS_BLEND_OVERRUN);
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
        if (Ia) Ia[roff + m] += (val - old);
      }
    }
    __syncwarp();
  }

  // ---- convergence ratios on the TOA / surface rows (whole half-row, read back from global) ----
  const bool toa = (t == 0), surf = (t == L - 1);
  if (Ia && (toa || surf)) {
    __threadfence_block();
    __syncwarp();
    double rmax = -INFINITY;
    bool nonfinite = false;
    const int a0 = toa ? max(M, c_lo) : c_lo;
    const int a1 = toa ? c_hi : min(M, c_hi);
    for (int m = a0 + lane; m < a1; m += 32) {
      const double r = Is[roff + m] / Ia[roff + m];
      if (isnan(r)) nonfinite = true; else rmax = fmax(rmax, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    nonfinite = __any_sync(0xffffffffu, nonfinite);
    if (lane == 0) {
      if (toa) g.state[s].ratio_toa = rmax;
      if (surf) g.state[s].ratio_surf = rmax;
      if (nonfinite || rmax == INFINITY) atomicOr(&g.state[s].status, SOS_STATUS_NONFINITE);  // -inf: no owned column
    }
  }
}

// ------------------------------------------------------------------------------------------
// convergence bookkeeping
// ------------------------------------------------------------------------------------------
// ascending list of the active scenarios (warp 0 of a one-CTA kernel; deterministic)
__device__ __forceinline__ void build_active_list(const GridDev& g) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int base = 0;
  for (int s0 = 0; s0 < g.S; s0 += 32) {
    const int s = s0 + lane;
    const bool act = s < g.S && g.state[s].active;
    const unsigned mask = __ballot_sync(0xffffffffu, act);
    if (act) g.active_flat[base + __popc(mask & ((1u << lane) - 1u))] = s;
    base += __popc(mask);
  }
  if (lane == 0) *g.n_active = base;
}

// ratio_part != nullptr: the fused order kernel left one {TOA, surface} maximum per strip; reduce them first
__global__ void converge_kernel(const GridDev g, int order_arg, int* order_counter, const double* __restrict__ ratio_part,
                                int nstrips, int* strip_ticket) {
  __shared__ int order_s;
  if (threadIdx.x == 0) {
    // order_arg < 0: take the order number from the device-side counter (CUDA-graph replays cannot
    // change kernel arguments); the counter always tracks the last order handled
    order_s = order_arg >= 0 ? order_arg : *order_counter + 1;
    *order_counter = order_s;
    if (strip_ticket) *strip_ticket = 0;
  }
  __syncthreads();
  const int order = order_s;
  for (int s = threadIdx.x; s < g.S; s += blockDim.x) {
    ScenState& st = g.state[s];
    if (st.active) {
      if (ratio_part) {
        double r0 = -INFINITY, r1 = -INFINITY;
        for (int k = 0; k < nstrips; ++k) {
          r0 = fmax(r0, ratio_part[(static_cast<size_t>(s) * nstrips + k) * 2]);
          r1 = fmax(r1, ratio_part[(static_cast<size_t>(s) * nstrips + k) * 2 + 1]);
        }
        st.ratio_toa = r0;
        st.ratio_surf = r1;
      }
      st.n_orders = order;
      const double r = fmax(st.ratio_toa, st.ratio_surf);
      if (!(r >= g.scen[s].threshold)) st.active = 0;
    }
  }
  __syncthreads();
  build_active_list(g);
}

// ratios with I_n := 1 (the reference initialises In = ones before the loop, :306-309)
__global__ void reset_kernel(const GridDev g, const double* __restrict__ I1, int* order_counter) {
  const int s = blockIdx.x;
  __shared__ double sc[32];
  const int L = g.L, M = g.M, N = g.N, ld = g.ld;
  const double* __restrict__ I = I1 + static_cast<size_t>(s) * L * ld;
  double r0 = -INFINITY, r1 = -INFINITY;
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    if (m >= M) { const double r = 1.0 / I[m]; if (!isnan(r)) r0 = fmax(r0, r); }
    else { const double r = 1.0 / I[static_cast<size_t>(L - 1) * ld + m]; if (!isnan(r)) r1 = fmax(r1, r); }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r0 = fmax(r0, __shfl_xor_sync(0xffffffffu, r0, o));
    r1 = fmax(r1, __shfl_xor_sync(0xffffffffu, r1, o));
  }
  if (lane == 0) sc[warp] = r0;
  __syncthreads();
  if (threadIdx.x == 0) { double r = -INFINITY; for (int w = 0; w < nwarps; ++w) r = fmax(r, sc[w]); g.state[s].ratio_toa = r; }
  __syncthreads();
  if (lane == 0) sc[warp] = r1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = -INFINITY; for (int w = 0; w < nwarps; ++w) r = fmax(r, sc[w]);
    ScenState& st = g.state[s];
    st.ratio_surf = r;
    st.n_orders = 1;
    st.status = 0;
    if (s == 0) *order_counter = 1;
    st.active = (fmax(st.ratio_toa, r) >= g.scen[s].threshold) ? 1 : 0;
  }
}

// ratios <-> caller buffer [S][2] (mu-block sharding reduces them across ranks with a MAX all-reduce)
__global__ void ratios_kernel(const GridDev g, double* buf, int set) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.S) return;
  if (set) { g.state[s].ratio_toa = buf[2 * s]; g.state[s].ratio_surf = buf[2 * s + 1]; }
  else { buf[2 * s] = g.state[s].ratio_toa; buf[2 * s + 1] = g.state[s].ratio_surf; }
}

__global__ void count_active_kernel(const GridDev g, int* strip_ticket) {
  if (threadIdx.x == 0 && strip_ticket) *strip_ticket = 0;
  build_active_list(g);
}

}  // namespace sossweep
