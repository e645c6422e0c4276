// Closed-form first order of scattering.
//   three regions : SOS_Aer_main_specular.py:104-292 (the Lambertian driver shares it after "repair A",
//                   SURVEY.md 8c)
//   single layer  : I1_NumInt, SOS_Aer_I1_In.py:13-58
//
// Every general column is   carry * e^{(tau_t - tau_b)/mu}
//                         + mu0/(mu0+mu) C[m]      F0/(4pi) (e^{-tau_t/mu0} - e^{-tau_d/mu0} e^{(tau_t-tau_d)/mu})
//                         + mu0/(mu0-mu) C[mirror] S /(4pi) (e^{-(T-tau_t)/mu0} - e^{-(T-tau_s)/mu0} e^{(tau_t-tau_s)/mu})
// with region-dependent (tau_b, tau_d, tau_s) -- the same expression for both hemispheres because
// -(tau_b - tau_t)/mu == (tau_t - tau_b)/mu.  The carries are the values of the same closed form at the
// region boundary rows, so each thread first re-evaluates its column's short carry chain (<= 5 values)
// and then fills its rows; no inter-thread dependency, one launch.
#pragma once
#include "common.cuh"

namespace sosfirst {

struct Col {
  double mu, mu0, F0q, Sq, T;
  double Cm_atm, Cmir_atm, Cm_mix, Cmir_mix;
  bool down, special;
};

// value of the closed form at optical depth tt for region parameters (carry, tau_b, tau_d, tau_s)
__device__ __forceinline__ double i1_value(const Col& c, bool mix, double tt, double carry, double tau_b, double tau_d,
                                           double tau_s) {
  const double Cm = mix ? c.Cm_mix : c.Cm_atm;
  const double Cr = mix ? c.Cmir_mix : c.Cmir_atm;
  const double e0 = exp(-tt / c.mu0);
  const double es = exp(-(c.T - tt) / c.mu0);
  double direct, surf;
  if (c.down && c.special) {  // |mu + mu0| < 1e-4 (:133-140)
    direct = Cm * c.F0q * e0 * (tt - tau_d) / c.mu0;
  } else {
    direct = (c.mu0 / (c.mu0 + c.mu)) * Cm * c.F0q * (e0 - exp(-tau_d / c.mu0) * exp((tt - tau_d) / c.mu));
  }
  if (!c.down && c.special) {  // |mu - mu0| < 1e-4 (:225-233)
    surf = Cr * c.Sq * es * (tau_s - tt) / c.mu0;
  } else {
    surf = (c.mu0 / (c.mu0 - c.mu)) * Cr * c.Sq * (es - exp(-(c.T - tau_s) / c.mu0) * exp((tt - tau_s) / c.mu));
  }
  double v = direct + surf;
  if (carry != 0.0) v = carry * exp((tt - tau_b) / c.mu) + v;
  return v;
}

// downward column value at row t (3 regions), chaining the carries from the top
__device__ double i1_down(const GridDev& g, const Col& c, const double* __restrict__ tau, int t, int* cached_region,
                          double* cached_carry) {
  // region of row t
  int k = 0;
  while (k + 1 < g.nreg && t >= g.rstart[k + 1]) ++k;
  double carry = 0.0;
  if (*cached_region == k) {
    carry = *cached_carry;
  } else {
    for (int r = 1; r <= k; ++r) {
      const int cb = g.rstart[r] - 1;  // carry row of region r
      const int rp = r - 1;
      const double tb = rp == 0 ? 0.0 : tau[g.rstart[rp] - 1];
      const double ts = rp == 0 ? 0.0 : tau[g.rstart[rp]];
      carry = i1_value(c, rp == 1, tau[cb], carry, tb, tb, ts);
    }
    *cached_region = k;
    *cached_carry = carry;
  }
  const double tb = k == 0 ? 0.0 : tau[g.rstart[k] - 1];
  const double ts = k == 0 ? 0.0 : tau[g.rstart[k]];
  return i1_value(c, k == 1, tau[t], carry, tb, tb, ts);
}

__global__ void __launch_bounds__(128)
first_order_regions_kernel(const GridDev g, const double* __restrict__ Cs /*[S][2][N]*/, double* __restrict__ I1,
                           int rows_per_block) {
  const int s = blockIdx.z;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.N) return;
  const int L = g.L, M = g.M, N = g.N, ld = g.ld;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const sos_scenario sc = g.scen[s];
  const double* __restrict__ Catm = Cs + static_cast<size_t>(s) * 2 * N;
  const double* __restrict__ Cmix = Catm + N;
  double* __restrict__ out = I1 + static_cast<size_t>(s) * L * ld;
  const int ta = blockIdx.y * rows_per_block;
  const int tb = min(L, ta + rows_per_block);
  const double PI = 3.14159265358979323846;
  const double mu0 = sc.mu0;
  const double F0 = PI / mu0;
  const double T = sc.tauStar_tot;
  const double q = 1.0 / (4.0 * PI);
  const double S = F0 * sc.grd_alb * exp(-T / mu0);

  if (m == M - 1 || m == M) {
    // mu = 0-: C[M-1] F0/(4pi) e^{-tau/mu0} + C[M] S/(4pi) e^{-(T-tau)/mu0}   (:124-131); mu = 0+ mirrored (:217-224)
    const int mir = N - 1 - m;
    for (int t = ta; t < tb; ++t) {
      int k = 0;
      while (k + 1 < g.nreg && t >= g.rstart[k + 1]) ++k;
      const double* C = (k == 1) ? Cmix : Catm;
      const double tt = tau[t];
      out[static_cast<size_t>(t) * ld + m] = (mu0 / (mu0 + g.mu[m])) * C[m] * (F0 * q) * exp(-tt / mu0) +
                                             (mu0 / (mu0 - g.mu[m])) * C[mir] * (S * q) * exp(-(T - tt) / mu0);
    }
    return;
  }

  Col c;
  c.mu0 = mu0; c.F0q = F0 * q; c.Sq = S * q; c.T = T;
  if (m < M - 1) {
    c.mu = g.mu[m];
    c.down = true;
    c.special = fabs(c.mu + mu0) < SOS_MU0_TOLERANCE;
    c.Cm_atm = Catm[m]; c.Cmir_atm = Catm[N - 1 - m];
    c.Cm_mix = Cmix[m]; c.Cmir_mix = Cmix[N - 1 - m];
    int creg = -1; double ccar = 0.0;
    for (int t = ta; t < tb; ++t) out[static_cast<size_t>(t) * ld + m] = i1_down(g, c, tau, t, &creg, &ccar);
    return;
  }

  // ---- upward column: needs the mirror downward column at the surface row first ----
  const int mir = N - 1 - m;
  Col d;
  d.mu0 = mu0; d.F0q = c.F0q; d.Sq = c.Sq; d.T = T;
  d.mu = g.mu[mir];
  d.down = true;
  d.special = fabs(d.mu + mu0) < SOS_MU0_TOLERANCE;
  d.Cm_atm = Catm[mir]; d.Cmir_atm = Catm[m];
  d.Cm_mix = Cmix[mir]; d.Cmir_mix = Cmix[m];
  int creg = -1; double ccar = 0.0;
  const double surf_down = i1_down(g, d, tau, L - 1, &creg, &ccar);

  c.mu = g.mu[m];
  c.down = false;
  c.special = fabs(c.mu - mu0) < SOS_MU0_TOLERANCE;
  c.Cm_atm = Catm[m]; c.Cmir_atm = Catm[mir];
  c.Cm_mix = Cmix[m]; c.Cmir_mix = Cmix[mir];

  // carries of the upward chain, bottom region first:
  //   last region : carry = rho * I1[L-1, mirror], tau_b = tau_d = tau[L-1], tau_s = T   (:206-216)
  //   region k    : carry = I1[rstart[k+1], m],    tau_b = tau_d = tau[rstart[k+1]], tau_s = tau[rstart[k+1]-1]
  const int R = g.nreg;
  double carry_k[3];
  carry_k[R - 1] = sc.grd_alb * surf_down;
  for (int k = R - 2; k >= 0; --k) {
    const int row = g.rstart[k + 1];  // first row of the region below = carry row
    const int kb = k + 1;
    const double tbb = (kb == R - 1) ? tau[L - 1] : tau[g.rstart[kb + 1]];
    const double tss = (kb == R - 1) ? T : tau[g.rstart[kb + 1] - 1];
    carry_k[k] = i1_value(c, kb == 1, tau[row], carry_k[kb], tbb, tbb, tss);
  }
  for (int t = ta; t < tb; ++t) {
    int k = 0;
    while (k + 1 < R && t >= g.rstart[k + 1]) ++k;
    const double tbb = (k == R - 1) ? tau[L - 1] : tau[g.rstart[k + 1]];
    const double tss = (k == R - 1) ? T : tau[g.rstart[k + 1] - 1];
    out[static_cast<size_t>(t) * ld + m] = i1_value(c, k == 1, tau[t], carry_k[k], tbb, tbb, tss);
  }
}

// I1_NumInt (SOS_Aer_I1_In.py:13-58): one homogeneous layer above a black surface
__global__ void __launch_bounds__(128)
first_order_single_kernel(const GridDev g, const double* __restrict__ Cs /*[S][2][N], plane 0 = alb*P0*/,
                          double* __restrict__ I1, int rows_per_block) {
  const int s = blockIdx.z;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.N) return;
  const int L = g.L, M = g.M, N = g.N, ld = g.ld;
  const double* __restrict__ tau = g.tau + static_cast<size_t>(s) * L;
  const sos_scenario sc = g.scen[s];
  const double C = Cs[static_cast<size_t>(s) * 2 * N + m];
  double* __restrict__ out = I1 + static_cast<size_t>(s) * L * ld;
  const int ta = blockIdx.y * rows_per_block;
  const int tb = min(L, ta + rows_per_block);
  const double PI = 3.14159265358979323846;
  const double mu0 = sc.mu0, mu = g.mu[m], T = sc.tauStar_tot;
  const double k = C / (4.0 * PI);
  const double norm = PI / mu0;  // (:58)
  const double eS = exp(-T / mu0);
  for (int t = ta; t < tb; ++t) {
    const double tt = tau[t];
    const double e0 = exp(-tt / mu0);
    double v;
    if (m == M - 1 || m == M) {
      v = k * (mu0 / (mu0 + mu)) * e0;                                  // (:39,:50)
    } else if (m < M - 1) {
      if (fabs(mu + mu0) < SOS_MU0_TOLERANCE) v = k * e0 * tt / mu0;    // (:41-43)
      else v = (mu0 / (mu0 + mu)) * k * (e0 - exp(tt / mu));            // (:34-37)
    } else {
      v = (mu0 / (mu0 + mu)) * k * (e0 - eS * exp(-(T - tt) / mu));     // (:54-55)
    }
    out[static_cast<size_t>(t) * ld + m] = v * norm;
  }
}

}  // namespace sosfirst
