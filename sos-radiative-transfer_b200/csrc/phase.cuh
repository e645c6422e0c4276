// Azimuth-averaged phase functions on the device: P0(mu, mu0) and P(mu, mu') of the
// analytic / tabulated families of SOS_Aer_phase_func.py:68-292 -- Rayleigh (:97), Henyey-Greenstein
// (:158), tabulated FWC with the reference's searchsorted + linear interpolation (:202-236).  Same
// quadrature as the reference: 25 azimuth nodes on [0, pi], both half-rings, composite trapezoid; the raw
// matrix is symmetric and every COLUMN is then normalised to trapz(P[:, n], mu) = 4 (:131).
// The reference spends 89-112 s per matrix at N = 1002 in a triple Python loop; this is two launches.
#pragma once
#include "common.cuh"

namespace sosphase {

constexpr int NPHI = 25;
enum Family { RAYLEIGH = 0, HG = 1, TABLE = 2 };

struct PhaseArgs {
  int family;
  double g;
  const double* tab_x;  // TABLE: abscissae (ascending), device
  const double* tab_y;
  int tab_n;
  double phi[NPHI];     // linspace(0, pi, 25) as computed by the host (NumPy), so that nodes are identical
  double cphi[NPHI];    // cos(0 - phi)
};

__device__ __forceinline__ double phase_value(const PhaseArgs& a, double c) {
  if (a.family == RAYLEIGH) return 0.75 * (1.0 + c * c);
  if (a.family == HG) {
    const double g = a.g;
    return (1.0 - g * g) / pow(1.0 + g * g - 2.0 * g * c, 1.5);
  }
  // np.clip + np.searchsorted(side='left') + linear interpolation (:214-236)
  c = fmin(fmax(c, -1.0), 1.0);
  int lo = 0, hi = a.tab_n;  // first index with tab_x[idx] >= c
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a.tab_x[mid] < c) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return a.tab_y[0];
  if (lo >= a.tab_n) return a.tab_y[a.tab_n - 1];
  const double x0 = a.tab_x[lo - 1], x1 = a.tab_x[lo];
  const double w = (c - x0) / (x1 - x0);
  return a.tab_y[lo - 1] + w * (a.tab_y[lo] - a.tab_y[lo - 1]);
}

// trapz over phi of f(-(cc + ss cos phi)) + f(-(cc - ss cos phi))
__device__ __forceinline__ double ring_integral(const PhaseArgs& a, double cc, double ss) {
  double sum = 0.0;
  double prev = 0.0;
#pragma unroll 5
  for (int k = 0; k < NPHI; ++k) {
    const double x = ss * a.cphi[k];
    const double y = phase_value(a, -(cc + x)) + phase_value(a, -(cc - x));
    if (k > 0) sum += (a.phi[k] - a.phi[k - 1]) * (y + prev) / 2.0;
    prev = y;
  }
  return sum;
}

// raw symmetric matrix R[m][n] (only m >= n is computed by the reference and mirrored; cc and ss are
// symmetric in (m, n) bit for bit, so computing every entry gives the same matrix)
__global__ void __launch_bounds__(256)
phase_raw_kernel(const PhaseArgs a, const double* __restrict__ mu, int N, double* __restrict__ P, int ldp) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (m >= N || n >= N) return;
  const double um = mu[m], un = mu[n];
  const double cc = um * un;
  const double ss = sqrt(1.0 - un * un) * sqrt(1.0 - um * um);
  const double PI = 3.14159265358979323846;
  P[static_cast<size_t>(m) * ldp + n] = ring_integral(a, cc, ss) / (2.0 * PI);
}

// column integrals trapz(R[:, n], mu): one thread per column, coalesced over n
__global__ void __launch_bounds__(128)
phase_colint_kernel(const double* __restrict__ mu, int N, const double* __restrict__ P, int ldp, double* __restrict__ colint) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s = 0.0;
  double prev = P[n];
  for (int m = 1; m < N; ++m) {
    const double cur = P[static_cast<size_t>(m) * ldp + n];
    s += (mu[m] - mu[m - 1]) * (cur + prev) / 2.0;
    prev = cur;
  }
  colint[n] = s;
}

__global__ void __launch_bounds__(256)
phase_normalise_kernel(int N, double* __restrict__ P, int ldp, const double* __restrict__ colint) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (m >= N || n >= N) return;
  const size_t i = static_cast<size_t>(m) * ldp + n;
  P[i] = 4.0 * P[i] / colint[n];
}

// P0(mu, mu0): one CTA; (:92-105)
__global__ void __launch_bounds__(256)
phase_p0_kernel(const PhaseArgs a, const double* __restrict__ mu, int N, double mu0, double* __restrict__ P0) {
  __shared__ double part[256];
  const double PI = 3.14159265358979323846;
  const double s0 = sqrt(1.0 - mu0 * mu0);
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    const double um = mu[m];
    P0[m] = ring_integral(a, um * mu0, s0 * sqrt(1.0 - um * um)) / (4.0 * PI);
  }
  __syncthreads();
  double s = 0.0;
  for (int m = 1 + threadIdx.x; m < N; m += blockDim.x) s += (mu[m] - mu[m - 1]) * (P0[m] + P0[m - 1]) / 2.0;
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  const double tot = part[0];
  __syncthreads();
  for (int m = threadIdx.x; m < N; m += blockDim.x) P0[m] = P0[m] / tot * 2.0;
}

}  // namespace sosphase
