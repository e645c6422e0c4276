// One scattering order in ONE pass over memory: the layer sweeps of In_NumInt (SOS_Aer_I1_In.py:77-130; inlined
// three-region form SOS_Aer_main_specular.py:327-449, Lambert surface SOS_Aer_main_lambertian.py:399,401), the
// accumulation I += I_n (:454-456), the convergence ratios (:309) and -- for rows whose contraction operand is the
// rank <= 2 molecular one (gemm_lowrank.cuh) -- the source function itself (:315-323), fused.
//
// sweep.cuh evaluates the exp(-dtau/|mu|) recurrence as a chunked scan (local aggregates, carry chain, apply, zone
// fix-up: four launches, J read twice, 40 B per element).  For a BATCH of scenarios there is enough parallelism
// without cutting the layer axis, so here one CTA owns a "strip": 128 downward columns of one scenario AND their 128
// mirror columns, and streams through all L rows twice with the running D / U in registers --
//   * down pass (rows 0 .. L-1), thread j <-> column d = M-1-(128 k + j): no carries, no look-back;
//   * the specular coupling is thread-local: the mirror column u = N-1-d belongs to the same thread, so the seed
//     rho * I_n[L-1, d] never leaves the register file (Lambert: one partial sum per strip, exchanged through L2);
//   * up pass (rows L-1 .. 0) on the mirror block, re-seeded from the blended boundary rows;
//   * strip 0 holds every column the reference treats specially next to mu = 0 (windowed / Taylor columns,
//     extrapolation targets and sources, the find-first blend): they are finished in shared memory before the tile
//     is written, so no second kernel touches the fields.
// Row tiles ([R rows][128 columns] of J and of I) arrive by TMA (3-D tensor maps [S][L][N]: out-of-range rows and
// columns are zero-filled on load and clipped on store, so ragged L, M need no special code) into an NS-stage
// mbarrier ring; the finished I_n and I tiles leave by TMA bulk stores from the same buffers.  Algorithmic traffic
// when J is read: J 8 + I_n 8 + I 16 = 32 B per element -- the minimum.
// Two hardware rules shape the column layout (measured with tools/tma_probe.cu on B200: both violations raise "illegal
// instruction"): the first element of a box must sit on a 16-byte boundary -- an EVEN column for doubles -- and a bulk
// STORE must not start at a negative coordinate (loads may; both may overhang the far edge).  N = 2M is even, so mirror
// blocks are even-aligned as soon as the down blocks are; for odd M (the reference's 501) the blocks therefore stop
// one column short of mu = 0 on either side -- d = M-2-(128 k + j), u = M+1+128 k + j -- and the pair of columns
// mu = 0-/0+ (M-1, M: no recurrence on either, SOS_Aer_I1_In.py:100,124-127) is carried by eight threads of strip 0
// through plain loads and stores, prefetched one stage ahead.  The strip at the mu = -1 end starts at column 0 and
// uses tensor maps cut off at its last column instead of a negative start.
//
// Generated source.  On rows that use the molecular operand alone (every row outside the aerosol layer,
// SOS_Aer_main_specular.py:323), A = Us Vt with Vt = [1; mu^2] (gemm_lowrank.cuh), so
//     J[t, m] = coef * (c0[t] + c1[t] mu_m^2),   c_r[t] = sum_k I_{n-1}[t, k] Us[k][r]
// and the two numbers c_r[t] per row are all the next order needs from this one.  The kernel therefore emits, for
// every finished row, the partial projections of its own columns (one m8n8k4 DMMA chain per warp: one slot per strip,
// half and warp, summed by the readers in a fixed tree: deterministic) and REBUILDS J from them in the next order instead of reading it: neither J
// nor I_n is ever written for those rows.  Per element and order that leaves the I read-modify-write, 16 B, on 746 of
// the 800 default rows; the dense (aerosol) rows keep the 32 B path and the DMMA contraction.
//
// Windowed columns (|mu| < 0.01, SOS_Aer_In_limit.py:96-107: trapezoid over tau' >= tau_t - 5|mu| inside the region).
// With D the same recurrence restarted at the region's first row, the window integral is
//     D_t - exp((tau_t - tau_k0)/mu) D_k0,      k0 = first row of the window,
// (the intervals above k0, decayed to t, cancel exactly), so the column costs one recurrence step, one exp and one
// look-up per row instead of a fresh sum over up to 5|mu|/dtau rows.  k0 depends only on tau and mu and is tabulated
// on the host with the reference's own rounding (tau_t - 5*abs(mu) as two operations); D_k0 is fetched from a small
// history array one stage ahead of its use.
#pragma once
#include "common.cuh"
#include "gemm_f64.cuh"
#include "sweep.cuh"

namespace sosstrip {

using sosgemm::mbar_arrive;
using sosgemm::mbar_expect_tx;
using sosgemm::mbar_init;
using sosgemm::mbar_wait;
using sosgemm::smem_u32;
using sossweep::exp_small;

constexpr int W = 128;          // columns per strip half = threads per CTA
constexpr int THREADS = 128;          // consumer threads: one per column of a strip half
constexpr int CTA_THREADS = THREADS + 32;  // + one producer warp that owns every TMA operation
constexpr int MAX_SMALL = 16;   // windowed / Taylor columns (|mu| < 0.01, without mu = 0-) a plan may have
constexpr int MAX_STRIPS = 16;  // M <= 2048
constexpr int STRIP_MIN_CTAS = 2;

struct StripParams {
  GridDev g;
  CUtensorMap map_J, map_In, map_I, map_S;  // [S][L][N] (strides ld, L*ld), box {W, R, 1}
  CUtensorMap lo_J, lo_In, lo_I, lo_S;      // the same fields cut off after the columns of the strip at the mu = -1 end
  const double* J;                          // raw pointers for the mu = 0 column pair (odd M)
  double* In;
  double* I;
  double* saved;
  const double* tau_pad;                    // [S][Lp]: tau with rows padded to whole stages (last value repeated)
  int Lp;
  const int* active;                        // compacted ids of the scenarios still iterating ([*g.n_active])
  int* ticket;                              // work counter, zeroed before the launch
  int nstrips, nslots;                      // strips per scenario; projection slots per row = 2 * nstrips
  int has_saved;                            // also store I_n into map_S (I_saved of SOS_Aer_main_specular.py:458)
  int store_all;                            // store I_n on every row (otherwise only where the next contraction is dense)
  const double* Ut[SOS_MAX_PHASE];          // low-rank factors of the operands ([4][ldr], gemm_lowrank.cuh); rank 0 = dense
  const double* Vt[SOS_MAX_PHASE];
  int rank[SOS_MAX_PHASE];
  int ldr;
  const double* proj_in;                    // [S][L][nslots][2] projections of I_{n-1} (read when J is generated)
  double* proj_out;                         //                    ... of I_n
  double* ratio_part;                       // [S][nstrips][2] {TOA, surface} ratio maxima of each strip
  double* lam_part;                         // [S][nstrips] Lambert partial sums
  unsigned* lam_flag;                       // [S][nstrips] epoch stamps of lam_part
  unsigned epoch;                           // unique per launch
  double* dhist;                            // [S][L][MAX_SMALL] region-restarted D of the windowed columns
  const int* k0tab;                         // [S][L][nsc] first row of every window (host-built)
  int nsc;                                  // small columns first_small .. M-2
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// barrier over the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

template <int R>
struct StageLayout {
  // [J / I_n tile][I tile][projections of the stage's rows][tau of the stage's rows]
  static constexpr int TILE = R * W * 8;
  __host__ __device__ static int bytes(int nslots) { return (2 * TILE + R * nslots * 16 + R * 8 + 127) / 128 * 128; }
};
template <int R, int NS>
__host__ __device__ inline int strip_smem_bytes(int nslots) { return NS * StageLayout<R>::bytes(nslots) + 128; }

__device__ __forceinline__ int region_of(const GridDev& g, int t) {
  int k = 0;
  while (k + 1 < g.nreg && t >= g.rstart[k + 1]) ++k;
  return k;
}

enum ColKind { K_INVALID = 0, K_STD = 1, K_M1 = 2, K_TAYLOR = 3, K_WINDOW = 4 };

// exp(x) for |x| <= 2^-10 (x^5/5! < 7e-18 relative): what almost every scan step of a thin atmosphere needs
constexpr double kTinyArg = 0.0009765625;
__device__ __forceinline__ double exp_tiny(double x) {
  double p = sossweep::kExpTaylor[4];
  p = fma(p, x, sossweep::kExpTaylor[5]);
  p = fma(p, x, sossweep::kExpTaylor[6]);
  p = fma(p, x, sossweep::kExpTaylor[7]);
  return fma(p, x, sossweep::kExpTaylor[7]);
}

constexpr int PROJ_WARPS = THREADS / 32;  // every warp of a strip half writes its own projection slot

template <int R, int NS>
__global__ void __launch_bounds__(CTA_THREADS, STRIP_MIN_CTAS) order_strip_kernel(const __grid_constant__ StripParams p) {
  static_assert(R == 8, "the projection uses one m8n8k4 tile of rows per stage");
  static_assert(R * MAX_SMALL <= THREADS && 16 * R <= THREADS, "stage shape");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[NS];   // TMA loads of a stage have landed (producer -> consumers)
  __shared__ uint64_t ready_bar[NS];  // the stage's tiles are final and fenced (consumers -> producer)
  __shared__ int s_ticket;
  __shared__ int s_istar;
  __shared__ double s_cj[2][R][2];
  __shared__ double s_us[2][W][2];  // Us of this strip's down / up columns (0 outside the strip)
  __shared__ double s_lkv[R][MAX_SMALL];
  __shared__ int s_lkk[R][MAX_SMALL];
  __shared__ double s_dh[R][MAX_SMALL];
  __shared__ double s_row[W + 1];
  __shared__ double s_red[8];
  __shared__ double s_pI[R], s_pJ[R], s_pV[R];  // mu = 0 column of the current pass (odd M): old I, J (dense rows), I_n

  const GridDev& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = g.L, M = g.M, N = g.N;
  const int o = M & 1;  // odd M: the blocks stop one column short of mu = 0 (see the header)
  const int nslots = p.nslots;
  const int stage_bytes = StageLayout<R>::bytes(nslots);
  const int total = *g.n_active * p.nstrips;
  const size_t ld = g.ld;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&ready_bar[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t seq = 0;  // stages consumed by this CTA so far (ring position and mbarrier phase)
  const int nst = (L + R - 1) / R, nseq = 2 * nst;
  const int side_r = tid - (W - R);  // threads W-R .. W-1 also carry row side_r of the mu = 0 column pair

  for (;;) {
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int tk = s_ticket;
    __syncthreads();
    if (tk >= total) break;
    const int ai = tk / p.nstrips, k = tk - ai * p.nstrips;
    const int s = p.active[ai];
    const sos_scenario sc = g.scen[s];
    const double* __restrict__ tau_g = g.tau + static_cast<size_t>(s) * L;
    const int op = sc.phase_atm;
    const int rk = p.rank[op];                        // 0: J is read from memory on every row
    const int a0 = (g.nreg == 3) ? g.rstart[1] : L;   // rows [a0, a1) keep the dense contraction (aerosol layer)
    const int a1 = (g.nreg == 3) ? g.rstart[2] : L;
    const bool lowstrip = M - o - W * (k + 1) < 0;    // the strip at the mu = -1 end: starts at column 0, clipped maps
    const int dcol0 = lowstrip ? 0 : M - o - W * (k + 1), ucol0 = M + o + W * k;
    const int dend = M - o - W * k;                   // one past this strip's last downward column
    const int d = M - 1 - o - (W * k + tid), u = ucol0 + tid;
    const bool vd = d >= 0, vu = u < N;
    const int ci = vd ? d - dcol0 : tid;              // tile column of d (W-1-tid except in the low strip, whose idle threads sit past its columns)
    const double mud = vd ? g.mu[d] : -1.0, muu = vu ? g.mu[u] : 1.0;
    const double imud = 1.0 / mud, imuu = 1.0 / muu;
    double vd0 = 0.0, vd1 = 0.0, vu0 = 0.0, vu1 = 0.0;
    if (rk > 0 && tid < THREADS) {
      const double* __restrict__ Vt = p.Vt[op];
      const double* __restrict__ Ut = p.Ut[op];
      if (vd) { vd0 = Vt[d]; vd1 = Vt[p.ldr + d]; }
      if (vu) { vu0 = Vt[u]; vu1 = Vt[p.ldr + u]; }
      // projection operands by TILE column (tid = tile column here)
      const int cdn = dcol0 + tid, cup = ucol0 + tid;
      s_us[0][tid][0] = (cdn < dend) ? Ut[cdn] : 0.0;
      s_us[0][tid][1] = (cdn < dend) ? Ut[p.ldr + cdn] : 0.0;
      s_us[1][tid][0] = (cup < N) ? Ut[cup] : 0.0;
      s_us[1][tid][1] = (cup < N) ? Ut[p.ldr + cup] : 0.0;
    }
    int kd = K_INVALID;
    if (vd) kd = (d == M - 1) ? K_M1 : (fabs(mud) >= SOS_MU_THRESHOLD ? K_STD : (fabs(mud) < SOS_MU_VERY_SMALL ? K_TAYLOR : K_WINDOW));
    const int csm = d - g.first_small;                // index among the small columns (valid for K_TAYLOR / K_WINDOW)
    int wmaxs = 0;
    for (int r = 0; r < g.nreg; ++r) wmaxs = max(wmaxs, sc.extrap_width[r]);
    const bool side = (o == 1 && k == 0);             // this CTA carries the column pair M-1, M outside its tiles
    const bool side_thread = side && side_r >= 0;
    double vM0 = 0.0, vM1 = 0.0, uM1a = 0.0, uM1b = 0.0, uMa = 0.0, uMb = 0.0;  // Vt[:, M]; Us[M-1], Us[M]
    if (side_thread && rk > 0) { vM0 = p.Vt[op][M]; vM1 = p.Vt[op][p.ldr + M]; }
    if (side && rk > 0 && warp == 0) { uM1a = p.Ut[op][M - 1]; uM1b = p.Ut[op][p.ldr + M - 1]; uMa = p.Ut[op][M]; uMb = p.Ut[op][p.ldr + M]; }

    const uint32_t seq0 = seq;
    auto stage_ptr = [&](uint32_t gq) { return smem + static_cast<size_t>(gq % NS) * stage_bytes; };
    auto stage_rows = [&](int q2, bool& down2, int& t02, int& rows2) {
      down2 = q2 < nst;
      const int st2 = down2 ? q2 : nseq - 1 - q2;
      t02 = st2 * R;
      rows2 = min(R, L - t02);
    };
    auto issue = [&](int q2) {  // thread 0: start the loads of sequence step q2 of this work item
      bool down2; int t02, rows2;
      stage_rows(q2, down2, t02, rows2);
      const uint32_t gq2 = seq0 + q2;
      uint8_t* sp = stage_ptr(gq2);
      uint64_t* bar = &full_bar[gq2 % NS];
      const bool dense2 = rk == 0 || (t02 < a1 && t02 + rows2 > a0);
      const bool gen2 = rk > 0 && (t02 < a0 || t02 + rows2 > a1);
      const uint32_t bytes = StageLayout<R>::TILE + (dense2 ? StageLayout<R>::TILE : 0) + (gen2 ? rows2 * nslots * 16 : 0) + R * 8;
      mbar_expect_tx(bar, bytes);
      const int c0 = down2 ? dcol0 : ucol0;
      const bool lo = down2 && lowstrip;
      tma_load_3d(smem_u32(sp + StageLayout<R>::TILE), lo ? &p.lo_I : &p.map_I, bar, c0, t02, s);
      if (dense2) tma_load_3d(smem_u32(sp), lo ? &p.lo_J : &p.map_J, bar, c0, t02, s);
      if (gen2)
        bulk_load_1d(smem_u32(sp + 2 * StageLayout<R>::TILE), p.proj_in + (static_cast<size_t>(s) * L + t02) * nslots * 2,
                     rows2 * nslots * 16, bar);
      bulk_load_1d(smem_u32(sp + 2 * StageLayout<R>::TILE + R * nslots * 16), p.tau_pad + static_cast<size_t>(s) * p.Lp + t02, R * 8, bar);
    };
    // c_r[t] = coef * (sum over slots of the partial projections), for the rows of sequence step q2, into s_cj[q2 & 1].
    // All 128 threads: (row, component, eighth of the slots), fixed summation tree -> deterministic.
    auto coefficients = [&](int q2) {
      bool down2; int t02, rows2;
      stage_rows(q2, down2, t02, rows2);
      if (!(rk > 0 && (t02 < a0 || t02 + rows2 > a1))) return;
      const uint32_t gq2 = seq0 + q2;
      mbar_wait(&full_bar[gq2 % NS], (gq2 / NS) & 1);
      const double* __restrict__ pr2 = reinterpret_cast<const double*>(stage_ptr(gq2) + 2 * StageLayout<R>::TILE);
      const int r = tid >> 4, c = (tid >> 3) & 1, part = tid & 7;
      double sum = 0.0;
      if (r < rows2)
        for (int j = part; j < nslots; j += 8) sum += pr2[(r * nslots + j) * 2 + c];
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (part == 0) s_cj[q2 & 1][r][c] = sc.coef_atm * sum;
    };
    if (tid >= THREADS) {
      // ===================== producer warp: every TMA load and store of this work item =====================
      // (a consumer thread that issued them would put the bulk store's shared-memory read latency on the critical path
      //  of its whole CTA: measured 6.7 us per 8-row stage with thread 0 as issuer)
      if (tid == THREADS) {
        bulk_wait_read<0>();  // the previous work item's stores have left the ring
        for (int q2 = 0; q2 < NS - 1 && q2 < nseq; ++q2) issue(q2);
        for (int q = 0; q < nseq; ++q) {
          bool down; int t0, rows;
          stage_rows(q, down, t0, rows);
          const uint32_t gq = seq0 + q;
          uint8_t* sp = stage_ptr(gq);
          const bool hasdense = rk == 0 || (t0 < a1 && t0 + rows > a0);
          mbar_wait(&ready_bar[gq % NS], (gq / NS) & 1);
          const int c0 = down ? dcol0 : ucol0;
          const bool lo = down && lowstrip;
          if (hasdense || p.store_all) tma_store_3d(lo ? &p.lo_In : &p.map_In, smem_u32(sp), c0, t0, s);
          tma_store_3d(lo ? &p.lo_I : &p.map_I, smem_u32(sp + StageLayout<R>::TILE), c0, t0, s);
          if (p.has_saved) tma_store_3d(lo ? &p.lo_S : &p.map_S, smem_u32(sp), c0, t0, s);
          bulk_commit();
          if (q + NS - 1 < nseq) {
            // the buffer of step q-1 is refilled: its consumers are done (they have signalled step q) and its stores
            // have read it (at most this step's group is still pending)
            bulk_wait_read<1>();
            issue(q + NS - 1);
          }
        }
      }
      seq += nseq;
      continue;
    }
    if (k == 0 && p.nsc > 0 && tid < min(R, L) * p.nsc) {  // windows of the first stage start inside it
      const int r = tid / p.nsc, c = tid - r * p.nsc;
      s_lkk[r][c] = p.k0tab[(static_cast<size_t>(s) * L + r) * p.nsc + c];
      s_lkv[r][c] = 0.0;
    }
    if (side_thread && side_r < min(R, L)) s_pI[side_r] = p.I[(static_cast<size_t>(s) * L + side_r) * ld + (M - 1)];
    consumer_sync();   // (thread 0 has issued step 0 before anyone waits for it)
    coefficients(0);
    consumer_sync();

    // running state of the two recurrences
    double D = 0.0, Jp = 0.0, tp = tau_g[0];
    double Dr = 0.0;                    // region-restarted recurrence of a windowed column
    int regd = 0, r0d = 0;              // region of the down pass (slow path)
    double U = 0.0, Jn = 0.0, tn = tau_g[L - 1];
    double lastJ = 0.0;                 // I_n[t, mu = 0+] = J of the last row finished (thread 0 of strip 0)
    double seed_src = 0.0;

    for (int q = 0; q < nseq; ++q) {
      bool down; int t0, rows;
      stage_rows(q, down, t0, rows);
      const int st = t0 / R;
      const uint32_t gq = seq0 + q;
      uint8_t* sp = stage_ptr(gq);
      double* __restrict__ Jt = reinterpret_cast<double*>(sp);
      double* __restrict__ It = reinterpret_cast<double*>(sp + StageLayout<R>::TILE);
      const double* __restrict__ pr = reinterpret_cast<const double*>(sp + 2 * StageLayout<R>::TILE);
      const double* __restrict__ ta = reinterpret_cast<const double*>(sp + 2 * StageLayout<R>::TILE + R * nslots * 16);
      const double (*cj)[2] = s_cj[q & 1];
      const bool hasdense = rk == 0 || (t0 < a1 && t0 + rows > a0);
      mbar_wait(&full_bar[gq % NS], (gq / NS) & 1);

      // coefficients of the NEXT step's generated source (published by the barrier that ends this step's main loop)
      if (q + 1 < nseq) coefficients(q + 1);

      // ---------------- strip 0: the mu = 0 column of this pass (odd M) ----------------
      if (k == 0) {
        if (side_thread && side_r < rows) {
          double v = 0.0;  // down: column M-1 is an extrapolation target (written by the fix-up below) or stays 0
          if (!down) {     // up: I_n[t, M] = J[t, M] (SOS_Aer_I1_In.py:100)
            const int t = t0 + side_r;
            v = (rk > 0 && (t < a0 || t >= a1)) ? fma(cj[side_r][1], vM1, cj[side_r][0] * vM0) : s_pJ[side_r];
          }
          s_pV[side_r] = v;
        }
        consumer_sync();  // also publishes s_lk* / s_pI / s_pJ written at the end of the previous step
      }

      // tau of the stage's rows; the largest step decides (warp-uniformly) which exp polynomial the scan uses
      double tcv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) tcv[r] = ta[r];

      if (down) {
        double dtmax = tcv[0] - tp;
#pragma unroll
        for (int r = 1; r < R; ++r) dtmax = fmax(dtmax, tcv[r] - tcv[r - 1]);
        const bool tiny = __all_sync(0xffffffffu, kd != K_STD || dtmax * fabs(imud) <= kTinyArg);
        if (kd == K_STD || kd == K_INVALID) {
          if (kd == K_STD) {
            double jv[R], xv[R], av[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const int t = t0 + r;
              const bool gen = rk > 0 && (t < a0 || t >= a1);
              jv[r] = gen ? fma(cj[r][1], vd1, cj[r][0] * vd0) : Jt[r * W + ci];
              xv[r] = (tcv[r] - (r ? tcv[r - 1] : tp)) * imud;
            }
            if (tiny) {
#pragma unroll
              for (int r = 0; r < R; ++r) av[r] = exp_tiny(xv[r]);
            } else {
#pragma unroll
              for (int r = 0; r < R; ++r) av[r] = exp_small(xv[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) xv[r] = (0.5 * xv[r]) * ((r ? jv[r - 1] : Jp) * av[r] + jv[r]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (r < rows) {
                D = (t0 + r == 0) ? 0.0 : D * av[r] - xv[r];
                Jt[r * W + ci] = D;
                It[r * W + ci] += D;
                Jp = jv[r];
                tp = tcv[r];
              }
            }
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) Jt[r * W + ci] = 0.0;  // idle threads of the low strip: keep the projection's operand finite
          }
        } else {
          // the few special columns next to mu = 0- (strip 0 only)
          for (int r = 0; r < rows; ++r) {
            const int t = t0 + r;
            const double tc = ta[r];
            while (regd + 1 < g.nreg && t >= g.rstart[regd + 1]) { ++regd; r0d = g.rstart[regd]; }
            const bool gen = rk > 0 && (t < a0 || t >= a1);
            const double jt = gen ? fma(cj[r][1], vd1, cj[r][0] * vd0) : Jt[r * W + ci];
            const double io = It[r * W + ci];
            const bool target = (M - 1 - d) < sc.extrap_width[regd];
            double val;
            if (kd == K_M1) {
              val = 0.0;  // extrapolation target, or 0 when nothing is extrapolated
            } else if (kd == K_TAYLOR) {  // -J + mu dJ/dtau (SOS_Aer_In_limit.py:79-93)
              const double slope = (t > r0d) ? (jt - Jp) / (tc - tp) : 0.0;
              val = -jt + mud * slope;
            } else {
              if (t == r0d) Dr = 0.0;
              else {
                const double dt = tc - tp;
                const double a = exp_small(dt * imud);
                Dr = Dr * a - (dt * 0.5) * (Jp * a + jt) * imud;
              }
              s_dh[r][csm] = Dr;
              p.dhist[(static_cast<size_t>(s) * L + t) * MAX_SMALL + csm] = Dr;
              const int k0 = s_lkk[r][csm];
              const double Dk = (k0 >= t0) ? s_dh[k0 - t0][csm] : s_lkv[r][csm];
              val = Dr - exp((tc - tau_g[k0]) / mud) * Dk;
              if (!isfinite(val)) val = -jt;  // (:104-105)
            }
            Jt[r * W + ci] = val;
            if (kd != K_M1 && !target) It[r * W + ci] = io + val;
            Jp = jt;
            tp = tc;
          }
        }
      } else {
        // ---------------- up pass ----------------
        bool special = false;  // the stage holds a gap row (first row above a region boundary): generic path, CTA-uniform
        if (g.nreg == 3) special = (g.rstart[1] - 1 >= t0 && g.rstart[1] - 1 < t0 + rows) || (g.rstart[2] - 1 >= t0 && g.rstart[2] - 1 < t0 + rows);
        const bool mu0p = (o == 0 && k == 0 && tid == 0);  // column M inside the tile: I_n = J (SOS_Aer_I1_In.py:100)
        const bool fast = !special && vu && !mu0p;
        double dtmax = tn - tcv[R - 1];  // (tau_pad repeats the last value past row L-1)
#pragma unroll
        for (int r = 1; r < R; ++r) dtmax = fmax(dtmax, tcv[r] - tcv[r - 1]);
        const bool tiny = __all_sync(0xffffffffu, !fast || dtmax * fabs(imuu) <= kTinyArg);
        if (fast) {
          double jv[R], xv[R], av[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int t = t0 + r;
            const bool gen = rk > 0 && (t < a0 || t >= a1);
            jv[r] = gen ? fma(cj[r][1], vu1, cj[r][0] * vu0) : Jt[r * W + tid];
          }
          {
            double tnx = tn;
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
              if (r < rows) {
                xv[r] = (tcv[r] - tnx) * imuu;  // = -dt / mu
                tnx = tcv[r];
              } else {
                xv[r] = 0.0;
              }
            }
          }
          if (tiny) {
#pragma unroll
            for (int r = 0; r < R; ++r) av[r] = exp_tiny(xv[r]);
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) av[r] = exp_small(xv[r]);
          }
          {
            double jnx = Jn;
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
              if (r < rows) {
                xv[r] = (0.5 * xv[r]) * (jv[r] + jnx * av[r]);  // = -b
                jnx = jv[r];
              }
            }
          }
#pragma unroll
          for (int r = R - 1; r >= 0; --r) {
            if (r < rows) {
              U = (t0 + r == L - 1) ? U : U * av[r] - xv[r];  // surface row: zero-length integral, the seed itself
              Jt[r * W + tid] = U;
              It[r * W + tid] += U;
            }
          }
          Jn = jv[0];
          tn = tcv[0];
        } else {
          for (int r = rows - 1; r >= 0; --r) {
            const int t = t0 + r;
            const double tc = ta[r];
            const bool gen = rk > 0 && (t < a0 || t >= a1);
            const double jt = gen ? fma(cj[r][1], vu1, cj[r][0] * vu0) : Jt[r * W + tid];
            const double io = It[r * W + tid];
            double val;
            if (special && (t + 1 == g.rstart[1] || t + 1 == g.rstart[2])) {
              // U holds the raw value at the carry row t+1: it is read after its blend ...
              if (k == 0) {
                // s_row[i] = raw I_n[t+1, M+i]
                if (!mu0p) s_row[tid + o] = U;
                if (tid == 0) s_row[0] = (o == 0) ? lastJ : ((r + 1 < rows) ? s_pV[r + 1] : lastJ);
                consumer_sync();
                if (warp == 0) {
                  const int lim = min(W + o, N - M);
                  int istar = -1;
                  for (int base = 1; base + 2 <= lim - 1 && istar < 0; base += 32) {
                    const int i = base + lane;
                    bool hit = false;
                    if (i + 2 <= lim - 1) {
                      const double x = s_row[i], y = s_row[i + 1], z = s_row[i + 2];
                      hit = !(fabs((x - y) - (y - z)) > SOS_BLEND_THRESHOLD);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, hit);
                    if (mask) istar = base + __ffs(mask) - 1 + 1;
                  }
                  if (lane == 0) {
                    s_istar = istar;
                    if (istar < 0) atomicOr(&g.state[s].status, (N - M <= W + o) ? SOS_STATUS_BLEND_OVERRUN : SOS_STATUS_STRIP_FALLBACK);
                  }
                }
                consumer_sync();
                const int istar = s_istar;
                const int ri = tid + o;
                if (istar > 0 && ri > 0 && ri < istar) {
                  const double w = muu / g.mu[M + istar];
                  U = (1.0 - w) * s_row[0] + w * s_row[istar];
                }
                consumer_sync();  // s_row is reused at the next boundary
              }
              // ... then crosses the gap with pure attenuation (SOS_Aer_main_specular.py:413,433)
              U = U * exp(-(tn - tc) / muu);
              val = U;
            } else if (t == L - 1) {
              val = U;
            } else {
              const double dt = tn - tc;
              const double a = exp_small(-dt * imuu);
              U = U * a + (dt * 0.5) * (jt + Jn * a) * imuu;
              val = U;
            }
            if (mu0p) { val = jt; lastJ = jt; }
            if (!vu) val = 0.0;
            Jt[r * W + tid] = val;
            if (vu) It[r * W + tid] = io + val;
            Jn = jt;
            tn = tc;
          }
        }
        if (side && tid == 0) lastJ = s_pV[0];  // the next stage may open with a gap row whose carry row is this stage's first
      }
      if (k != 0) fence_async_smem();  // tiles of the other strips are final here
      consumer_sync();
      if (k != 0 && tid == 0) mbar_arrive(&ready_bar[gq % NS]);  // -> producer: store them, refill the ring

      // ---------------- strip 0: prefetch for the next step, then finish the columns next to mu = 0 in shared memory -------
      double lk_val = 0.0;
      int lk_k0 = 0, lk_r = -1, lk_c = 0;
      double nx_I = 0.0, nx_J = 0.0;
      bool nx_have = false;
      if (k == 0) {
        if (q + 1 < nseq) {
          bool downn; int t0n, rowsn;
          stage_rows(q + 1, downn, t0n, rowsn);
          if (downn && p.nsc > 0 && tid < rowsn * p.nsc) {
            lk_r = tid / p.nsc;
            lk_c = tid - lk_r * p.nsc;
            lk_k0 = p.k0tab[(static_cast<size_t>(s) * L + t0n + lk_r) * p.nsc + lk_c];
            if (lk_k0 < t0n) lk_val = p.dhist[(static_cast<size_t>(s) * L + lk_k0) * MAX_SMALL + lk_c];
          }
          if (side_thread && side_r < rowsn) {
            const int t = t0n + side_r;
            const size_t rowoff = (static_cast<size_t>(s) * L + t) * ld;
            nx_have = true;
            nx_I = p.I[rowoff + (downn ? M - 1 : M)];
            if (!downn && !(rk > 0 && (t < a0 || t >= a1))) nx_J = p.J[rowoff + M];
          }
        }
        if (down) {
          if (wmaxs > 0) {
            for (int e = tid; e < rows * wmaxs; e += THREADS) {
              const int r = e / wmaxs, i = e - r * wmaxs;
              const int reg = region_of(g, t0 + r);
              const int idxw = sc.extrap_width[reg];
              if (i >= idxw) continue;
              const int wc = sossweep::width_class(g, idxw);
              const int ns = g.wns[wc];
              const int src0 = (idxw < 2) ? (M - idxw - 2) : (M - idxw - ns);
              const double* __restrict__ Wm = g.W + g.woff[wc];
              double v = 0.0;
              for (int j = 0; j < ns; ++j) v += Wm[i * ns + j] * Jt[r * W + (src0 + j - dcol0)];
              const int m = M - 1 - i, idx = m - dcol0;
              if (idx < W) {
                const bool stdc = (m < M - 1) && fabs(g.mu[m]) >= SOS_MU_THRESHOLD;
                const double raw = Jt[r * W + idx];
                Jt[r * W + idx] = v;
                It[r * W + idx] += stdc ? (v - raw) : v;  // standard targets were accumulated raw
              } else {
                s_pV[r] = v;  // column M-1 of an odd grid lives outside the tile
              }
            }
          }
        } else {
          const int lim = min(W + o, N - M);
          for (int r = warp; r < rows; r += THREADS / 32) {
            double* row = Jt + r * W - o;  // row[i] = I_n[t, M+i] for i >= o
            const double v0 = o ? s_pV[r] : row[0];
            int istar = -1;
            for (int base = 1; base + 2 <= lim - 1 && istar < 0; base += 32) {
              const int i = base + lane;
              bool hit = false;
              if (i + 2 <= lim - 1) {
                const double x = row[i], y = row[i + 1], z = row[i + 2];
                hit = !(fabs((x - y) - (y - z)) > SOS_BLEND_THRESHOLD);
              }
              const unsigned mask = __ballot_sync(0xffffffffu, hit);
              if (mask) istar = base + __ffs(mask) - 1 + 1;
            }
            if (istar < 0) {
              if (lane == 0) atomicOr(&g.state[s].status, (N - M <= W + o) ? SOS_STATUS_BLEND_OVERRUN : SOS_STATUS_STRIP_FALLBACK);
            } else {
              const double v1 = row[istar];
              const double mus = g.mu[M + istar];
              __syncwarp();
              for (int i = 1 + lane; i < istar; i += 32) {
                const double w = g.mu[M + i] / mus;
                const double val = (1.0 - w) * v0 + w * v1;
                const double old = row[i];
                row[i] = val;
                It[r * W + i - o] += val - old;
              }
            }
          }
        }
        fence_async_smem();
        consumer_sync();
        if (tid == 0) mbar_arrive(&ready_bar[gq % NS]);
      }

      // ---------------- rows are final: surface seeds and convergence ratios (two stages per work item) ----------------
      const bool last_down = down && st == nst - 1;
      const bool toa_up = !down && st == 0;
      if (last_down || toa_up) {
        const int r = last_down ? (L - 1 - t0) : 0;
        const int cix = last_down ? ci : tid;
        const bool valid = last_down ? vd : vu;
        const double val = Jt[r * W + cix];
        double rmax = -INFINITY;
        bool nonfinite = false;
        if (valid) {
          const double ratio = val / It[r * W + cix];
          if (isnan(ratio)) nonfinite = true; else rmax = ratio;
        }
        if (side && tid == 0) {  // the mu = 0 column of this pass takes part in the ratio (SOS_Aer_main_specular.py:309)
          const double ratio = s_pV[r] / (s_pI[r] + s_pV[r]);
          if (isnan(ratio)) nonfinite = true; else rmax = fmax(rmax, ratio);
        }
        double lam = 0.0;
        if (last_down) {
          seed_src = valid ? val : 0.0;
          if (g.surface == SOS_SURFACE_LAMBERT && vd && d <= M - 2) {
            // -2 rho trapz(I mu, mu) over columns M-2 .. 0 as a weighted sum (SOS_Aer_main_lambertian.py:399,401)
            double wgt = 0.0;
            if (d >= 1) wgt += (g.mu[d - 1] - g.mu[d]) * 0.5;
            if (d + 1 <= M - 2) wgt += (g.mu[d] - g.mu[d + 1]) * 0.5;
            lam = wgt * val * mud;
          }
        }
#pragma unroll
        for (int oo = 16; oo > 0; oo >>= 1) {
          rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, oo));
          lam += __shfl_xor_sync(0xffffffffu, lam, oo);
        }
        nonfinite = __any_sync(0xffffffffu, nonfinite);
        if (lane == 0) { s_red[warp] = rmax; s_red[4 + warp] = lam; }
        if (nonfinite && lane == 0) atomicOr(&g.state[s].status, SOS_STATUS_NONFINITE);
        consumer_sync();
        if (tid == 0) {
          const double rm = fmax(fmax(s_red[0], s_red[1]), fmax(s_red[2], s_red[3]));
          if (rm == INFINITY) atomicOr(&g.state[s].status, SOS_STATUS_NONFINITE);
          p.ratio_part[(static_cast<size_t>(s) * p.nstrips + k) * 2 + (last_down ? 1 : 0)] = rm;
        }
        if (last_down) {
          if (g.surface == SOS_SURFACE_SPECULAR) {
            U = sc.grd_alb * seed_src;
          } else if (g.surface == SOS_SURFACE_LAMBERT) {
            if (tid == 0) {
              p.lam_part[static_cast<size_t>(s) * p.nstrips + k] = (s_red[4] + s_red[5]) + (s_red[6] + s_red[7]);
              __threadfence();
              atomicExch(&p.lam_flag[static_cast<size_t>(s) * p.nstrips + k], p.epoch);
              // every strip of the scenario holds a neighbouring ticket, so its CTA is running (or done)
              double totl = 0.0;
              for (int kk = 0; kk < p.nstrips; ++kk) {
                volatile unsigned* fl = p.lam_flag + static_cast<size_t>(s) * p.nstrips + kk;
                while (*fl != p.epoch) __nanosleep(64);
                __threadfence();
                totl += *reinterpret_cast<volatile double*>(p.lam_part + static_cast<size_t>(s) * p.nstrips + kk);
              }
              s_red[0] = -2.0 * sc.grd_alb * totl;
            }
            consumer_sync();
            U = s_red[0];
          } else {
            U = 0.0;
          }
        }
        consumer_sync();
      }

      // ---------------- projections of the finished rows onto the molecular factors: one m8n8k4 DMMA chain per warp ----------
      // C[row][n] += sum_k I_n[row][32 warp + k] Us[k][n]; lanes with lane % 4 == 0 end up with (n = 0, 1) of row lane / 4
      if (rk > 0) {
        const int half = down ? 0 : 1;
        const int gq4 = lane >> 2, t4 = lane & 3;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const int col = 32 * warp + 4 * ks + t4;
          const double a = Jt[gq4 * W + col];
          const double b = (gq4 < 2) ? s_us[half][col][gq4] : 0.0;
          sosgemm::dmma884(c0, c1, a, b);
        }
        if (side && warp == 0 && t4 == 0) {  // the mu = 0 column of this pass
          const double x = s_pV[gq4];
          c0 = fma(x, down ? uM1a : uMa, c0);
          c1 = fma(x, down ? uM1b : uMb, c1);
        }
        if (t4 == 0 && gq4 < rows) {
          double2* dst = reinterpret_cast<double2*>(p.proj_out + ((static_cast<size_t>(s) * L + t0 + gq4) * nslots + (2 * k + half) * PROJ_WARPS + warp) * 2);
          *dst = make_double2(c0, c1);
        }
      }
      if (k == 0) {
        // the mu = 0 column of this pass goes back with plain stores
        if (side_thread && side_r < rows) {
          const size_t off = (static_cast<size_t>(s) * L + t0 + side_r) * ld + (down ? M - 1 : M);
          const double v = s_pV[side_r];
          if (hasdense || p.store_all) p.In[off] = v;
          if (p.has_saved) p.saved[off] = v;
          p.I[off] = s_pI[side_r] + v;
        }
        consumer_sync();  // s_pV / s_pI / s_lk* of this stage have been read by everyone
        if (lk_r >= 0) { s_lkv[lk_r][lk_c] = lk_val; s_lkk[lk_r][lk_c] = lk_k0; }
        if (nx_have) { s_pI[side_r] = nx_I; s_pJ[side_r] = nx_J; }
      }
    }
    seq += nseq;
  }
  if (tid == THREADS) bulk_wait_read<0>();
}

// Projections of a whole field (the first order, before the loop): proj[s][t][0][r] = sum_k I[t, k] Us[k][r], other slots 0.
__global__ void __launch_bounds__(256) strip_project_kernel(const GridDev g, const double* __restrict__ I, const double* const* Ut_tab,
                                                            const int* rank_tab, int ldr, int nslots, double* __restrict__ proj) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y, t = blockIdx.x * 8 + warp;
  if (t >= g.L) return;
  const int op = g.scen[s].phase_atm;
  double* dst = proj + (static_cast<size_t>(s) * g.L + t) * nslots * 2;
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
    p0 += __shfl_xor_sync(0xffffffffu, p0, o);
    p1 += __shfl_xor_sync(0xffffffffu, p1, o);
  }
  for (int j = lane; j < nslots * 2; j += 32) dst[j] = (j == 0) ? p0 : (j == 1 ? p1 : 0.0);
}

}  // namespace sosstrip
