// Folded source contraction: half the DMMA work when the operand is centrosymmetric.
//
// The contraction operand A[k][m] = w_k/4 * P[m][N-1-k] (SOS_Aer_I1_In.py:73) inherits
// P(mu, mu') = P(-mu, -mu') from every phase function of SOS_Aer_phase_func.py (azimuthal average of a
// function of the scattering angle) and w_k = w_{N-1-k} from the symmetric mu grid
// (SOS_Aer_main_specular.py:59-61), so A[N-1-k][N-1-m] = A[k][m].  Writing a row of I_{n-1} as
// x = [x1 | x2] (downward | upward half, M = N/2 columns each) and R for the reversal of M columns,
//     u = x1 + x2 R,  v = x1 - x2 R,
//     B+ = (A11 + A12 R) / 2,  B- = (A11 - A12 R) / 2        (M x M each, built once per operand)
//     J[:, j]       = u B+ + v B-
//     J[:, N-1-j]   = u B+ - v B-        (j < M)
// i.e. two M x M x rows contractions instead of one N x N x rows: half of the multiply-adds for the
// same sum (reassociated; agreement with the unfolded kernel is ~1e-14 relative, tests/test_gpu_parity.py).
// The fold costs no extra pass over memory: the producer fetches, for every k-step, the I segment and its
// mirror image (columns N-16-k0 .. N-1-k0), the consumers form u and v while loading the A fragments
// (one DADD each per 4 DMMAs) and keep two accumulator sets that are combined in the epilogue.
// Operands that are not centrosymmetric (a user matrix, a non-symmetric mu grid) keep the general kernel.
#pragma once
#include "gemm_f64.cuh"

namespace sosgemm {

struct FoldParams {
  CUtensorMap map_I;                 // [rows_total][N] (stride ld), box {BK, SEG_ROWS}, 128B swizzle
  CUtensorMap map_F[SOS_MAX_PHASE];  // folded operand [Kp][2*Mh] = [B+ | B-] (stride ldf), box {BN+8, BK}
  CUtensorMap map_mix;               // premixed aerosol operands [S][Kp][2*Mh]: c1_s F[atm_s] + c2_s F[aer_s] (class 2 groups)
  const TilePlan* plan;
  int* work_counter;
  const int* active_list;
  const int* seg_row[2];
  const int* seg_valid[2];
  int nseg[2];
  int n_col_tiles;   // ceil(M / BN)
  int split_passes;  // see GemmParams
  int ksplit;        // 0: chosen per launch on the device (J rows zeroed by the caller; see the kernel); 1, or 2: every output tile is computed by two tiles (halves of the k range) that add their partial into a
                     //    zeroed J (exactly two partials per element: order independent) -- launches with fewer tiles than SMs / 2
  int seg_begin, seg_end;  // row-restricted launch (one single-region group: layer-sharded plans): only the segments [seg_begin, seg_end)
  int L, N, M, Mh, ld;
  double* J;
  const sos_scenario* scen;
};

// Tile shapes: WM x WN consumer warps, each MB x NB blocks of 8 x 8: <4, 2, 4, 4> = 64 x 128 (the default), <3, 2, 4, 4> = 48 x 128,
// <7, 1, 8, 2> = 56 x 128.  One tile per SM is ~70 us of DMMA work at M = 512, so a launch with few tiles (one large grid, or
// a layer block of it) is quantised in whole tiles per SM: the host picks the 48-row shape when (waves x rows per tile) is
// smaller, see source_impl; and an aerosol layer of 54 rows (7 segments per scenario) fills a 56-row tile exactly where a
// 64-row tile issues an eighth of its DMMAs on nothing.
template <int MB_, int WM_ = 2, int WN_ = 4, int NB_ = 4>
struct FoldCfgT {
  static constexpr int WM = WM_, WN = WN_, MB = MB_, NB = NB_;
  static constexpr int NST = 4;
  static constexpr int BM = 8 * MB * WM;    // 64 rows
  static constexpr int BN = 8 * NB * WN;    // 128 folded columns j (-> J columns j and N-1-j)
  static constexpr int BN_PAD = BN + 8;
  static constexpr int SEGS = BM / SEG_ROWS;
  static constexpr int CONSUMER_WARPS = WM * WN;
  static constexpr int THREADS = 32 * (CONSUMER_WARPS + 4);
  static constexpr int LAUNCH_REGS = (65536 / THREADS) / 8 * 8;
  static constexpr int REGS_PRODUCER = 40;
  static constexpr int REGS_CONSUMER_RAW = ((THREADS * LAUNCH_REGS - 128 * REGS_PRODUCER) / (32 * CONSUMER_WARPS)) / 8 * 8;
  static constexpr int REGS_CONSUMER = REGS_CONSUMER_RAW > 232 ? 232 : REGS_CONSUMER_RAW;
  static constexpr int A_HALF = BM * BK * 8;          // the I segments, then their mirror images
  static constexpr int A_BYTES = 2 * A_HALF;
  static constexpr int B_HALF = BK * BN_PAD * 8;      // B+ rows, then B- rows
  static constexpr int B_BYTES = 2 * B_HALF;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int INFO_SLOTS = NST + 2;
  static constexpr int SMEM = NST * STAGE_BYTES + 1024 + 3 * NST * 8 + NST * 8 + INFO_SLOTS * static_cast<int>(sizeof(TileInfo<SEGS>));
};
using FoldCfg = FoldCfgT<4>;

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ SegRef seg_lookup_fold(const FoldParams& p, int cls, int nactive, int list_off, int q) {
  SegRef r;
  const int nseg = p.nseg[cls];
  const int rank = q / nseg;
  if (rank >= nactive) { r.scen = -1; r.row = 0; r.valid = 0; return r; }
  const int j = q - rank * nseg;
  r.scen = p.active_list[list_off + rank];
  r.row = r.scen * p.L + p.seg_row[cls][j];
  r.valid = p.seg_valid[cls][j];
  return r;
}

// XFORM: the three otherwise idle warps of the producer warpgroup turn every landed stage from (x, mirror x) into
// (u, v) in place, once per CTA, so that the consumer warps load u and v directly (otherwise each of the WN consumer
// warps of a row block forms them again while loading its fragments: one DADD per 4 DMMAs on the same FP64 pipe).
template <bool XFORM, int MBT = 4, int WMT = 2, int WNT = 4, int NBT = 4>
__global__ void __launch_bounds__(FoldCfgT<MBT, WMT, WNT, NBT>::THREADS, 1) jn_gemm_fold_kernel(const __grid_constant__ FoldParams p) {
  using C = FoldCfgT<MBT, WMT, WNT, NBT>;
  static_assert(C::BN == 128 && C::CONSUMER_WARPS == 8, "operand boxes and the warp roles assume 128 folded columns and eight consumer warps");
  constexpr int NST = C::NST, MB = C::MB, NB = C::NB, WN = C::WN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NST * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* ready_bar = empty_bar + NST;  // XFORM: u, v of the stage are in place
  volatile int* tile_ring = reinterpret_cast<volatile int*>(ready_bar + NST);
  using Info = TileInfo<C::SEGS>;
  Info* tile_info = reinterpret_cast<Info*>(const_cast<int*>(tile_ring) + 2 * NST);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::CONSUMER_WARPS);
      mbar_init(&ready_bar[s], 3);
      tile_ring[s] = 0;  // the transformer warps look at every stage's slot (only a sentinel is negative)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const TilePlan* plan = p.plan;
  const int ksteps = (p.M + BK - 1) / BK;
  const int last_slabs = (p.M - (ksteps - 1) * BK + 3) / 4;  // k-slabs of the last k-step that hold k < M
  // row-restricted launch: tiles cover the segments [seg_begin, seg_end) of the (single) group's list
  const bool restricted = p.seg_begin > 0 || p.seg_end != 0x7fffffff;
  const int seg_stop = restricted ? min(p.seg_end, plan->group_nactive[0] * p.nseg[plan->group_cls[0]]) : 0;
  const int n_row_tiles = restricted ? max(0, (seg_stop - p.seg_begin + C::SEGS - 1) / C::SEGS) : plan->n_row_tiles;
  // ksplit == 0: a launch is whole tiles per CTA, so halving the tiles pays whenever it saves half a wave: with T output tiles on
  // G CTAs, ceil(2T / G) half-tiles per CTA against 2 ceil(T / G) (T = 156 on 148: 3 instead of 4 half-tile times)
  // With at least one output tile per CTA the launch is not cut into tiles at all ("stream-k"): the T x ksteps k-steps of
  // the launch are dealt to the CTAs as equal contiguous ranges, a range covers the tail of one tile, some whole tiles and
  // the head of another.  A range is at least one tile long, so no tile is shared by more than two CTAs: its two partials are
  // added into the zeroed J (order independent), whole tiles are stored.  T = 384 tiles on 148 CTAs: 2.6 tile times instead of 3.
  int ksplit = p.ksplit;
  bool streamk = false;
  if (ksplit == 0) {
    const int T = n_row_tiles * p.n_col_tiles, G = static_cast<int>(gridDim.x);
    streamk = T >= G && !p.split_passes;
    ksplit = (!streamk && T > 0 && p.M >= 64 && (2 * T + G - 1) / G < 2 * ((T + G - 1) / G)) ? 2 : 1;
  }
  const int n_tiles = n_row_tiles * p.n_col_tiles * ksplit;
  const uint32_t smem_base = smem_u32(smem);

  if (warp >= C::CONSUMER_WARPS) {
    // ===================== TMA producer warp =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_PRODUCER));
    if (warp != C::CONSUMER_WARPS) {
      if constexpr (XFORM) {
        // ===================== transformer warps: (x, mirror x) -> (u, v) in place =====================
        // a stage holds SEGS segments x 8 rows x 8 sixteen-byte chunks; chunk c of the direct box (k = 2c, 2c+1)
        // pairs with chunk 7-c of the mirror box (offsets 15-k: 14-2c, 15-2c)
        const int t96 = (warp - C::CONSUMER_WARPS - 1) * 32 + lane;
        int stage = 0;
        uint32_t phase = 0;
        while (true) {
          mbar_wait(&full_bar[stage], phase);
          if (tile_ring[stage] < 0) {  // sentinel: pass it on
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready_bar[stage]);
            break;
          }
          const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
#pragma unroll 2
          for (int idx = t96; idx < C::SEGS * 64; idx += 96) {
            const int row = (idx >> 3) & 7, c = idx & 7;
            const uint32_t base = static_cast<uint32_t>((idx >> 6) * 1024 + row * 128);
            const uint32_t o1 = sA + base + static_cast<uint32_t>((c ^ row) << 4);
            const uint32_t o2 = sA + static_cast<uint32_t>(C::A_HALF) + base + static_cast<uint32_t>(((7 - c) ^ row) << 4);
            double x1a, x1b, x2a, x2b;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x1a), "=d"(x1b) : "r"(o1));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x2a), "=d"(x2b) : "r"(o2));
            // direct k = 2c pairs with mirror offset 15-2c (second of the chunk), k = 2c+1 with 14-2c (first)
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(o1), "d"(x1a + x2b), "d"(x1b + x2a) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(o2), "d"(x1b - x2a), "d"(x1a - x2b) : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the next TMA fill overwrites these generic writes
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready_bar[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
      return;
    }
    int stage = 0;
    uint32_t phase = 0;
    int seq = 0;
    // stream-k: this CTA's range of the launch's k-steps
    const long long all_steps = static_cast<long long>(n_tiles) * ksteps;
    long long pos = streamk ? all_steps * blockIdx.x / gridDim.x : 0;
    const long long pos_end = streamk ? all_steps * (blockIdx.x + 1) / gridDim.x : 0;
    while (true) {
      int tile = 0;
      if (streamk) tile = pos < pos_end ? static_cast<int>(pos / ksteps) : n_tiles;
      else {
        if (lane == 0) tile = atomicAdd(p.work_counter, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
      }
      if (tile >= n_tiles) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          tile_ring[stage] = -1;
          mbar_arrive(&full_bar[stage]);
        }
        break;
      }
      const int t2 = tile / ksplit;
      const int kh = tile - t2 * ksplit;
      int ks0 = kh * ksteps / ksplit, ks1 = (kh + 1) * ksteps / ksplit;
      if (streamk) {
        ks0 = static_cast<int>(pos - static_cast<long long>(tile) * ksteps);
        ks1 = static_cast<int>(min(static_cast<long long>(ksteps), pos_end - static_cast<long long>(tile) * ksteps));
        pos += ks1 - ks0;
      }
      // tile order: row tile major (the column tiles of a row tile are handed out together and run side by side on neighbouring
      // CTAs, sharing their I rows through L2).  Stream-k ranges are contiguous in tile order, so there the order is column tile
      // major: the column tiles of a row tile then sit n_row_tiles apart = in different CTAs' ranges at about the same offset, and
      // again run at the same time instead of one after the other (1.49 x the algorithmic DRAM bytes otherwise: the rows are
      // re-read after L2 has moved on)
      int rt = t2 / p.n_col_tiles;
      int ct = t2 - rt * p.n_col_tiles;
      if (streamk) { ct = t2 / n_row_tiles; rt = t2 - ct * n_row_tiles; }
      const int g = restricted ? 0 : find_group(plan, rt);
      const int cls = plan->group_cls[g];
      int lt = restricted ? rt : rt - plan->group_tile_start[g];
      const bool split = p.split_passes && cls == 1;
      const int only_pass = split ? (lt & 1) : 0;
      if (split) lt >>= 1;
      SegRef sr;
      sr.scen = -1; sr.row = 0; sr.valid = 0;
      int mix_scen = -1;  // class 2: the tile belongs to one scenario and uses that scenario's premixed operand
      if (cls == 2) {
        const int tps = (p.nseg[1] + C::SEGS - 1) / C::SEGS;
        const int rank = lt / tps;
        mix_scen = p.active_list[plan->group_list_off[g] + rank];
        const int j = (lt - rank * tps) * C::SEGS + lane;
        if (lane < C::SEGS && j < p.nseg[1]) {
          sr.scen = mix_scen;
          sr.row = mix_scen * p.L + p.seg_row[1][j];
          sr.valid = p.seg_valid[1][j];
        }
      } else if (lane < C::SEGS) {
        const int q = (restricted ? p.seg_begin : 0) + lt * C::SEGS + lane;
        if (!restricted || q < seg_stop) sr = seg_lookup_fold(p, cls, plan->group_nactive[g], plan->group_list_off[g], q);
      }
      const unsigned have = __ballot_sync(0xffffffffu, sr.valid > 0);
      const uint32_t tx = static_cast<uint32_t>(__popc(have)) * (2 * SEG_ROWS * BK * 8) + C::B_BYTES;
      Info* info = &tile_info[seq % C::INFO_SLOTS];
      if (lane < C::SEGS) {
        double coef = 0.0, resc = 0.0;
        if (sr.scen >= 0) {
          const sos_scenario& sc = p.scen[sr.scen];
          if (cls == 2) coef = 1.0;  // the mixing coefficients are inside the operand
          else if (cls == 1 && split) { coef = only_pass ? sc.coef_mix_aer : sc.coef_mix_atm; resc = 1.0; }
          else if (cls == 1) { coef = sc.coef_mix_aer; resc = sc.coef_mix_atm / sc.coef_mix_aer; }
          else coef = sc.coef_atm;
        }
        info->coef[lane] = coef;
        info->rescale[lane] = resc;
        info->row[lane] = sr.row;
        info->valid[lane] = sr.valid;
      }
      const int passes = (cls == 1 && !split) ? 2 : 1;
      if (lane == 0) {
        info->ct = ct; info->passes = passes; info->first_pass = only_pass; info->atomic_out = (split || ksplit > 1 || ks1 - ks0 < ksteps) ? 1 : 0;
        info->ks0 = ks0; info->ks1 = ks1;
      }
      __syncwarp();
      ++seq;
      for (int pass = 0; pass < passes; ++pass) {
        const CUtensorMap* mapF = &p.map_F[(pass + only_pass) == 0 ? plan->group_phaseA[g] : plan->group_phaseB[g]];
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t dst = smem_base + stage * C::STAGE_BYTES;
          if (lane == 0) {
            if (pass == 0 && ks == ks0) tile_ring[stage] = tile;
            mbar_expect_tx(&full_bar[stage], tx);
          }
          if (sr.valid > 0) {
            // columns k0 .. k0+15 and their mirror images N-16-k0 .. N-1-k0 (column N-1-k sits at offset 15-(k-k0))
            tma_load_2d(dst + lane * (SEG_ROWS * BK * 8), &p.map_I, &full_bar[stage], ks * BK, sr.row);
            tma_load_2d(dst + C::A_HALF + lane * (SEG_ROWS * BK * 8), &p.map_I, &full_bar[stage], p.N - BK - ks * BK, sr.row);
          }
          if (mix_scen >= 0) {
            if (lane == C::SEGS) tma_load_3d(dst + C::A_BYTES, &p.map_mix, &full_bar[stage], ct * C::BN, ks * BK, mix_scen);
            if (lane == C::SEGS + 1) tma_load_3d(dst + C::A_BYTES + C::B_HALF, &p.map_mix, &full_bar[stage], p.Mh + ct * C::BN, ks * BK, mix_scen);
          } else {
            if (lane == C::SEGS) tma_load_2d(dst + C::A_BYTES, mapF, &full_bar[stage], ct * C::BN, ks * BK);
            if (lane == C::SEGS + 1) tma_load_2d(dst + C::A_BYTES + C::B_HALF, mapF, &full_bar[stage], p.Mh + ct * C::BN, ks * BK);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumers: DMMA on (u, B+) and (v, B-) =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_CONSUMER));
  const int warp_m = warp / WN;
  const int warp_n = warp - warp_m * WN;
  const int g8 = lane >> 2;
  const int t4 = lane & 3;

  // fragment of m-block i, k-slab j: row g8 of segment warp_m*MB + i; k = 4 j + t4 in the direct box,
  // 15 - k in the mirror box (same 128B swizzle: 16-byte chunk index ^ row)
  uint32_t a1_off[4], a2_off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = 4 * j + t4, km = 15 - k;
    const uint32_t rowb = static_cast<uint32_t>((warp_m * 8 * MB + g8) * 128);
    a1_off[j] = rowb + static_cast<uint32_t>((((k >> 1) ^ g8) << 4) | ((k & 1) << 3));
    a2_off[j] = static_cast<uint32_t>(C::A_HALF) + rowb + static_cast<uint32_t>((((km >> 1) ^ g8) << 4) | ((km & 1) << 3));
  }
  const uint32_t b_off = static_cast<uint32_t>(C::A_BYTES + (t4 * C::BN_PAD + warp_n * 8 * NB + g8) * 8);

  int stage = 0;
  uint32_t phase = 0;
  int seq = 0;
  uint64_t* const landed_bar = XFORM ? ready_bar : full_bar;
  while (true) {
    if constexpr (XFORM) mbar_wait(&full_bar[stage], phase);  // (already complete: acquires the TMA writes of the operand)
    mbar_wait(&landed_bar[stage], phase);
    const int tile = tile_ring[stage];
    if (tile < 0) break;
    const Info* info = &tile_info[seq % C::INFO_SLOTS];
    ++seq;
    const int ct = info->ct;
    const int passes = info->passes;
    const int atomic_out = info->atomic_out;
    const int ks0 = info->ks0, ks1 = info->ks1;

    double accP[MB][NB][2], accM[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int q = 0; q < NB; ++q) accP[i][q][0] = accP[i][q][1] = accM[i][q][0] = accM[i][q][1] = 0.0;

    for (int pass = 0; pass < passes; ++pass) {
      if (pass == 1) {
#pragma unroll
        for (int i = 0; i < MB; ++i) {
          const double r = info->rescale[warp_m * MB + i];
#pragma unroll
          for (int q = 0; q < NB; ++q) { accP[i][q][0] *= r; accP[i][q][1] *= r; accM[i][q][0] *= r; accM[i][q][1] *= r; }
        }
      }
      for (int ks = ks0; ks < ks1; ++ks) {
        if (pass != 0 || ks != ks0) {
          if constexpr (XFORM) mbar_wait(&full_bar[stage], phase);
          mbar_wait(&landed_bar[stage], phase);
        }
        const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sB = sA + b_off;
        // one k-slab (4 k values): form u, v while loading the A fragments, then 2 x MB x NB DMMAs
        auto slab = [&](const uint32_t o1, const uint32_t o2, const uint32_t ob) {
          double fu[MB], fv[MB], fp[NB], fm[NB];
#pragma unroll
          for (int i = 0; i < MB; ++i) {
            const double x1 = lds64(sA + o1 + static_cast<uint32_t>(i * 1024));
            const double x2 = lds64(sA + o2 + static_cast<uint32_t>(i * 1024));
            if constexpr (XFORM) { fu[i] = x1; fv[i] = x2; }
            else { fu[i] = x1 + x2; fv[i] = x1 - x2; }
          }
#pragma unroll
          for (int q = 0; q < NB; ++q) {
            fp[q] = lds64(sB + ob + static_cast<uint32_t>(8 * q * 8));
            fm[q] = lds64(sB + ob + static_cast<uint32_t>(C::B_HALF + 8 * q * 8));
          }
#pragma unroll
          for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int q = 0; q < NB; ++q) {
              dmma884(accP[i][q][0], accP[i][q][1], fu[i], fp[q]);
              dmma884(accM[i][q][0], accM[i][q][1], fv[i], fm[q]);
            }
        };
        if (ks + 1 < ksteps || last_slabs == 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) slab(a1_off[j], a2_off[j], static_cast<uint32_t>(4 * j * C::BN_PAD * 8));
        } else {
          // ragged last k-step: the k-slabs at or beyond M multiply zero operand rows -- skip them
#pragma unroll 1
          for (int j = 0; j < last_slabs; ++j) {
            const int k = 4 * j + t4, km = 15 - k;
            const uint32_t rowb = static_cast<uint32_t>((warp_m * 8 * MB + g8) * 128);
            slab(rowb + static_cast<uint32_t>((((k >> 1) ^ g8) << 4) | ((k & 1) << 3)),
                 static_cast<uint32_t>(C::A_HALF) + rowb + static_cast<uint32_t>((((km >> 1) ^ g8) << 4) | ((km & 1) << 3)),
                 static_cast<uint32_t>(4 * j * C::BN_PAD * 8));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }

    // ---- epilogue: lane owns folded columns c, c+1 -> J[r][c], J[r][c+1] and J[r][N-1-c], J[r][N-2-c] ----
    const int col_base = ct * C::BN + warp_n * 8 * NB + 2 * t4;
#pragma unroll
    for (int i = 0; i < MB; ++i) {
      const int sg = warp_m * MB + i;
      if (g8 < info->valid[sg]) {
        const double coef = info->coef[sg];
        double* out = p.J + static_cast<size_t>(info->row[sg] + g8) * p.ld;
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const int c = col_base + 8 * q;
          const double d0 = coef * (accP[i][q][0] + accM[i][q][0]);
          const double d1 = coef * (accP[i][q][1] + accM[i][q][1]);
          const double m0 = coef * (accP[i][q][0] - accM[i][q][0]);
          const double m1 = coef * (accP[i][q][1] - accM[i][q][1]);
          if (atomic_out) {
            if (c < p.M) { atomicAdd(out + c, d0); atomicAdd(out + (p.N - 1 - c), m0); }
            if (c + 1 < p.M) { atomicAdd(out + c + 1, d1); atomicAdd(out + (p.N - 2 - c), m1); }
          } else if (c + 1 < p.M) {
            *reinterpret_cast<double2*>(out + c) = make_double2(d0, d1);
            *reinterpret_cast<double2*>(out + (p.N - 2 - c)) = make_double2(m1, m0);
          } else if (c < p.M) {
            out[c] = d0;
            out[p.N - 1 - c] = m0;
          }
        }
      }
    }
  }
}

// Folded operand F = [B+ | B-] of one contraction operand A, plus the centrosymmetry defect of A.
//   As[k][j] = (A[k][j] + A[N-1-k][N-1-j]) / 2           (symmetrised: both halves of A contribute)
//   B+-[k][j] = (As[k][j] +- As[k][N-1-j]) / 2 ,  k, j < M ; zero elsewhere (rows up to Kp, columns up to Mh)
// stats[0] = max |A[k][j] - A[N-1-k][N-1-j]|, stats[1] = max |A| (bit patterns of non-negative doubles order
// like unsigned integers, so atomicMax on the bits is exact).
__global__ void build_folded_kernel(const double* __restrict__ A, int lda, int N, double* __restrict__ F, int ldf, int Kp, int Mh,
                                    unsigned long long* __restrict__ stats) {
  const int M = N / 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. Mh-1
  const int k = blockIdx.y;                             // 0 .. Kp-1
  double defect = 0.0, amax = 0.0;
  if (j < Mh) {
    double bp = 0.0, bm = 0.0;
    if (k < M && j < M) {
      const double a = A[static_cast<size_t>(k) * lda + j], ar = A[static_cast<size_t>(N - 1 - k) * lda + (N - 1 - j)];
      const double b = A[static_cast<size_t>(k) * lda + (N - 1 - j)], br = A[static_cast<size_t>(N - 1 - k) * lda + j];
      const double s1 = 0.5 * (a + ar), s2 = 0.5 * (b + br);
      bp = 0.5 * (s1 + s2);
      bm = 0.5 * (s1 - s2);
      defect = fmax(fabs(a - ar), fabs(b - br));
      amax = fmax(fmax(fabs(a), fabs(ar)), fmax(fabs(b), fabs(br)));
    }
    F[static_cast<size_t>(k) * ldf + j] = bp;
    F[static_cast<size_t>(k) * ldf + Mh + j] = bm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, o));
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (defect > 0.0) atomicMax(&stats[0], static_cast<unsigned long long>(__double_as_longlong(defect)));
    if (amax > 0.0) atomicMax(&stats[1], static_cast<unsigned long long>(__double_as_longlong(amax)));
  }
}

// Premixed aerosol operands: out[s] = c1_s * F[atm_s] + c2_s * F[aer_s] (c = the two source coefficients of the aerosol rows,
// SOS_Aer_main_specular.py:321), so that an aerosol-row tile needs ONE operand pass instead of two.
struct MixSources {
  const double* F[SOS_MAX_PHASE];
};
__global__ void mix_folded_kernel(MixSources src, const sos_scenario* __restrict__ scen, double* __restrict__ out, size_t elems) {
  const int s = blockIdx.y;
  const sos_scenario sc = scen[s];
  const double* __restrict__ Fa = src.F[sc.phase_atm];
  const double* __restrict__ Fe = src.F[sc.phase_aer];
  double* __restrict__ o = out + static_cast<size_t>(s) * elems;
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < elems; e += static_cast<size_t>(gridDim.x) * blockDim.x)
    o[e] = sc.coef_mix_atm * Fa[e] + sc.coef_mix_aer * Fe[e];
}

}  // namespace sosgemm
