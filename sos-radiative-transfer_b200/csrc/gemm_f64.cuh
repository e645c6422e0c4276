// Source contraction J = coef * (I_{n-1} . A)  -- Jn_NumInt (SOS_Aer_I1_In.py:62-74) and the inlined
// two-operand version for aerosol rows (SOS_Aer_main_specular.py:315-323) as ONE FP64 GEMM over the
// stacked rows of all scenarios that are still iterating.
//
// tcgen05.mma has no f64 kind, so the FP64 tensor path is mma.sync.aligned.m8n8k4.f64 (SASS
// DMMA.8x8x4; measured 37.1 TFLOP/s on B200 against 34.1 for a pure DFMA loop,
// profiles/r01_fp64_peak.json -- and a register-tiled DFMA version of this kernel reached 19.6
// TFLOP/s where the DMMA version reached 24.7 on the same tiles, profiles/r01_gemm_variants.md).
// Everything around the DMMA is Blackwell-native:
//   * operand tiles arrive by TMA (cp.async.bulk.tensor.2d) into a 5-stage full/empty mbarrier ring
//     fed by a dedicated producer warp; OOB zero-fill handles the ragged N = 1002 edge in k and n;
//   * the I operand is fetched in 8-row SEGMENTS (one 1 KB, 128B-swizzled box each).  A row tile is
//     any 8 or 16 segments taken from a device-built list of the ACTIVE scenarios' segments, so
//     tiles are always full (no padding at region or scenario boundaries), aerosol rows of different
//     scenarios share two-operand tiles, and converged scenarios cost nothing;
//   * CTAs are persistent (one per SM); the producer draws tiles from a global atomic counter and
//     hands the tile index to the consumers through the stage ring (heavy two-operand tiles first);
//   * setmaxnreg moves the producer warpgroup's registers to the consumer warpgroups.
//
// Warp tile = MB x NB mma blocks (shipped: 32 x 32 and 32 x 48); fragments:
//   A (8x4, row)  lane (g = lane/4, t = lane%4) holds I[r0+g][k0+t]; in the swizzled segment the
//                 8 rows x 32 B fall on 4 distinct chunk pairs -> 2 wavefronts (the minimum)
//   B (4x8, col)  lane holds A[k0+t][n0+g]; operand rows are padded to BN+8 doubles by loading a
//                 wider TMA box, so the 4 k-rows sit 64 B apart modulo 128 B -> 2 wavefronts
//   C (8x8)       lane holds J[r0+g][n0+2t], J[r0+g][n0+2t+1]
#pragma once
#include "common.cuh"

namespace sosgemm {

constexpr int BK = 16;
constexpr int NSTAGES = 5;
constexpr int SEG_ROWS = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GemmParams {
  CUtensorMap map_I;                 // [rows_total][N] (stride ld), box {BK, SEG_ROWS}, 128B swizzle
  CUtensorMap map_peer[SOS_MAX_PEERS];  // same field on every mu-block owner (peer GPU memory over NVLink)
  int n_peers;                       // 0: single-GPU operand (map_I); G: columns [peer_col[r], peer_col[r+1]) live on peer r
  int peer_col[SOS_MAX_PEERS + 1];
  CUtensorMap map_A[SOS_MAX_PHASE];  // [N][N] (stride lda), box {BN+8, BK}, no swizzle
  const TilePlan* plan;              // device, rebuilt after every convergence update
  int* work_counter;                 // device, zeroed before every launch
  const int* active_list;            // compacted scenario ids per group
  const int* seg_row[2];             // local first row of each segment, per class
  const int* seg_valid[2];           // valid rows (1..8) of each segment, per class
  int nseg[2];
  int n_col_tiles;                   // column tiles this plan computes ...
  int ct0;                           // ... starting at this one (mu-block sharding)
  int split_passes;                  // 1: the two operand passes of class-1 tiles are SEPARATE tiles that add their
                                     //    scaled partial into a zeroed J (2 partials: order independent) -- small batches
  int seg_begin, seg_end;            // restrict every group to these segments of its list (row-chunked launches
                                     // that overlap the all-gather of the other chunks); [0, INT_MAX) = all
  int L, N, ld;
  double* J;
  const sos_scenario* scen;
};

// What the consumers need to know about a tile; written once by the producer warp into shared memory
// so that the consumer warps never touch global memory except for their J stores.
template <int SEGS>
struct TileInfo {
  double coef[SEGS];     // final scale of each segment's rows
  double rescale[SEGS];  // accumulator rescale between the two passes of a class-1 tile
  int row[SEGS];         // stacked first row of the segment
  int valid[SEGS];       // valid rows (0 = empty segment)
  int ct;                // column tile
  int passes;            // 1 or 2
  int first_pass;        // operand of the first (or only) pass: 0 = A, 1 = B of the group
  int atomic_out;        // 1: add into J instead of storing
  int ks0, ks1;          // k-steps of this tile (folded kernel with split k: two tiles share an output tile)
};

template <int WM, int WN, int MB, int NB>
struct Cfg {
  // warp tile = (8*MB rows) x (8*NB columns): MB x NB mma blocks
  static constexpr int BM = 8 * MB * WM;
  static constexpr int BN = 8 * NB * WN;
  static constexpr int BN_PAD = BN + 8;
  static constexpr int SEGS = BM / SEG_ROWS;
  static constexpr int CONSUMER_WARPS = WM * WN;
  // the producer gets a whole warpgroup (one working warp) so that setmaxnreg can move its registers
  // to the consumer warpgroups (register allocation is per warpgroup)
  static constexpr int THREADS = 32 * (CONSUMER_WARPS + 4);
  static constexpr bool REBALANCE = (THREADS > 256);
  // setmaxnreg can only redistribute the registers the CTA was launched with: the kernel is compiled
  // for LAUNCH_REGS per thread (what __launch_bounds__(THREADS, 1) allows), the producer warpgroup
  // drops to REGS_PRODUCER and the consumer warpgroups share what that frees.  Asking for more than
  // the launch pool holds would block forever.
  // (capping the kernel at half of the register file so that another stream's HBM-bound sweeps can be
  //  co-resident was tried: the kernels co-run but contend for L2/HBM, 5 % net gain at best -- not shipped)
  static constexpr int MIN_CTAS = 1;
  static constexpr int LAUNCH_REGS = (65536 / (THREADS * MIN_CTAS)) / 8 * 8 > 255 ? 248 : (65536 / (THREADS * MIN_CTAS)) / 8 * 8;
  static constexpr int REGS_PRODUCER = 40;
  static constexpr int REGS_CONSUMER_RAW = ((THREADS * LAUNCH_REGS - 128 * REGS_PRODUCER) / (32 * CONSUMER_WARPS)) / 8 * 8;
  static constexpr int REGS_CONSUMER = REGS_CONSUMER_RAW > 232 ? 232 : REGS_CONSUMER_RAW;
  static constexpr int A_BYTES = BM * BK * 8;
  static constexpr int B_BYTES = BK * BN_PAD * 8;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int INFO_SLOTS = NSTAGES + 2;  // producer runs at most NSTAGES stages ahead and a stage is released before the epilogue
  static constexpr int SMEM = NSTAGES * STAGE_BYTES + 1024 /*align*/ + 2 * NSTAGES * 8 + NSTAGES * 8 +
                              INFO_SLOTS * static_cast<int>(sizeof(TileInfo<SEGS>));
};

// segment q of tile `lt` of a group -> (scenario, stacked first row, valid rows); valid = 0 if none
struct SegRef {
  int scen, row, valid;
};
__device__ __forceinline__ SegRef seg_lookup(const GemmParams& p, int cls, int nactive, int list_off, int q) {
  SegRef r;
  const int nseg = p.nseg[cls];
  const int rank = q / nseg;
  if (rank >= nactive || q >= p.seg_end) { r.scen = -1; r.row = 0; r.valid = 0; return r; }
  const int j = q - rank * nseg;
  r.scen = p.active_list[list_off + rank];
  r.row = r.scen * p.L + p.seg_row[cls][j];
  r.valid = p.seg_valid[cls][j];
  return r;
}

__device__ __forceinline__ int find_group(const TilePlan* plan, int rt) {
  int g = 0;
  while (g + 1 < plan->n_groups && rt >= plan->group_tile_start[g + 1]) ++g;
  return g;
}

template <int WM, int WN, int MB, int NB>
__global__ void __launch_bounds__(Cfg<WM, WN, MB, NB>::THREADS, Cfg<WM, WN, MB, NB>::MIN_CTAS)
jn_gemm_dmma_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<WM, WN, MB, NB>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: every 8-row segment is one swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGES;
  volatile int* tile_ring = reinterpret_cast<volatile int*>(empty_bar + NSTAGES);  // [NSTAGES] (+pad)
  using Info = TileInfo<C::SEGS>;
  Info* tile_info = reinterpret_cast<Info*>(const_cast<int*>(tile_ring) + 2 * NSTAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const TilePlan* plan = p.plan;
  const int ksteps = (p.N + BK - 1) / BK;
  // row-restricted launch (single group only): tiles cover segments [seg_begin, seg_end)
  const bool restricted = p.seg_begin > 0 || p.seg_end != 0x7fffffff;
  const int n_row_tiles = restricted ? (min(p.seg_end, plan->group_nactive[0] * p.nseg[plan->group_cls[0]]) - p.seg_begin + C::SEGS - 1) / C::SEGS
                                     : plan->n_row_tiles;
  const int n_tiles = max(n_row_tiles, 0) * p.n_col_tiles;
  const uint32_t smem_base = smem_u32(smem);

  if (warp >= C::CONSUMER_WARPS) {
    // ===================== TMA producer warp =====================
    if constexpr (C::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_PRODUCER));
    if (warp != C::CONSUMER_WARPS) return;
    int stage = 0;
    uint32_t phase = 0;
    int seq = 0;
    while (true) {
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.work_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= n_tiles) {
        // sentinel stage: tells the consumers to stop
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          tile_ring[stage] = -1;
          mbar_arrive(&full_bar[stage]);
        }
        break;
      }
      const int rt = tile / p.n_col_tiles;
      const int ct = p.ct0 + tile - rt * p.n_col_tiles;
      const int g = restricted ? 0 : find_group(plan, rt);
      const int cls = plan->group_cls[g];
      int lt = restricted ? rt + p.seg_begin / C::SEGS : rt - plan->group_tile_start[g];
      // split class-1 tiles: local tile index = 2 * row tile + pass
      const bool split = p.split_passes && cls == 1;
      const int only_pass = split ? (lt & 1) : 0;
      if (split) lt >>= 1;
      // lanes 0..SEGS-1 own one I segment each, lane SEGS owns the operand tile
      SegRef sr;
      sr.scen = -1; sr.row = 0; sr.valid = 0;
      if (lane < C::SEGS) sr = seg_lookup(p, cls, plan->group_nactive[g], plan->group_list_off[g], lt * C::SEGS + lane);
      const unsigned have = __ballot_sync(0xffffffffu, sr.valid > 0);
      const uint32_t tx = static_cast<uint32_t>(__popc(have)) * (SEG_ROWS * BK * 8) + C::B_BYTES;
      // tile info for the consumers (slot reuse is safe: see INFO_SLOTS)
      Info* info = &tile_info[seq % C::INFO_SLOTS];
      if (lane < C::SEGS) {
        double coef = 0.0, resc = 0.0;
        if (sr.scen >= 0) {
          const sos_scenario& sc = p.scen[sr.scen];
          if (cls == 1 && split) { coef = only_pass ? sc.coef_mix_aer : sc.coef_mix_atm; resc = 1.0; }
          else if (cls == 1) { coef = sc.coef_mix_aer; resc = sc.coef_mix_atm / sc.coef_mix_aer; }
          else coef = sc.coef_atm;
        }
        info->coef[lane] = coef;
        info->rescale[lane] = resc;
        info->row[lane] = sr.row;
        info->valid[lane] = sr.valid;
      }
      const int passes = (cls == 1 && !split) ? 2 : 1;
      if (lane == 0) { info->ct = ct; info->passes = passes; info->first_pass = only_pass; info->atomic_out = split ? 1 : 0; }
      __syncwarp();
      ++seq;
      for (int pass = 0; pass < passes; ++pass) {
        const CUtensorMap* mapA = &p.map_A[(pass + only_pass) == 0 ? plan->group_phaseA[g] : plan->group_phaseB[g]];
        int owner = 0;
        for (int ks = 0; ks < ksteps; ++ks) {
          // fused all-gather: the k-range of this step is fetched straight from the GPU that owns those mu
          // columns (TMA over NVLink peer memory), so the exchange overlaps the DMMA work tile by tile
          const CUtensorMap* mapI = &p.map_I;
          if (p.n_peers > 0) {
            while (owner + 1 < p.n_peers && ks * BK >= p.peer_col[owner + 1]) ++owner;
            mapI = &p.map_peer[owner];
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t dst = smem_base + stage * C::STAGE_BYTES;
          if (lane == 0) {
            if (pass == 0 && ks == 0) tile_ring[stage] = tile;
            mbar_expect_tx(&full_bar[stage], tx);
          }
          if (sr.valid > 0) tma_load_2d(dst + lane * (SEG_ROWS * BK * 8), mapI, &full_bar[stage], ks * BK, sr.row);
          if (lane == C::SEGS) tma_load_2d(dst + C::A_BYTES, mapA, &full_bar[stage], ct * C::BN, ks * BK);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumers: DMMA =====================
  if constexpr (C::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_CONSUMER));
  const int warp_m = warp / WN;
  const int warp_n = warp - warp_m * WN;
  const int g8 = lane >> 2;  // 0..7
  const int t4 = lane & 3;   // 0..3

  // A fragment of m-block i (= segment warp_m*8 + i), k-slab j: row g8 of the segment, k = 4 j + t4
  //   byte = seg*1024 + g8*128 + (((k >> 1) ^ g8) << 4) + (k & 1)*8 ; (4j + t4) >> 1 = 2j + (t4 >> 1)
  uint32_t a_off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    a_off[j] = static_cast<uint32_t>((warp_m * 8 * MB + g8) * 128 + ((((2 * j + (t4 >> 1)) ^ g8) << 4) | ((t4 & 1) << 3)));
  // B fragment of n-block q, k-slab j: row k = 4 j + t4, col n = warp_n*32 + 8 q + g8
  const uint32_t b_off = static_cast<uint32_t>(C::A_BYTES + (t4 * C::BN_PAD + warp_n * 8 * NB + g8) * 8);

  int stage = 0;
  uint32_t phase = 0;
  int seq = 0;
  while (true) {
    mbar_wait(&full_bar[stage], phase);
    const int tile = tile_ring[stage];
    if (tile < 0) break;
    const Info* info = &tile_info[seq % C::INFO_SLOTS];
    ++seq;
    const int ct = info->ct;
    const int passes = info->passes;
    const int atomic_out = info->atomic_out;

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int q = 0; q < NB; ++q) acc[i][q][0] = acc[i][q][1] = 0.0;

    for (int pass = 0; pass < passes; ++pass) {
      if (pass == 1) {
        // J = c1*(I.A1) + c2*(I.A2) = c2 * ((c1/c2)*(I.A1) + I.A2): rescale once per segment, one accumulator
#pragma unroll
        for (int i = 0; i < MB; ++i) {
          const double r = info->rescale[warp_m * MB + i];
#pragma unroll
          for (int q = 0; q < NB; ++q) { acc[i][q][0] *= r; acc[i][q][1] *= r; }
        }
      }
      for (int ks = 0; ks < ksteps; ++ks) {
        if (pass != 0 || ks != 0) mbar_wait(&full_bar[stage], phase);
        const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sB = sA + b_off;
        if constexpr (MB == 8) {
          // 2 warps per SMSP: prefetch the next k-slab's fragments while this slab's DMMAs issue
          double fa[2][MB], fb[2][NB];
#pragma unroll
          for (int i = 0; i < MB; ++i) fa[0][i] = lds64(sA + a_off[0] + static_cast<uint32_t>(i * 1024));
#pragma unroll
          for (int q = 0; q < NB; ++q) fb[0][q] = lds64(sB + static_cast<uint32_t>(8 * q * 8));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cur = j & 1, nxt = cur ^ 1;
            if (j + 1 < 4) {
#pragma unroll
              for (int i = 0; i < MB; ++i) fa[nxt][i] = lds64(sA + a_off[j + 1] + static_cast<uint32_t>(i * 1024));
#pragma unroll
              for (int q = 0; q < NB; ++q)
                fb[nxt][q] = lds64(sB + static_cast<uint32_t>((4 * (j + 1) * C::BN_PAD + 8 * q) * 8));
            }
#pragma unroll
            for (int i = 0; i < MB; ++i)
#pragma unroll
              for (int q = 0; q < NB; ++q) dmma884(acc[i][q][0], acc[i][q][1], fa[cur][i], fb[cur][q]);
          }
        } else {
          // 4 warps per SMSP hide the fragment-load latency; registers are the scarce resource
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double fa[MB], fb[NB];
#pragma unroll
            for (int i = 0; i < MB; ++i) fa[i] = lds64(sA + a_off[j] + static_cast<uint32_t>(i * 1024));
#pragma unroll
            for (int q = 0; q < NB; ++q) fb[q] = lds64(sB + static_cast<uint32_t>((4 * j * C::BN_PAD + 8 * q) * 8));
#pragma unroll
            for (int i = 0; i < MB; ++i)
#pragma unroll
              for (int q = 0; q < NB; ++q) dmma884(acc[i][q][0], acc[i][q][1], fa[i], fb[q]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
    }

    // ---- epilogue: lane owns J[r][c], J[r][c+1]; a quad writes 64 contiguous bytes ----
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
          if (atomic_out) {
            if (c < p.N) atomicAdd(out + c, coef * acc[i][q][0]);
            if (c + 1 < p.N) atomicAdd(out + c + 1, coef * acc[i][q][1]);
          } else if (c + 1 < p.N) {
            *reinterpret_cast<double2*>(out + c) = make_double2(coef * acc[i][q][0], coef * acc[i][q][1]);
          } else if (c < p.N) {
            out[c] = coef * acc[i][q][0];
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tile planning: compact the active scenarios of every group and lay the row tiles out.
// One CTA; called after every convergence update (and at plan creation / reset).
// ---------------------------------------------------------------------------------------------
struct GroupTable {
  int n_groups;
  int cls[SOS_MAX_GROUPS];
  int phaseA[SOS_MAX_GROUPS];
  int phaseB[SOS_MAX_GROUPS];
  int member_off[SOS_MAX_GROUPS + 1];  // into `members`
};

__device__ __forceinline__ void plan_tiles_block(const GroupTable& gt, const int* __restrict__ members,
                                                 const ScenState* __restrict__ state, int* __restrict__ active_list,
                                                 TilePlan* plan, const int nseg0, const int nseg1, const int segs_per_tile,
                                                 const int split_passes) {
  // warp w handles groups w, w + nwarps, ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __shared__ int nact[SOS_MAX_GROUPS];
  for (int g = warp; g < gt.n_groups; g += nwarps) {
    const int off = gt.member_off[g], cnt = gt.member_off[g + 1] - off;
    int total = 0;
    for (int base = 0; base < cnt; base += 32) {
      const int i = base + lane;
      int s = -1;
      bool on = false;
      if (i < cnt) { s = members[off + i]; on = state[s].active != 0; }
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) active_list[off + total + __popc(m & ((1u << lane) - 1))] = s;
      total += __popc(m);
    }
    if (lane == 0) nact[g] = total;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int g = 0; g < gt.n_groups; ++g) {
      plan->group_tile_start[g] = t;
      plan->group_cls[g] = gt.cls[g];
      plan->group_nactive[g] = nact[g];
      plan->group_list_off[g] = gt.member_off[g];
      plan->group_phaseA[g] = gt.phaseA[g];
      plan->group_phaseB[g] = gt.phaseB[g];
      if (gt.cls[g] == 3) {
        // low-rank operand: these rows are contracted by jn_lowrank_kernel, no dense tiles
      } else if (gt.cls[g] == 2) {
        // premixed aerosol rows (folded kernel): tiles never mix scenarios, each scenario has its own operand
        t += nact[g] * ((nseg1 + segs_per_tile - 1) / segs_per_tile);
      } else {
        const int segs = nact[g] * (gt.cls[g] == 1 ? nseg1 : nseg0);
        t += ((segs + segs_per_tile - 1) / segs_per_tile) * ((split_passes && gt.cls[g] == 1) ? 2 : 1);
      }
    }
    plan->group_tile_start[gt.n_groups] = t;
    plan->n_groups = gt.n_groups;
    plan->n_row_tiles = t;
  }
}

}  // namespace sosgemm
