// Source contraction J = coef * (I_{n-1} . A)  -- Jn_NumInt (SOS_Aer_I1_In.py:62-74) and the
// inlined two-operand version for aerosol rows (SOS_Aer_main_specular.py:315-323) as ONE FP64
// GEMM over all stacked scenario rows.
//
// tcgen05.mma has no f64 kind, so the FP64 pipe is driven with register-tiled DFMA micro-kernels;
// everything around them is Blackwell-native: operand tiles arrive by TMA
// (cp.async.bulk.tensor.2d, 128B-swizzled I tile, OOB zero-fill handles the ragged N=1002 edge),
// a dedicated producer warp runs a 4-stage full/empty mbarrier ring, CTAs are persistent
// (one per SM) and walk a host-built tile list that never straddles a region or a scenario.
//
// Tile: BM = 32*WM rows x BN = 64*WN columns, BK = 16.  Consumer warp = 4 (rows) x 8 (cols)
// threads, thread tile 8 x 8:
//   rows  warp_m*32 + i*4 + ty      (i = 0..7)  -> the 4 ty's of a warp hit 4 different 16-byte
//                                                 chunks of the swizzled I tile (no bank conflict)
//   cols  warp_n*64 + j*16 + 2*tx   (j = 0..3)  -> 8 tx's read/write 128 contiguous bytes
#pragma once
#include "common.cuh"

namespace sosgemm {

constexpr int BK = 16;
constexpr int STAGES = 4;

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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

struct GemmParams {
  CUtensorMap map_I;                  // [rows_total][N] (stride ld), box {BK, BM}, 128B swizzle
  CUtensorMap map_A[SOS_MAX_PHASE];   // [N][N] (stride lda), box {BN, BK}, no swizzle
  const GemmTile* tiles;              // row tiles
  int n_row_tiles;
  int n_col_tiles;
  int N;
  int ld;                             // leading dimension of J
  double* J;
  const sos_scenario* scen;
  const ScenState* state;
};

template <int WM, int WN>
struct Cfg {
  static constexpr int BM = 32 * WM;
  static constexpr int BN = 64 * WN;
  static constexpr int CONSUMER_WARPS = WM * WN;
  // the producer gets a whole warpgroup (4 warps, one working lane) so that setmaxnreg can move its
  // registers to the consumer warpgroups (register allocation is per warpgroup)
  static constexpr int THREADS = 32 * (CONSUMER_WARPS + 4);
  static constexpr bool REBALANCE = (THREADS > 256);
  static constexpr int REGS_PRODUCER = 40;
  static constexpr int REGS_CONSUMER = 232;
  static constexpr int A_BYTES = BM * BK * 8;   // I tile
  static constexpr int B_BYTES = BK * BN * 8;   // A tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 2 * STAGES * 8;
};

template <int WM, int WN>
__global__ void __launch_bounds__(Cfg<WM, WN>::THREADS, 1)
jn_gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<WM, WN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled TMA destination
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ksteps = (p.N + BK - 1) / BK;
  const int n_tiles = p.n_row_tiles * p.n_col_tiles;

  if (warp >= C::CONSUMER_WARPS) {
    // ===================== TMA producer (one elected lane) =====================
    if constexpr (C::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_PRODUCER));
    if (warp == C::CONSUMER_WARPS && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rt = tile / p.n_col_tiles;
        const int ct = tile - rt * p.n_col_tiles;
        const GemmTile t = p.tiles[rt];
        if (!p.state[t.scen].active) continue;
        const sos_scenario sc = p.scen[t.scen];
        const int passes = (t.mix && sc.coef_mix_aer != 0.0) ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
          const CUtensorMap* mapA = &p.map_A[pass == 0 ? sc.phase_atm : sc.phase_aer];
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* dst = smem + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(dst, &p.map_I, &full_bar[stage], ks * BK, t.row0);
            tma_load_2d(dst + C::A_BYTES, mapA, &full_bar[stage], ct * C::BN, ks * BK);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }

  // ===================== consumers: DFMA micro-kernels =====================
  if constexpr (C::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_CONSUMER));
  const int warp_m = warp / WN;
  const int warp_n = warp - warp_m * WN;
  const int ty = lane >> 3;  // 0..3
  const int tx = lane & 7;   // 0..7

  // byte offsets inside the I tile: row r at r*128, 16-byte chunk c of row r at (c ^ (r & 7)) * 16
  uint32_t a_pre[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp_m * 32 + i * 4 + ty;
    a_pre[i] = static_cast<uint32_t>(r * 128 + ((r & 7) << 4));
  }
  const uint32_t b_off = static_cast<uint32_t>((warp_n * 64 + 2 * tx) * 8);
  const uint32_t smem_base = smem_u32(smem);

  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int rt = tile / p.n_col_tiles;
    const int ct = tile - rt * p.n_col_tiles;
    const GemmTile t = p.tiles[rt];
    if (!p.state[t.scen].active) continue;
    const sos_scenario sc = p.scen[t.scen];
    const int passes = (t.mix && sc.coef_mix_aer != 0.0) ? 2 : 1;
    const double coef_first = t.mix ? sc.coef_mix_atm : sc.coef_atm;
    const double coef_last = passes == 2 ? sc.coef_mix_aer : coef_first;

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    for (int pass = 0; pass < passes; ++pass) {
      if (pass == 1) {
        // J = c1*(I.A1) + c2*(I.A2) = c2 * ((c1/c2)*(I.A1) + I.A2): rescale once, keep one accumulator
        const double r = coef_first / coef_last;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] *= r;
      }
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sB = sA + C::A_BYTES + b_off;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 2) {
          double2 ra[8];
          double2 rb0[4], rb1[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) ra[i] = lds128(sA + (a_pre[i] ^ static_cast<uint32_t>((kk >> 1) << 4)));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rb0[j] = lds128(sB + static_cast<uint32_t>((kk * C::BN + 16 * j) * 8));
            rb1[j] = lds128(sB + static_cast<uint32_t>(((kk + 1) * C::BN + 16 * j) * 8));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][2 * j] = fma(ra[i].x, rb0[j].x, acc[i][2 * j]);
              acc[i][2 * j + 1] = fma(ra[i].x, rb0[j].y, acc[i][2 * j + 1]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][2 * j] = fma(ra[i].y, rb1[j].x, acc[i][2 * j]);
              acc[i][2 * j + 1] = fma(ra[i].y, rb1[j].y, acc[i][2 * j + 1]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }

    // ---- epilogue: scale and store (8 tx's -> 128 contiguous bytes per row) ----
    const int col_base = ct * C::BN + warp_n * 64 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp_m * 32 + i * 4 + ty;
      if (r < t.nrows) {
        double* out = p.J + static_cast<size_t>(t.row0 + r) * p.ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = col_base + 16 * j;
          if (c + 1 < p.N) {
            *reinterpret_cast<double2*>(out + c) = make_double2(coef_last * acc[i][2 * j], coef_last * acc[i][2 * j + 1]);
          } else if (c < p.N) {
            out[c] = coef_last * acc[i][2 * j];
          }
        }
      }
    }
  }
}


// =============================================================================================
// DMMA variant: the same TMA/mbarrier pipeline, but the inner product runs on the FP64 tensor
// path (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4; measured 37.1 TFLOP/s on B200 against
// 34.1 for a pure DFMA loop, profiles/r01_fp64_peak.json).  One DMMA keeps an SMSP's FP64 pipe
// busy for 16 cycles, so issue slots and shared-memory bandwidth stop being the limiter.
//
// Warp tile 64 x 32 (8 x 4 mma blocks, 64 accumulator doubles per thread); fragments:
//   A (8x4, row)  lane (g = lane/4, t = lane%4) holds I[r0+g][k0+t]   -- 128B-swizzled I tile:
//                 8 rows x 32 B fall on 4 distinct chunk pairs -> 2 wavefronts (the minimum)
//   B (4x8, col)  lane holds A[k0+t][n0+g]                            -- operand rows are padded to
//                 BN+8 doubles by loading a wider TMA box, so the 4 k-rows of a fragment sit
//                 64 B apart modulo 128 B -> 2 wavefronts (the minimum)
//   C (8x8)       lane holds J[r0+g][n0+2t], J[r0+g][n0+2t+1]
template <int WM, int WN>
struct CfgT {
  static constexpr int BM = 64 * WM;
  static constexpr int BN = 32 * WN;
  static constexpr int BN_PAD = BN + 8;
  static constexpr int CONSUMER_WARPS = WM * WN;
  static constexpr int THREADS = 32 * (CONSUMER_WARPS + 4);
  static constexpr bool REBALANCE = (THREADS > 256);
  static constexpr int REGS_PRODUCER = 40;
  static constexpr int REGS_CONSUMER = 232;
  static constexpr int A_BYTES = BM * BK * 8;
  static constexpr int B_BYTES = BK * BN_PAD * 8;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int NSTAGES = 5;
  static constexpr int SMEM = NSTAGES * STAGE_BYTES + 1024 + 2 * NSTAGES * 8;
};

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

template <int WM, int WN>
__global__ void __launch_bounds__(CfgT<WM, WN>::THREADS, 1)
jn_gemm_dmma_kernel(const __grid_constant__ GemmParams p) {
  using C = CfgT<WM, WN>;
  constexpr int NS = C::NSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NS * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NS;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ksteps = (p.N + BK - 1) / BK;
  const int n_tiles = p.n_row_tiles * p.n_col_tiles;

  if (warp >= C::CONSUMER_WARPS) {
    if constexpr (C::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_PRODUCER));
    if (warp == C::CONSUMER_WARPS && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rt = tile / p.n_col_tiles;
        const int ct = tile - rt * p.n_col_tiles;
        const GemmTile t = p.tiles[rt];
        if (!p.state[t.scen].active) continue;
        const sos_scenario sc = p.scen[t.scen];
        const int passes = (t.mix && sc.coef_mix_aer != 0.0) ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
          const CUtensorMap* mapA = &p.map_A[pass == 0 ? sc.phase_atm : sc.phase_aer];
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* dst = smem + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::A_BYTES + C::B_BYTES);
            tma_load_2d(dst, &p.map_I, &full_bar[stage], ks * BK, t.row0);
            tma_load_2d(dst + C::A_BYTES, mapA, &full_bar[stage], ct * C::BN, ks * BK);
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }

  if constexpr (C::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_CONSUMER));
  const int warp_m = warp / WN;
  const int warp_n = warp - warp_m * WN;
  const int g = lane >> 2;  // 0..7
  const int t4 = lane & 3;  // 0..3

  // A fragment of m-block i, k-slab j: row r = warp_m*64 + 8 i + g (r & 7 == g), k = 4 j + t4
  //   byte = r*128 + (((k >> 1) ^ g) << 4) + (k & 1)*8 ; ((4j + t4) >> 1) = 2j + (t4 >> 1)
  uint32_t a_off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    a_off[j] = static_cast<uint32_t>((warp_m * 64 + g) * 128 + ((((2 * j + (t4 >> 1)) ^ g) << 4) | ((t4 & 1) << 3)));
  // B fragment of n-block q, k-slab j: row k = 4 j + t4, col n = warp_n*32 + 8 q + g
  const uint32_t b_off = static_cast<uint32_t>(C::A_BYTES + (t4 * C::BN_PAD + warp_n * 32 + g) * 8);
  const uint32_t smem_base = smem_u32(smem);

  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int rt = tile / p.n_col_tiles;
    const int ct = tile - rt * p.n_col_tiles;
    const GemmTile t = p.tiles[rt];
    if (!p.state[t.scen].active) continue;
    const sos_scenario sc = p.scen[t.scen];
    const int passes = (t.mix && sc.coef_mix_aer != 0.0) ? 2 : 1;
    const double coef_first = t.mix ? sc.coef_mix_atm : sc.coef_atm;
    const double coef_last = passes == 2 ? sc.coef_mix_aer : coef_first;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q][0] = acc[i][q][1] = 0.0;

    for (int pass = 0; pass < passes; ++pass) {
      if (pass == 1) {
        const double r = coef_first / coef_last;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[i][q][0] *= r; acc[i][q][1] *= r; }
      }
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sA = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sB = sA + b_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double fa[8], fb[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) fa[i] = lds64(sA + a_off[j] + static_cast<uint32_t>(i * 8 * 128));
#pragma unroll
          for (int q = 0; q < 4; ++q) fb[q] = lds64(sB + static_cast<uint32_t>((4 * j * C::BN_PAD + 8 * q) * 8));
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) dmma884(acc[i][q][0], acc[i][q][1], fa[i], fb[q]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }

    // ---- epilogue: lane owns J[r][c], J[r][c+1]; a quad writes 64 contiguous bytes ----
    const int col_base = ct * C::BN + warp_n * 32 + 2 * t4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp_m * 64 + 8 * i + g;
      if (r < t.nrows) {
        double* out = p.J + static_cast<size_t>(t.row0 + r) * p.ld;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = col_base + 8 * q;
          if (c + 1 < p.N) {
            *reinterpret_cast<double2*>(out + c) = make_double2(coef_last * acc[i][q][0], coef_last * acc[i][q][1]);
          } else if (c < p.N) {
            out[c] = coef_last * acc[i][q][0];
          }
        }
      }
    }
  }
}

}  // namespace sosgemm
