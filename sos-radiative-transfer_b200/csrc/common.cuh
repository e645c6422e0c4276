// Shared device/host declarations of libsos_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include "../../include/sos_b200.h"

// thresholds of the reference (SOS_Aer_global_va.py:5-7, SOS_Aer_I1_In.py:41,103)
#define SOS_MU_THRESHOLD 0.01
#define SOS_MU_VERY_SMALL 0.001
#define SOS_BLEND_THRESHOLD 0.0001
#define SOS_MU0_TOLERANCE 0.0001

// internal status bit: with a generated source (sweep.cuh: SrcGen) only the first 128 upward columns keep their raw I_n;
// a mu -> 0+ blend that reaches further cannot be finished, and the solve is repeated with every row stored (sos_solve
// returns SOS_ERR_RETRY after switching the plan over)
#define SOS_STATUS_STRIP_FALLBACK 0x100u
// internal status bit: a layer-sharded rank waited more than a few seconds for a peer's flag (layer_shard.cuh)
#define SOS_STATUS_PEER_TIMEOUT 0x200u

#define SOS_MAX_PHASE 16
#define SOS_MAX_GROUPS 48
#define SOS_MAX_PEERS 8

// Per-scenario mutable state (device).
struct ScenState {
  double ratio_toa;
  double ratio_surf;
  int n_orders;
  int active;
  unsigned status;
  int pad;
};

// Everything the kernels need to know about the grid; passed by value.
struct GridDev {
  int L, M, N, S, ld;
  int nreg;
  int rstart[4];
  int surface;
  int nchunks;
  const int* chunk_start;   // [nchunks+1]
  const int* chunk_region;  // [nchunks]
  const int* row_chunk;     // [L]
  const double* mu;         // [N]
  const double* wmu;        // [N] composite-trapezoid weights of the mu grid
  const double* tau;        // [S][L]
  const sos_scenario* scen; // [S]
  ScenState* state;         // [S]
  int* n_active;            // [1]
  int* active_flat;         // [S] ids of the active scenarios, ascending (rebuilt with n_active)
  const double* W;          // extrapolation matrices
  int widx[4], wns[4], woff[4];
  int first_small;          // first downward column with |mu| < MU_THRESHOLD (M-1 if none)
  int col0, col1;           // mu columns this plan owns (mu-block sharding); [0, N) by default
  int row0, row1;           // layers this plan owns (layer-block sharding, layer_shard.cuh); [0, L) by default
  int c_lo, c_hi;           // ... = the scan chunks [c_lo, c_hi); [0, nchunks) by default
};

__device__ __forceinline__ unsigned long long sos_global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Row-tile layout of the source contraction, rebuilt on the device whenever the set of active
// scenarios changes.  A group = scenarios sharing the contraction operand(s); class 1 groups
// (aerosol rows, two operands) come first so that the dynamic scheduler hands out the heavy tiles
// first.  Tile t of group g covers segments [t*SEGS, (t+1)*SEGS) of the concatenation
// "active scenario 0: seg 0..nseg-1, active scenario 1: ...".
struct TilePlan {
  int n_row_tiles;
  int n_groups;
  int group_tile_start[SOS_MAX_GROUPS + 1];
  int group_cls[SOS_MAX_GROUPS];
  int group_nactive[SOS_MAX_GROUPS];
  int group_list_off[SOS_MAX_GROUPS];
  int group_phaseA[SOS_MAX_GROUPS];
  int group_phaseB[SOS_MAX_GROUPS];
};
