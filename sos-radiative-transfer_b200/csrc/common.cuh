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

#define SOS_MAX_PHASE 16

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
  const double* W;          // extrapolation matrices
  int widx[4], wns[4], woff[4];
  int first_small;          // first downward column with |mu| < MU_THRESHOLD (M-1 if none)
};

// One row-tile of the source contraction.
struct GemmTile {
  int row0;    // first stacked row
  int nrows;   // valid rows (<= BM)
  int scen;
  int mix;     // 0: rows outside the aerosol region, 1: aerosol rows (two operands)
};
