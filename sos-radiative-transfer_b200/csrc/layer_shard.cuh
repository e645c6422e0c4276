// Layer-block sharding of ONE large grid over the GPUs of a node (BASELINE configs[3]: ~10 000 layers x 1024 mu;
// SURVEY.md 8e, third row).  Rank r owns a contiguous block of scan chunks = layers [row0, row1).
//
//   * The source contraction (SOS_Aer_I1_In.py:62-74) is row-local: a rank contracts its own rows (plus the few halo
//     rows whose J its sweeps read) and needs nothing from its peers.
//   * The layer sweeps (SOS_Aer_I1_In.py:86-129) are the chunked scan of sweep.cuh: the chunk-local pass and the apply pass
//     run on the rank's own chunks; what crosses a block boundary is the chunk aggregates, N doubles per chunk and
//     direction.  Every rank writes its aggregates straight into the aggregate tables of the ranks that chain through them
//     (downward: the ranks below, upward: the ranks above) and every rank then runs the SAME carry chain over the same
//     numbers as the unsharded solve -- the result is bit-identical to it.
//   * After the sweeps a rank writes the rows its neighbours read as halos (the row below a block: the upward recurrence's
//     first trapezoid; the rows above it: the downward one and the tau-window of the |mu| < 0.01 columns,
//     SOS_Aer_In_limit.py:96-107) into the neighbours' I_n fields, and the owners of the TOA / surface rows publish the two
//     convergence ratios of SOS_Aer_main_specular.py:309 to everybody.
//
// Both exchanges are FUSED into the kernels that consume what arrives: plain stores into peer memory (CUDA IPC mappings over
// NVLink / NVSwitch), one system-scope fence, one flag per sender in the receiver's mailbox, and the receiver spins on its
// own memory.
//   phase 0  sweep_carry_cols_kernel (sweep.cuh): CTA b owns 16 mu columns; it pushes its columns of this rank's aggregates to
//            the peers, raises flag (sender, b) there and waits for the peers' flags b only -- no rank-wide barrier, the
//            column groups of the carry chain proceed independently.
//   phase 1  order_end_kernel (sos_abi.cu): one CTA pushes the halo rows and ratios, raises its flag at every peer, waits for
//            all peers (this is the one all-to-all synchronisation per order: it also orders the next order's writes into the
//            aggregate tables and halo rows after this order's reads), then does the convergence bookkeeping.
// No NCCL call and no host round trip per order: 16 KB-class messages are latency, not bandwidth.  Flags carry a
// monotonically increasing epoch kept on the device, so the kernels are replayed from a CUDA graph; once the solve has
// converged (identical state on every rank: the ratios are the same bits everywhere) every kernel returns at once, so ranks
// may run ahead by different numbers of no-op orders without waiting for each other.
#pragma once
#include "common.cuh"

namespace soslayer {

constexpr int COL_GROUP = 16;    // mu columns per CTA of the carry chain = per phase-0 flag
constexpr int MAX_GROUPS = 256;  // N <= 4096

struct Mailbox {
  double* aggD;               // [nchunks][N]
  double* aggU;               // [nchunks][N]
  double* ratios;             // [2]: ratio_toa (from the owner of row 0), ratio_surf (from the owner of row L-1)
  unsigned long long* flags;  // [SOS_MAX_PEERS senders] of phase 1, then [SOS_MAX_PEERS][MAX_GROUPS] of phase 0
  unsigned long long* epoch;  // [1]: orders exchanged so far (only ever touched by its own GPU)
};

struct LayerPeers {
  int rank, n;                 // n <= 1: not sharded
  Mailbox box[SOS_MAX_PEERS];  // every rank's mailbox as mapped into THIS process
  double* In[SOS_MAX_PEERS];   // every rank's I_n field
  int halo_above;              // rows above row0 whose J this rank's sweeps read (>= 1 unless row0 == 0)
  int next_halo_above;         // ... of the rank below
  unsigned long long timeout_ns;  // 0: never wait (one rank profiled alone)
};

// carve a mailbox out of one allocation; returns the size in bytes when base == nullptr
inline size_t mailbox_layout(void* base, int nchunks, int N, Mailbox* out) {
  const size_t nagg = static_cast<size_t>(nchunks) * N;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 127) / 128 * 128; return o; };
  const size_t oD = take(nagg * 8), oU = take(nagg * 8), oR = take(2 * 8), oF = take((SOS_MAX_PEERS + SOS_MAX_PEERS * MAX_GROUPS) * 8), oE = take(8);
  if (base && out) {
    char* b = static_cast<char*>(base);
    out->aggD = reinterpret_cast<double*>(b + oD);
    out->aggU = reinterpret_cast<double*>(b + oU);
    out->ratios = reinterpret_cast<double*>(b + oR);
    out->flags = reinterpret_cast<unsigned long long*>(b + oF);
    out->epoch = reinterpret_cast<unsigned long long*>(b + oE);
  }
  return off;
}

__device__ __forceinline__ unsigned long long* flag_phase1(const Mailbox& m, int sender) { return m.flags + sender; }
__device__ __forceinline__ unsigned long long* flag_phase0(const Mailbox& m, int sender, int group) {
  return m.flags + SOS_MAX_PEERS + sender * MAX_GROUPS + group;
}

// One full warp: lane q waits until flag(q) has reached epoch e.  Returns true (in every lane) if a peer did not show up
// within the time limit (a crashed rank must not hang the GPU).
template <typename FlagOf>
__device__ __forceinline__ bool wait_flags(const LayerPeers& lp, unsigned long long e, FlagOf flag_of) {
  const int lane = threadIdx.x & 31;
  bool late = false;
  if (lane < lp.n && lane != lp.rank && lp.timeout_ns > 0) {
    const volatile unsigned long long* f = flag_of(lane);
    const unsigned long long t0 = sos_global_timer_ns();
    while (*f < e) {
      if (sos_global_timer_ns() - t0 > lp.timeout_ns) { late = true; break; }
      __nanosleep(32);
    }
  }
  __threadfence_system();
  return __any_sync(0xffffffffu, late);
}

// ---- phase 0, called by every thread of carry-chain CTA `group` (columns [group * 32, +32)); contains __syncthreads ----
// Pushes this rank's aggregates of those columns to the ranks that chain through them, raises the flags, waits for the
// peers'.  Returns true if a peer was late.
__device__ __forceinline__ bool exchange_aggregates(const GridDev& g, const LayerPeers& lp, int group) {
  __shared__ int s_late;
  const int me = lp.rank, N = g.N;
  const Mailbox& mine = lp.box[me];
  const unsigned long long e = *mine.epoch + 1;  // (advanced by order_end_kernel, later on this stream)
  const int tx = threadIdx.x & (COL_GROUP - 1), ty = threadIdx.x / COL_GROUP, rows = blockDim.x / COL_GROUP;
  const int m = group * COL_GROUP + tx;
  if (m < N) {
    for (int c = g.c_lo + ty; c < g.c_hi; c += rows) {
      const size_t o = static_cast<size_t>(c) * N + m;
      const double d = mine.aggD[o], u = mine.aggU[o];
      for (int q = me + 1; q < lp.n; ++q) lp.box[q].aggD[o] = d;  // the downward chain of the ranks below runs through these
      for (int q = 0; q < me; ++q) lp.box[q].aggU[o] = u;         // the upward chain of the ranks above
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the barrier orders the CTA's stores before this thread; its system-scope fence (cumulative) orders them before the flags
    __threadfence_system();
    for (int q = 0; q < lp.n; ++q)
      if (q != me) *reinterpret_cast<volatile unsigned long long*>(flag_phase0(lp.box[q], me, group)) = e;
  }
  if (threadIdx.x < 32) {
    const bool late = wait_flags(lp, e, [&](int q) { return flag_phase0(mine, q, group); });
    if (threadIdx.x == 0) s_late = late ? 1 : 0;
  }
  __syncthreads();
  return s_late != 0;
}

__device__ __forceinline__ void copy16(double* dst, const double* src, size_t n_doubles) {
  // both sides are 16-byte aligned (rows of ld doubles, ld even).  Eight independent 16-byte loads per thread in flight, then
  // the (posted) stores: the copy is latency, not bandwidth
  double2* d = reinterpret_cast<double2*>(dst);
  const double2* s = reinterpret_cast<const double2*>(src);
  const size_t n = n_doubles / 2, step = static_cast<size_t>(blockDim.x) * 8;
  for (size_t i0 = threadIdx.x; i0 < n; i0 += step) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t i = i0 + static_cast<size_t>(u) * blockDim.x;
      if (i < n) v[u] = s[i];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t i = i0 + static_cast<size_t>(u) * blockDim.x;
      if (i < n) d[i] = v[u];
    }
  }
}

// ---- phase 1, called by every thread of the (single) CTA of order_end_kernel; contains __syncthreads ----
// Halo rows of I_n to the two neighbours, convergence ratios to everybody (incl. this rank's own mailbox), flags, wait for all
// peers; thread 0 then advances the epoch.  Returns true if a peer was late.
__device__ __forceinline__ bool exchange_halos(const GridDev& g, const LayerPeers& lp) {
  __shared__ int s_late;
  const int me = lp.rank;
  const Mailbox& mine = lp.box[me];
  const unsigned long long e = *mine.epoch + 1;
  const double* src = lp.In[me];
  if (me > 0) {  // the row below the upper neighbour's block = this rank's first row
    const size_t o = static_cast<size_t>(g.row0) * g.ld;
    copy16(lp.In[me - 1] + o, src + o, g.ld);
  }
  if (me + 1 < lp.n) {  // the rows above the lower neighbour's block = this rank's last rows
    const int h = lp.next_halo_above;
    const size_t o = static_cast<size_t>(g.row1 - h) * g.ld;
    copy16(lp.In[me + 1] + o, src + o, static_cast<size_t>(h) * g.ld);
  }
  if (static_cast<int>(threadIdx.x) < lp.n) {
    if (g.row0 == 0) lp.box[threadIdx.x].ratios[0] = g.state[0].ratio_toa;
    if (g.row1 == g.L) lp.box[threadIdx.x].ratios[1] = g.state[0].ratio_surf;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int q = 0; q < lp.n; ++q)
      if (q != me) *reinterpret_cast<volatile unsigned long long*>(flag_phase1(lp.box[q], me)) = e;
  }
  if (threadIdx.x < 32) {
    const bool late = wait_flags(lp, e, [&](int q) { return flag_phase1(mine, q); });
    if (threadIdx.x == 0) {
      s_late = late ? 1 : 0;
      *mine.epoch = e;
    }
  }
  __syncthreads();
  return s_late != 0;
}

}  // namespace soslayer
