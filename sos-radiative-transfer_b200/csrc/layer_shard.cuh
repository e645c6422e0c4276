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
// Both exchanges are plain stores into peer memory (CUDA IPC mappings over NVLink / NVSwitch) issued by the kernels below,
// followed by a system-scope fence and one flag per (phase, sender) in the receiver's mailbox; the receiver spins on its own
// memory (inside the kernel that consumes the data).  No NCCL call and no host round trip per order: 16 KB-class messages
// are latency, not bandwidth.  Flags carry a monotonically increasing epoch kept on the device, so the kernels can be replayed
// from a CUDA graph; once the solve has converged (identical state on every rank: the ratios are the same bits everywhere)
// every kernel returns at once, so ranks may run ahead by different numbers of no-op orders without waiting for each other.
#pragma once
#include "common.cuh"

namespace soslayer {

constexpr int PUSH_THREADS = 1024;

struct Mailbox {
  double* aggD;               // [nchunks][N]
  double* aggU;               // [nchunks][N]
  double* ratios;             // [2]: ratio_toa (from the owner of row 0), ratio_surf (from the owner of row L-1)
  unsigned long long* flags;  // [2 phases][SOS_MAX_PEERS senders]
  unsigned long long* epoch;  // [2 phases]: exchanges completed so far (only ever touched by its own GPU)
};

struct LayerPeers {
  int rank, n;
  Mailbox box[SOS_MAX_PEERS];  // every rank's mailbox as mapped into THIS process
  double* In[SOS_MAX_PEERS];   // every rank's I_n field
  int halo_above;              // rows above row0 whose J this rank's sweeps read (>= 1 unless row0 == 0)
  int next_halo_above;         // ... of the rank below
};

// carve a mailbox out of one allocation; returns the size in bytes when base == nullptr
__host__ __device__ inline size_t mailbox_layout(void* base, int nchunks, int N, Mailbox* out) {
  const size_t nagg = static_cast<size_t>(nchunks) * N;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 127) / 128 * 128; return o; };
  const size_t oD = take(nagg * 8), oU = take(nagg * 8), oR = take(2 * 8), oF = take(2 * SOS_MAX_PEERS * 8), oE = take(2 * 8);
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

__device__ __forceinline__ void copy16(double* dst, const double* src, size_t n_doubles) {
  // both sides are 16-byte aligned (rows of ld doubles, ld even; aggregate rows of N = 2M doubles)
  double2* d = reinterpret_cast<double2*>(dst);
  const double2* s = reinterpret_cast<const double2*>(src);
  // eight independent 16-byte loads per thread in flight, then the (posted) stores: the copy is latency, not bandwidth
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

// One CTA per receiving rank q.  phase 0 (after the chunk-local pass): this rank's chunk aggregates.  phase 1 (after the
// apply / zone passes): halo rows of I_n to the two neighbours, convergence ratios to everybody.
__global__ void __launch_bounds__(PUSH_THREADS) layer_push_kernel(const GridDev g, const LayerPeers lp, int phase) {
  if (!g.state[0].active) return;
  const int q = blockIdx.x, me = lp.rank;
  const Mailbox& mine = lp.box[me];
  const Mailbox& theirs = lp.box[q];
  const unsigned long long e = mine.epoch[phase] + 1;  // (advanced by order_end_kernel, later on this stream)
  const int N = g.N;
  if (phase == 0) {
    if (q != me) {
      const size_t o = static_cast<size_t>(g.c_lo) * N, n = static_cast<size_t>(g.c_hi - g.c_lo) * N;
      if (q > me) copy16(theirs.aggD + o, mine.aggD + o, n);  // the downward chain of the ranks below runs through these
      else copy16(theirs.aggU + o, mine.aggU + o, n);         // the upward chain of the ranks above
    }
  } else {
    const double* src = lp.In[me];
    double* dst = lp.In[q];
    if (q == me - 1) {
      copy16(dst + static_cast<size_t>(g.row0) * g.ld, src + static_cast<size_t>(g.row0) * g.ld, g.ld);
    } else if (q == me + 1) {
      const int h = lp.next_halo_above;
      copy16(dst + static_cast<size_t>(g.row1 - h) * g.ld, src + static_cast<size_t>(g.row1 - h) * g.ld, static_cast<size_t>(h) * g.ld);
    }
    if (threadIdx.x == 0) {
      if (g.row0 == 0) theirs.ratios[0] = g.state[0].ratio_toa;
      if (g.row1 == g.L) theirs.ratios[1] = g.state[0].ratio_surf;
    }
  }
  if (q == me) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    // the barrier orders the CTA's stores before this thread; its system-scope fence (cumulative) then orders them before the flag
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(theirs.flags + phase * SOS_MAX_PEERS + me) = e;
  }
}

// The waits live in the kernels that consume what arrived: sweep_carry_cols_kernel (sweep.cuh) waits for the peers'
// aggregates (phase 0), order_end_kernel (sos_abi.cu) for the halo rows and ratios (phase 1); the latter then takes the two
// ratios over as this rank's convergence state and advances both epochs (one exchange of each phase per order).
inline LayerWait wait_for(const LayerPeers& lp, int phase, unsigned long long timeout_ns) {
  LayerWait w;
  w.flags = lp.n > 1 ? lp.box[lp.rank].flags + phase * SOS_MAX_PEERS : nullptr;
  w.epoch = lp.n > 1 ? lp.box[lp.rank].epoch + phase : nullptr;
  w.n = lp.n;
  w.me = lp.rank;
  w.timeout_ns = timeout_ns;
  return w;
}

}  // namespace soslayer
