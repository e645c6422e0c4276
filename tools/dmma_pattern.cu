// How close to the DMMA peak can a GEMM-like instruction stream get?  The peak loop (tools/fp64_peak.cu) feeds every
// DMMA the same A and B registers; a real register-tiled contraction gives every DMMA of an (MB x NB) warp tile its own
// (A_i, B_q) pair and accumulator.  This probe issues exactly that stream from registers only (no memory, no barriers)
// with 8 warps per SM (2 per scheduler, like the contraction kernels), for several warp-tile shapes.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MB, int NB, int SETS, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) tile_kernel(double* out, int iters) {
  double acc[SETS][MB][NB][2], fa[SETS][MB], fb[SETS][NB];
#pragma unroll
  for (int s = 0; s < SETS; ++s) {
#pragma unroll
    for (int i = 0; i < MB; ++i) fa[s][i] = 1.0 + 1e-3 * (threadIdx.x + 7 * i + 3 * s);
#pragma unroll
    for (int q = 0; q < NB; ++q) fb[s][q] = 1e-9 * (threadIdx.x + 5 * q + s);
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int q = 0; q < NB; ++q) acc[s][i][q][0] = acc[s][i][q][1] = 0.0;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int q = 0; q < NB; ++q)
#pragma unroll
        for (int s = 0; s < SETS; ++s)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[s][i][q][0]), "+d"(acc[s][i][q][1]) : "d"(fa[s][i]), "d"(fb[s][q]));
  }
  double t = 0;
#pragma unroll
  for (int s = 0; s < SETS; ++s)
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int q = 0; q < NB; ++q) t += acc[s][i][q][0] + acc[s][i][q][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MB, int NB, int SETS, int WARPS>
static int run(const char* name, int sms, double* out) {
  const int iters = 2048;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) tile_kernel<MB, NB, SETS, WARPS><<<sms, 32 * WARPS>>>(out, iters);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    tile_kernel<MB, NB, SETS, WARPS><<<sms, 32 * WARPS>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * 256 * MB * NB * SETS * (double)iters * WARPS * sms;
  printf("{\"pattern\": \"%s\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", name, WARPS, flops / best * 1e-9);
  return 0;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * sms));
  if (run<1, 8, 1, 8>("1x8 blocks, shared A (peak-loop like)", sms, out)) return 1;
  if (run<4, 4, 1, 8>("4x4 blocks, 1 accumulator set (general kernel warp tile)", sms, out)) return 1;
  if (run<4, 4, 2, 8>("4x4 blocks, 2 accumulator sets (folded kernel warp tile)", sms, out)) return 1;
  if (run<4, 6, 1, 12>("4x6 blocks, 1 set, 12 warps (128x144 general tile)", sms, out)) return 1;
  if (run<2, 8, 2, 8>("2x8 blocks, 2 sets", sms, out)) return 1;
  if (run<4, 2, 2, 16>("4x2 blocks, 2 sets, 16 warps", sms, out)) return 1;
  if (run<4, 4, 2, 4>("4x4 blocks, 2 sets, 4 warps (1 per scheduler)", sms, out)) return 1;
  return 0;
}
