// FP64 peak micro-benchmarks for B200 (sm_100a): DFMA (vector pipe) and DMMA
// (mma.sync f64, the only FP64 tensor path -- tcgen05.mma has no f64 kind).
// MEASURED_PEAKS.json carries no FP64 figure, so the source-contraction kernel's
// roofline denominator is measured here, on the box, with CUDA events.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
//   ./fp64_peak            -> one JSON object on stdout
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync.aligned.m8n8k4.row.col.f64: D(8x8) += A(8x4) * B(4x8); per thread a:1 b:1 c:2 doubles
template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync.aligned.m16n8k8.row.col.f64 (sm_90+): a:4 b:2 c:4 doubles per thread
template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync.aligned.m16n8k16.row.col.f64 (sm_90+): a:8 b:4 c:4 doubles per thread
template <int NACC>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                     "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                       "d"(a), "d"(b), "d"(a), "d"(b));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int bps = 1; bps <= 4; bps *= 2) {  // blocks per SM (8, 16, 32 warps/SM)
    int grid = sms * bps;
    {
      double ms = time_ms([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 8 * 8 * (double)iters * 256 * grid;
      printf(", \"dfma_ilp8_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 8 * 8 * 4 * 4.0 * 8 * (double)iters * 8 * grid;  // per warp-mma 512 flop
      printf(", \"dmma_m8n8k4_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma1688_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 16 * 8 * 8 * 4.0 * 8 * (double)iters * 8 * grid;
      printf(", \"dmma_m16n8k8_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma16816_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 16 * 8 * 16 * 2.0 * 8 * (double)iters * 8 * grid;
      printf(", \"dmma_m16n8k16_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9);
    }
  }
  // sustained DFMA (about 2 s back to back) -- the figure for a kernel timed inside a long step
  {
    int grid = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 400; ++n) dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * 8 * (double)iters * 256 * grid * n;
    printf(", \"dfma_sustained_tflops\": %.3f, \"dfma_sustained_seconds\": %.2f", fl / ms * 1e-9, ms * 1e-3);
  }
  {
    int grid = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 400; ++n) dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * 8 * 4 * 4.0 * 8 * (double)iters * 8 * grid * n;
    printf(", \"dmma_m8n8k4_sustained_tflops\": %.3f, \"dmma_sustained_seconds\": %.2f", fl / ms * 1e-9, ms * 1e-3);
  }
  printf("}\n");
  return 0;
}
