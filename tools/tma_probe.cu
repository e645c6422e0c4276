// Probe of the TMA features the fused order kernel relies on, one feature per run (a fault poisons the context):
//   tma_probe <mode>   0: 3-D tiled load   1: + 1-D bulk load   2: + 3-D tiled store   3: store at a negative column   4: load at negative column
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct Params { CUtensorMap map; const double* lin; double* out; int mode; int c0; int bytes; };

__global__ void probe(const __grid_constant__ Params p) {
  __shared__ __align__(128) double tile[16 * 136];
  __shared__ __align__(16) double lin[8];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bytes = p.bytes + (p.mode >= 1 ? 64 : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)),
                 "l"(reinterpret_cast<uint64_t>(&p.map)), "r"(smem_u32(&bar)), "r"(p.mode == 4 ? -5 : p.c0), "r"(8), "r"(1) : "memory");
    if (p.mode >= 1)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(lin)), "l"(p.lin), "r"(64),
                   "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  double v = tile[threadIdx.x] + (p.mode >= 1 ? lin[threadIdx.x & 7] : 0.0);
  p.out[threadIdx.x] = v;
  tile[threadIdx.x] = v + 1000.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (p.mode >= 2 && threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&p.map)), "r"(smem_u32(tile)),
                 "r"(p.mode == 3 ? -5 : p.c0), "r"(16), "r"(2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int N = 502, ld = 512, L = 96, S = 7;
  std::vector<double> h(static_cast<size_t>(S) * L * ld);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<double>(i % 100000);
  double *d, *lin, *out;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&lin, 1024); cudaMalloc(&out, 1024 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(lin, h.data(), 1024, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Params p;
  const int dummy_ = 0; (void)dummy_;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)L, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)L * ld * 8};
  const int bc = argc > 2 ? atoi(argv[2]) : 128, br = argc > 3 ? atoi(argv[3]) : 8;
  cuuint32_t box[3] = {(cuuint32_t)bc, (cuuint32_t)br, 1}, es[3] = {1, 1, 1};
  p.bytes = bc * br * 8;
  CUresult r = ((PFN_encodeTiled)fp)(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  p.lin = lin; p.out = out; p.mode = mode; p.c0 = argc > 4 ? atoi(argv[4]) : 123;
  probe<<<1, 128>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  double o[4]; cudaMemcpy(o, out, 32, cudaMemcpyDeviceToHost);
  printf("mode %d: %s  out0 %.1f expected %.1f\n", mode, cudaGetErrorString(e), o[0], h[(size_t)(1 * L + 8) * ld + p.c0] + (mode >= 1 ? h[0] : 0));
  if (mode >= 2 && e == cudaSuccess) {
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    printf("  stored[2][16][c0] = %.1f (want out0 + 1000)\n", h[(size_t)(2 * L + 16) * ld + (mode == 3 ? 0 : 123)]);
  }
  return 0;
}
