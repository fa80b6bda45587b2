// Stand-alone check of the TMA box load used by the tile kernel (4-D fp32 tensor map, zero OOB fill).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct alignas(64) Maps { CUtensorMap t; };
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__constant__ int dBW, dBH, dBC;
static int BW = 36, BH = 20, BC = 3, RANK = 4;
__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap* gmap, const float* src, int variant, float* out, int x0, int y0, int b, int* status) {
  extern __shared__ __align__(128) float sm[];
  const uint32_t mbar = smem_addr(sm + 60);
  float* dst = sm + 64;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((uint32_t)(dBW * dBH * dBC * 4)) : "memory");
    if (variant == 1) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(dst)), "l"(src), "r"((uint32_t)(dBW * dBH * dBC * 4)), "r"(mbar) : "memory");
    } else if (variant == 3) {
      asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(smem_addr(dst)), "l"((uint64_t)&maps.t), "r"(x0), "r"(y0), "r"(0), "r"(b), "r"(mbar) : "memory");
    } else if (variant == 4) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_addr(dst)), "l"((uint64_t)&maps.t), "r"(x0), "r"(y0), "r"(mbar) : "memory");
    } else {
      const uint64_t mp = variant == 2 ? (uint64_t)gmap : (uint64_t)&maps.t;
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(smem_addr(dst)), "l"(mp), "r"(x0), "r"(y0), "r"(0), "r"(b), "r"(mbar) : "memory");
    }
  }
  __syncthreads();
  int ok = 0;
  for (int it = 0; it < (1 << 20); ++it) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
    if (done) { ok = 1; break; }
  }
  if (threadIdx.x == 0) *status = ok;
  if (ok) for (int i = threadIdx.x; i < dBW * dBH * dBC; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  if (argc > 5) { RANK = atoi(argv[2]); BW = atoi(argv[3]); BH = atoi(argv[4]); BC = atoi(argv[5]); }
  cudaMemcpyToSymbol(dBW, &BW, 4); cudaMemcpyToSymbol(dBH, &BH, 4); cudaMemcpyToSymbol(dBC, &BC, 4);
  const int B = 2, H = 32, W = 64;
  std::vector<float> h((size_t)B * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o; int* st;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 256 * 256 * 4); cudaMalloc(&st, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, ptr);
  Maps maps;
  const cuuint64_t dims[4] = {W, H, 3, B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BC, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)ptr)(&maps.t, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)RANK, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d  rank %d box %d %d %d\n", (int)r, RANK, BW, BH, BC);
  for (int i = 0; i < 16; ++i) printf("%016llx ", (unsigned long long)maps.t.opaque[i]); printf("\n");
  CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &maps.t, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int t = 0; t < 2; ++t) {
    const int x0 = t == 0 ? -2 : 30, y0 = t == 0 ? -2 : 14, b = t;
    k<<<1, 128, 64 * 1024>>>(maps, gmap, d, variant, o, x0, y0, b, st);
    e = cudaDeviceSynchronize();
    int s = -1; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
    std::vector<float> r2(BW * BH * BC);
    cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
      const int gy = y0 + y, gx = x0 + x;
      const float want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0.f : h[((size_t)(b * 3 + c) * H + gy) * W + gx];
      if (r2[(c * BH + y) * BW + x] != want) ++bad;
    }
    printf("variant %d launch %d: err=%s status=%d mismatches=%d first=%g\n", variant, t, cudaGetErrorString(e), s, bad, r2[0]);
  }
  return 0;
}
