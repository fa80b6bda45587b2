// Walk from the libcu++ TMA load (works) to the hand-written PTX path one difference at a time.
// argv[1] bitmask: 1 = L2_128B promotion, 2 = hand-written PTX barrier/copy (count 1, try_wait), 4 = dynamic smem,
//                  8 = negative start coordinates, 16 = 4-D map / box with 3 channels
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 36, BH = 20;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int flags, int x, int y, float* out, int nch) {
  __shared__ alignas(128) float sbuf[3 * BH * BW + 64];
  extern __shared__ __align__(128) float dyn[];
  float* buf = (flags & 4) ? dyn + 64 : sbuf + 64;
  const uint32_t bytes = BW * BH * nch * 4;
  if (flags & 2) {
    const uint32_t mbar = smem_addr(((flags & 4) ? dyn : sbuf) + 60);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      if (flags & 16)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(smem_addr(buf)), "l"((uint64_t)&tensor_map), "r"(x), "r"(y), "r"(0), "r"(0), "r"(mbar) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_addr(buf)), "l"((uint64_t)&tensor_map), "r"(x), "r"(y), "r"(mbar) : "memory");
    }
    __syncthreads();
    for (int it = 0; it < (1 << 20); ++it) {
      uint32_t done;
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
      if (done) break;
    }
  } else {
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
      if (flags & 16) cde::cp_async_bulk_tensor_4d_global_to_shared(buf, &tensor_map, x, y, 0, 0, bar);
      else cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &tensor_map, x, y, bar);
      token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else {
      token = bar.arrive();
    }
    bar.wait(std::move(token));
  }
  for (int i = threadIdx.x; i < BW * BH * nch; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  const int flags = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 2, H = 96, W = 64, nch = (flags & 16) ? 3 : 1;
  std::vector<float> h((size_t)B * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * 3 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[4] = {W, H, 3, B}; const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)3 * W * H * 4};
  const cuuint32_t box[4] = {BW, BH, (cuuint32_t)nch, 1}; const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)ptr)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (flags & 16) ? 4 : 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    (flags & 1) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int x0 = (flags & 32) ? -4 : ((flags & 8) ? -2 : 8), y0 = (flags & (8 | 32)) ? -2 : 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<1, 128, (flags & 4) ? 64 * 1024 : 0>>>(map, flags, x0, y0, o, nch);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> r2(BW * BH * nch); cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < nch; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
    const int gy = y0 + y, gx = x0 + x;
    const float want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0.f : h[((size_t)c * H + gy) * W + gx];
    if (r2[(c * BH + y) * BW + x] != want) ++bad;
  }
  printf("flags %2d encode %d: err=%s mismatches=%d\n", flags, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
