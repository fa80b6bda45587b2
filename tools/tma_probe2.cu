// CUDA programming guide style TMA load (libcu++), to diff against the hand-written PTX path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 32, BH = 16;
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float* out) {
  __shared__ alignas(128) float smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}
int main() {
  const int H = 96, W = 64;
  std::vector<float> h((size_t)H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[2] = {W, H}; const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {BW, BH}; const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)ptr)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  k<<<1, 128>>>(map, 8, 4, o);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> r2(BW * BH); cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) if (r2[y * BW + x] != h[(size_t)(4 + y) * W + 8 + x]) ++bad;
  printf("libcu++ path: err=%s mismatches=%d\n", cudaGetErrorString(e), bad);
  return 0;
}
