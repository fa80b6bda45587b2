// md2_nvtx.h - NVTX ranges around the C entry points (tracing).  nvtx3 is header-only and a no-op unless
// a profiler is attached; the ranges name the calls on Nsight timelines.
#ifndef MD2_NVTX_H_
#define MD2_NVTX_H_
#include <nvtx3/nvToolsExt.h>

namespace md2 {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace md2
#endif  // MD2_NVTX_H_
