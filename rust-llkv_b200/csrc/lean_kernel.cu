// lean_kernel.cu — ahead-of-time build of the lean kernel (lean_kernel.cuh) as an interpreter: the program and layout
// arrive with the __grid_constant__ kernel parameter.  Plans that run repeatedly are recompiled by jit.cpp with the
// program as a compile-time constant; this build is what runs first, and whenever NVRTC is not available.
#include "lean_kernel.cuh"

namespace llkv {

template <int R>
__global__ void __launch_bounds__(288, 1) lean_scan_kernel(const __grid_constant__ LeanPlan p) {
  lean_body<R, LeanDynCfg>(p);
}

cudaError_t launch_lean(const LeanPlan& plan, uint32_t grid, cudaStream_t stream) {
  const uint32_t block = plan.s.nc + 32;
  const uint32_t smem = plan.s.smem_total;
#define LLKV_LAUNCH_LEAN(RR)                                                                                                  \
  do {                                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(lean_scan_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return e;                                                                                           \
    lean_scan_kernel<RR><<<grid, block, smem, stream>>>(plan);                                                                \
    return cudaGetLastError();                                                                                                \
  } while (0)
  if (plan.s.rows_per_thread == 8) LLKV_LAUNCH_LEAN(8);
  if (plan.s.rows_per_thread == 4) LLKV_LAUNCH_LEAN(4);
  if (plan.s.rows_per_thread == 2) LLKV_LAUNCH_LEAN(2);
  LLKV_LAUNCH_LEAN(1);
#undef LLKV_LAUNCH_LEAN
}

}  // namespace llkv
