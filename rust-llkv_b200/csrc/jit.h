// jit.h — run-time specialisation of the lean kernel on one plan shape (jit.cpp).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "plan.h"

namespace llkv {

// Compiles the lean kernel specialised on `shape` to an sm_100a cubin with NVRTC.  Needs no GPU.  Returns 0, or -1 with
// the reason (NVRTC missing, compile log) in `log`.
int jit_compile_cubin(const LeanShape& shape, int ctas_per_sm, std::vector<char>& cubin, std::string& log);

// Launches the specialised kernel for plan.s (compiled and cached per device on first use).  `*used` is false, with
// cudaSuccess returned, when specialisation is not possible (NVRTC missing, compile error): the caller then launches the
// interpreted kernel instead.
// compiles (or finds) the specialised kernel of the plan's shape; false = NVRTC missing or the compile failed
bool jit_ready(int device, const LeanPlan& plan, int ctas_per_sm);
cudaError_t jit_launch(int device, const LeanPlan& plan, int ctas_per_sm, uint32_t grid, cudaStream_t stream, bool* used, std::string* why);

}  // namespace llkv
