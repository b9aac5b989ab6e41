// Persistent TMA-pipelined fused step for the production mode (placeholder until the kernel lands).
#pragma once

#include "d3pm_step_rows.cuh"

namespace d3pm {

inline bool stream_kernel_eligible(const StepParams&) { return false; }
inline int launch_step_stream(const StepParams&, cudaStream_t) { return D3PM_ERR_UNSUPPORTED; }

}  // namespace d3pm
