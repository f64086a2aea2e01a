// tcgen05 tensor-core path (placeholder until the kernel lands): nothing is eligible.
#pragma once
#include "common.cuh"
namespace curv {
static inline bool tc_gather_eligible(const Geom&) { return false; }
static inline int tc_launch_gather_gemm(const GatherGemmArgs&, int, cudaStream_t) { return -1; }
}  // namespace curv
