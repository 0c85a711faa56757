// tcgen05 / TMEM / TMA GEMM path (sm_100a). Placeholder interface: filled in by the tensor-core milestone.
#pragma once
#include "common.cuh"

namespace ptts {

struct TcPlanCache;
inline TcPlanCache* tc_plan_cache_create() { return nullptr; }
inline void tc_plan_cache_destroy(TcPlanCache*) {}

template <typename T>
inline bool tc_gemm_supported(int R, int N, int K, const RowMap& amap, int a_rps) { return false; }

template <typename T>
inline int tc_gemm_launch(TcPlanCache*, const T* A, RowMap amap, int a_rps, const T* W, int R, int N, int K, const Epi& epi, cudaStream_t stream) { return 0; }

}  // namespace ptts
