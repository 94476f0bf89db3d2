// knn2_tc.cu -- tcgen05 engine of the brute-force 2-NN search (placeholder until the tensor
// engine lands; the POPC engine is selected automatically while this reports "unsupported").
#include "internal.cuh"

bool knn2_tc_supported() { return false; }

int knn2_tc_run(orbgpu_ctx *, const orbgpu_db *, int64_t, const uint4 *, uint64_t *, uint32_t *, int *, int64_t)
{
    return orbgpu_fail(ORBGPU_ERR_INVALID, "tcgen05 engine not built into this library");
}
