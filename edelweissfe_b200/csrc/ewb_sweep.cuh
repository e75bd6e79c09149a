// Fused BoxGen sweep assembly (placeholder until the kernel lands; the generic path is used).
#pragma once
#include "ewb_tile.cuh"
struct ewb_buffers;
namespace ewb {
struct SweepPlan {
    int64_t nX = 0, nY = 0, nZ = 0;
    int build(int64_t nx, int64_t ny, int64_t nz) { nX = nx; nY = ny; nZ = nz; return 0; }
    void release() {}
    int launch(int, int, const MatParams&, const ewb_buffers*, int*, int, cudaStream_t, int*) { return -3; }
};
}  // namespace ewb
