// Host-callable launchers implemented in rmp2_kernels.cu, used by the C ABI in rmp2_api.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "rmp2_tables.h"

#ifndef RMP2_BLOCK_THREADS
#define RMP2_BLOCK_THREADS 128
#endif

#ifndef RMP2_SPHERES_BLOCK
#define RMP2_SPHERES_BLOCK 128      // threads per block of rmp2_spheres_kernel = E environments x L obstacle leaves
#endif

struct LeafVec {
  float v[3 * RMP2_MAX_JOINTS];
};

int rmp2_pick_width(int n);
size_t rmp2_step_smem(const StepTables& T, int block);

cudaError_t rmp2_launch_frames(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream);
size_t rmp2_spheres_smem(const SphereTables& ST, int n_spheres, bool use_tma, bool early_out);
cudaError_t rmp2_launch_spheres(const SphereTables& ST, const StepArgs& A, bool use_tma, cudaStream_t stream);
cudaError_t rmp2_launch_step(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream);
cudaError_t rmp2_launch_resolve(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream);
cudaError_t rmp2_launch_fallback(const StepTables& T, const StepArgs& A, int max_blocks, cudaStream_t stream);
cudaError_t rmp2_launch_pinv(int n, float rcond, bool pivot, int mode, long long B, const float* M, const float* f,
                             float* x, cudaStream_t stream);
// which: 0 frames, 1 spheres, 2 step (fused resolve), 3 step (split), 4 resolve, 5 resolve fallback (Jacobi)
cudaError_t rmp2_kernel_attributes(int n, int which, bool use_tma, int block, size_t smem, int* regs,
                                   int* blocks_per_sm);
cudaError_t rmp2_launch_feed(const StepTables& T, const FeedLinks& LK, const FeedArgs& A, cudaStream_t stream);
cudaError_t rmp2_launch_fk(const StepTables& T, long long B, const float* q, const float* qd, float* x, float* xd,
                           float* J, float* c, cudaStream_t stream);
cudaError_t rmp2_launch_leaf(const LeafTab& L, const LeafVec& V, int m, long long K, const float* x, const float* xd,
                             const float* aux, float* xdd, float* M, cudaStream_t stream);
