// Tree-specialised kernels: the frames and step kernels of ONE compiled tree, rebuilt at run time by NVRTC
// with the tree's tables as a compile-time constant (see rmp2_tree_kernels.cuh).  Host side only.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "rmp2_tables.h"

struct SpecModule;   // opaque: the loaded module and its kernels

// Compile (and, unless compile_only, load on the current device) the specialised kernels of `T` for kernel
// width `width`.  Returns 0 and a module (nullptr when compile_only), or nonzero with `err` filled in.
int rmp2_jit_build(const StepTables& T, int width, bool compile_only, SpecModule** out, std::string& err);
void rmp2_jit_destroy(SpecModule* m);
double rmp2_jit_seconds(const SpecModule* m);

// which: 0 frames, 1 step (resolve fused), 2 step (split), 3 step (resolve fused) compiled for exactly
// rmp2_jit_big_block() threads per block.  Same grid/block/shared-memory geometry as the generic launchers in
// rmp2_kernels.cu.
cudaError_t rmp2_jit_launch(const SpecModule* m, int which, const StepArgs& A, unsigned blocks, unsigned block,
                            size_t smem, cudaStream_t stream, std::string& err);
int rmp2_jit_registers(const SpecModule* m, int which);
// Block size the module's `which = 3` kernel was compiled for (0: that kernel must not be used, e.g. the tuning hook
// RMP2_JIT_EXTRA redefined the block size).
int rmp2_jit_big_block(const SpecModule* m, int width);
