// Host-side launch helpers shared by the .cu files.
#pragma once
#include <algorithm>
#include "common.cuh"

// thread shape up to this many observations; above it one warp (or a CTA) owns a point
#define PQ_THREAD_SHAPE_MAX_N 128
// keep at least two CTAs per SM when staging constants in shared memory
#define PQ_SMEM_STAGE_LIMIT (100 * 1024)

static inline size_t pq_pick_smem(const peuqse_problem* p, int* use_smem) {
    if (p->obs_smem_bytes <= (size_t)PQ_SMEM_STAGE_LIMIT && p->obs_smem_bytes <= (size_t)p->max_smem_optin) {
        *use_smem = 1;
        return p->obs_smem_bytes;
    }
    *use_smem = 0;
    return 0;
}

static inline cudaError_t pq_allow_smem(const void* kern, size_t smem) {
    // static + dynamic shared memory above 48 KB needs the opt-in; the kernels' static part is at most a few KB
    if (smem <= 40 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
