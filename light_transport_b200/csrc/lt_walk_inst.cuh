// lt_walk_inst.cuh — host dispatch over the (traversal, semantics) kernel variants; included once
// per arithmetic type by lt_walk_f32.cu / lt_walk_f64.cu so the two compile in parallel (and so the
// float64 unit can be built with -fmad=false: its products and sums must round like the reference's
// separate NumPy operations, see DESIGN.md "numerics").
#pragma once
#include "lt_megakernel.cuh"
#include "lt_wavefront.cuh"
#include "lt_record.cuh"

namespace lt {

template <typename R, int TRAV, int SEM, int PH>
static int launch_ph(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    auto k = walk_megakernel<R, TRAV, SEM, PH>;
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, LT_BLOCK, smem, stream>>>(P);
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}
// PH: phase-aligned iterations (lt_megakernel.cuh) — a compile-time variant because the aligned loop relies on the
// query kind being warp-uniform.  The LBVH walk is never aligned.
template <typename R, int TRAV, int SEM>
static int launch_one(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    if (TRAV != LT_TRAV_LBVH && TRAV != LT_TRAV_LBVH2 && !P.phased) return launch_ph<R, TRAV, SEM, 0>(P, grid, smem, stream);
    return launch_ph<R, TRAV, SEM, 1>(P, grid, smem, stream);
}

template <typename R, int TRAV>
static int launch_sem(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    if (P.semantics == LT_SEM_REFERENCE) return launch_one<R, TRAV, LT_SEM_REFERENCE>(P, grid, smem, stream);
    return launch_one<R, TRAV, LT_SEM_ITERATIVE>(P, grid, smem, stream);
}

template <typename R>
int launch_walk(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    switch (P.traversal) {
        case LT_TRAV_COMPAT: return launch_sem<R, LT_TRAV_COMPAT>(P, grid, smem, stream);
        case LT_TRAV_NEAREST: return launch_sem<R, LT_TRAV_NEAREST>(P, grid, smem, stream);
        case LT_TRAV_LBVH: return launch_sem<R, LT_TRAV_LBVH>(P, grid, smem, stream);
        case LT_TRAV_LBVH2: return launch_sem<R, LT_TRAV_LBVH2>(P, grid, smem, stream);
    }
    set_error("unknown traversal %d", P.traversal);
    return LT_E_INVALID;
}

template <typename R, int TRAV, int SEM, int PH>
static int attrs_ph(int* regs, int* local_bytes, int* max_blocks, size_t smem) {
    auto k = walk_megakernel<R, TRAV, SEM, PH>;
    cudaFuncAttributes a;
    LT_CUDA_OK(cudaFuncGetAttributes(&a, k));
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (cudaSharedmemCarveoutMaxL1 for the LBVH walk, which keeps nothing in shared memory, was measured: 2.42 vs 2.75 Grays/s
    //  on the C3 probe — the default carveout stays)
    *regs = a.numRegs;
    *local_bytes = (int)a.localSizeBytes;
    LT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks, k, LT_BLOCK, smem));
    return LT_OK;
}

// the variant launch_one() will launch: PH = 0 only exists for the shared-memory scenes
template <typename R, int TRAV, int SEM>
static int attrs_one(int phased, int* regs, int* local_bytes, int* max_blocks, size_t smem) {
    if (TRAV != LT_TRAV_LBVH && TRAV != LT_TRAV_LBVH2 && !phased) return attrs_ph<R, TRAV, SEM, 0>(regs, local_bytes, max_blocks, smem);
    return attrs_ph<R, TRAV, SEM, 1>(regs, local_bytes, max_blocks, smem);
}

template <typename R>
int walk_kernel_attrs(int trav, int sem, int phased, int* regs, int* local_bytes, int* max_blocks, size_t smem) {
#define LT_ATTR(T, S) \
    if (trav == T && sem == S) return attrs_one<R, T, S>(phased, regs, local_bytes, max_blocks, smem);
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_REFERENCE)
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_ITERATIVE)
    LT_ATTR(LT_TRAV_NEAREST, LT_SEM_REFERENCE)
    LT_ATTR(LT_TRAV_NEAREST, LT_SEM_ITERATIVE)
    LT_ATTR(LT_TRAV_LBVH, LT_SEM_REFERENCE)
    LT_ATTR(LT_TRAV_LBVH, LT_SEM_ITERATIVE)
    LT_ATTR(LT_TRAV_LBVH2, LT_SEM_REFERENCE)
    LT_ATTR(LT_TRAV_LBVH2, LT_SEM_ITERATIVE)
#undef LT_ATTR
    set_error("unknown kernel variant");
    return LT_E_INVALID;
}

template <typename R, int TRAV, int SEM>
static int wf_launch_one(const WalkParams<R>& P, unsigned long long* gc, int grid, size_t smem, cudaStream_t stream) {
    auto k = walk_wavefront<R, TRAV, SEM>;
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, WfCfg<R>::NT, smem, stream>>>(P, gc);
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}
template <typename R, int TRAV, int SEM>
static int wf_attrs_one(int* regs, int* max_blocks, size_t smem) {
    auto k = walk_wavefront<R, TRAV, SEM>;
    cudaFuncAttributes a;
    LT_CUDA_OK(cudaFuncGetAttributes(&a, k));
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (cudaSharedmemCarveoutMaxL1 for the LBVH walk, which keeps nothing in shared memory, was measured: 2.42 vs 2.75 Grays/s
    //  on the C3 probe — the default carveout stays)
    *regs = a.numRegs;
    LT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks, k, WfCfg<R>::NT, smem));
    return LT_OK;
}
#define LT_WF_DISPATCH(CALL)                                                                              \
    if (trav == LT_TRAV_COMPAT && sem == LT_SEM_REFERENCE) return CALL(LT_TRAV_COMPAT, LT_SEM_REFERENCE);   \
    if (trav == LT_TRAV_COMPAT && sem == LT_SEM_ITERATIVE) return CALL(LT_TRAV_COMPAT, LT_SEM_ITERATIVE);   \
    if (trav == LT_TRAV_NEAREST && sem == LT_SEM_REFERENCE) return CALL(LT_TRAV_NEAREST, LT_SEM_REFERENCE); \
    if (trav == LT_TRAV_NEAREST && sem == LT_SEM_ITERATIVE) return CALL(LT_TRAV_NEAREST, LT_SEM_ITERATIVE); \
    if (trav == LT_TRAV_LBVH && sem == LT_SEM_REFERENCE) return CALL(LT_TRAV_LBVH, LT_SEM_REFERENCE);       \
    if (trav == LT_TRAV_LBVH && sem == LT_SEM_ITERATIVE) return CALL(LT_TRAV_LBVH, LT_SEM_ITERATIVE);

template <typename R>
int launch_wavefront(const WalkParams<R>& P, unsigned long long* gc, int grid, size_t smem, cudaStream_t stream) {
    const int trav = P.traversal, sem = P.semantics;
#define LT_WF_L(T, S) wf_launch_one<R, T, S>(P, gc, grid, smem, stream)
    LT_WF_DISPATCH(LT_WF_L)
#undef LT_WF_L
    set_error("unknown wavefront kernel variant");
    return LT_E_INVALID;
}
template <typename R>
int wavefront_kernel_attrs(int trav, int sem, int* regs, int* max_blocks, size_t smem) {
#define LT_WF_A(T, S) wf_attrs_one<R, T, S>(regs, max_blocks, smem)
    LT_WF_DISPATCH(LT_WF_A)
#undef LT_WF_A
    set_error("unknown wavefront kernel variant");
    return LT_E_INVALID;
}

template <typename R>
int launch_record(const WalkParams<R>& P, const RecParams& RP, int grid, cudaStream_t stream) {
    switch (P.traversal) {
        case LT_TRAV_COMPAT: record_kernel<R, LT_TRAV_COMPAT><<<grid, LT_BLOCK, 0, stream>>>(P, RP); break;
        case LT_TRAV_NEAREST: record_kernel<R, LT_TRAV_NEAREST><<<grid, LT_BLOCK, 0, stream>>>(P, RP); break;
        case LT_TRAV_LBVH: record_kernel<R, LT_TRAV_LBVH><<<grid, LT_BLOCK, 0, stream>>>(P, RP); break;
        default: set_error("unknown traversal %d", P.traversal); return LT_E_INVALID;
    }
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}

template <typename R>
int launch_intersect(const IntersectParams<R>& P, int grid, cudaStream_t stream) {
    switch (P.traversal) {
        case LT_TRAV_COMPAT: intersect_kernel<R, LT_TRAV_COMPAT><<<grid, LT_BLOCK, 32 + (size_t)((((P.S.n_cleaves + 1) >> 1) * LT_SLAB_PAIR_FLOATS * 4 + 16 + 15) / 16 * 16) + (sizeof(R) == 4 ? LT_CHAIN_BYTES : 0), stream>>>(P); break;
        case LT_TRAV_NEAREST: intersect_kernel<R, LT_TRAV_NEAREST><<<grid, LT_BLOCK, 0, stream>>>(P); break;
        case LT_TRAV_LBVH: intersect_kernel<R, LT_TRAV_LBVH><<<grid, LT_BLOCK, 0, stream>>>(P); break;
        case LT_TRAV_LBVH2: intersect_kernel<R, LT_TRAV_LBVH2><<<grid, LT_BLOCK, 0, stream>>>(P); break;
        default: set_error("unknown traversal %d", P.traversal); return LT_E_INVALID;
    }
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}

}  // namespace lt
