// lt_walk_inst.cuh — host dispatch over the (traversal, semantics, visited-words) kernel variants;
// included once per arithmetic type by lt_walk_f32.cu / lt_walk_f64.cu so the two compile in parallel
// (and so the float64 unit can be built with -fmad=false: its products and sums must round like the
// reference's separate NumPy operations, see DESIGN.md "numerics").
#pragma once
#include "lt_megakernel.cuh"

namespace lt {

template <typename R, int TRAV, int SEM, int NW>
static int launch_one(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    auto k = walk_megakernel<R, TRAV, SEM, NW>;
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, LT_BLOCK, smem, stream>>>(P);
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}

template <typename R, int TRAV, int NW>
static int launch_sem(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream) {
    if (P.semantics == LT_SEM_REFERENCE) return launch_one<R, TRAV, LT_SEM_REFERENCE, NW>(P, grid, smem, stream);
    return launch_one<R, TRAV, LT_SEM_ITERATIVE, NW>(P, grid, smem, stream);
}

template <typename R>
int launch_walk(const WalkParams<R>& P, int nw, int grid, size_t smem, cudaStream_t stream) {
    switch (P.traversal) {
        case LT_TRAV_COMPAT:
            return nw <= 1 ? launch_sem<R, LT_TRAV_COMPAT, 1>(P, grid, smem, stream) : launch_sem<R, LT_TRAV_COMPAT, 2>(P, grid, smem, stream);
        case LT_TRAV_NEAREST: return launch_sem<R, LT_TRAV_NEAREST, 1>(P, grid, smem, stream);
        case LT_TRAV_LBVH: return launch_sem<R, LT_TRAV_LBVH, 1>(P, grid, smem, stream);
    }
    set_error("unknown traversal %d", P.traversal);
    return LT_E_INVALID;
}

template <typename R, int TRAV, int SEM, int NW>
static int attrs_one(int* regs, int* local_bytes, int* max_blocks, size_t smem) {
    auto k = walk_megakernel<R, TRAV, SEM, NW>;
    cudaFuncAttributes a;
    LT_CUDA_OK(cudaFuncGetAttributes(&a, k));
    if (smem > 48 * 1024) LT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *regs = a.numRegs;
    *local_bytes = (int)a.localSizeBytes;
    LT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks, k, LT_BLOCK, smem));
    return LT_OK;
}

template <typename R>
int walk_kernel_attrs(int trav, int sem, int nw, int* regs, int* local_bytes, int* max_blocks, size_t smem) {
#define LT_ATTR(T, S, N) \
    if (trav == T && sem == S && (T != LT_TRAV_COMPAT || (nw <= 1) == (N == 1))) return attrs_one<R, T, S, N>(regs, local_bytes, max_blocks, smem);
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_REFERENCE, 1)
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_REFERENCE, 2)
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_ITERATIVE, 1)
    LT_ATTR(LT_TRAV_COMPAT, LT_SEM_ITERATIVE, 2)
    LT_ATTR(LT_TRAV_NEAREST, LT_SEM_REFERENCE, 1)
    LT_ATTR(LT_TRAV_NEAREST, LT_SEM_ITERATIVE, 1)
    LT_ATTR(LT_TRAV_LBVH, LT_SEM_REFERENCE, 1)
    LT_ATTR(LT_TRAV_LBVH, LT_SEM_ITERATIVE, 1)
#undef LT_ATTR
    set_error("unknown kernel variant");
    return LT_E_INVALID;
}

template <typename R>
int launch_intersect(const IntersectParams<R>& P, int nw, int grid, cudaStream_t stream) {
    switch (P.traversal) {
        case LT_TRAV_COMPAT:
            if (nw <= 1) intersect_kernel<R, LT_TRAV_COMPAT, 1><<<grid, LT_BLOCK, 0, stream>>>(P);
            else intersect_kernel<R, LT_TRAV_COMPAT, 2><<<grid, LT_BLOCK, 0, stream>>>(P);
            break;
        case LT_TRAV_NEAREST: intersect_kernel<R, LT_TRAV_NEAREST, 1><<<grid, LT_BLOCK, 0, stream>>>(P); break;
        case LT_TRAV_LBVH: intersect_kernel<R, LT_TRAV_LBVH, 1><<<grid, LT_BLOCK, 0, stream>>>(P); break;
        default: set_error("unknown traversal %d", P.traversal); return LT_E_INVALID;
    }
    LT_CUDA_OK(cudaGetLastError());
    return LT_OK;
}

}  // namespace lt
