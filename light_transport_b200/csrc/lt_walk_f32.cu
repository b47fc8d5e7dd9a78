// float32 instantiation of the walk / intersect kernels (throughput path).
#include "lt_walk_inst.cuh"
namespace lt {
template int launch_walk<float>(const WalkParams<float>&, int, size_t, cudaStream_t);
template int launch_intersect<float>(const IntersectParams<float>&, int, cudaStream_t);
template int walk_kernel_attrs<float>(int, int, int, int*, int*, int*, size_t);
template int launch_wavefront<float>(const WalkParams<float>&, unsigned long long*, int, size_t, cudaStream_t);
template int wavefront_kernel_attrs<float>(int, int, int*, int*, size_t);
template size_t wavefront_smem_bytes<float>(size_t);
template int launch_record<float>(const WalkParams<float>&, const RecParams&, int, cudaStream_t);
}
