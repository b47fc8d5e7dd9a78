// float64 instantiation of the walk / intersect kernels (replay-parity path; built with -fmad=false).
#include "lt_walk_inst.cuh"
namespace lt {
template int launch_walk<double>(const WalkParams<double>&, int, size_t, cudaStream_t);
template int launch_intersect<double>(const IntersectParams<double>&, int, cudaStream_t);
template int walk_kernel_attrs<double>(int, int, int, int*, int*, int*, size_t);
template int launch_wavefront<double>(const WalkParams<double>&, unsigned long long*, int, size_t, cudaStream_t);
template int wavefront_kernel_attrs<double>(int, int, int*, int*, size_t);
template size_t wavefront_smem_bytes<double>(size_t);
template int launch_record<double>(const WalkParams<double>&, const RecParams&, int, cudaStream_t);
}
