// Column-marching loss kernels, block-size class 256 (see col_kernel_impl.cuh).
#include "col_kernel_impl.cuh"

namespace usl {
template int col_launch_class<256>(const ColPlan*, int, bool, int, cudaStream_t);
}
