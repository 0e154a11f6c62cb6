// Column-marching loss kernels, block-size class 128 (see col_kernel_impl.cuh).
#include "col_kernel_impl.cuh"

namespace usl {
template int col_launch_class<128>(const ColPlan*, int, bool, int, cudaStream_t);
}
