// The 32-column-chunk half of the ag_gemm_fwd kernel table (see gemm_conv_kernel.cuh); compiled as its own translation
// unit so that the 128 instantiations of the kernel template build in parallel.
#include "gemm_conv_kernel.cuh"

namespace agb {

static const MakeGemmTable<32, F_COUNT>::type kGemmTableWide;

GemmKernelFn gemm_wide_kernel(int cg, int flags) { return kGemmTableWide.fn[cg - 1][flags]; }

}  // namespace agb
