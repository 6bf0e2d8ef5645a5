// Kernels of the bf16 tensor-core GEMM with 192-column tiles (one tile width per file: the files
// compile in parallel).
#include "gemm_tc_kernel.cuh"

namespace npt {
template int launch_tc_bn<192>(const TcParams&, const GemmEpilogue&, float*, const CUtensorMap&, const CUtensorMap&,
                               const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, bool, bool, int, int, cudaStream_t);
}  // namespace npt
