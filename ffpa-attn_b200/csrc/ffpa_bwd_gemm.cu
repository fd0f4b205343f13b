// instantiations of the GEMM-only dK / dV kernel (stash path of the backward)
#include "ffpa_bwd_gemm_sm100.cuh"
namespace ffpa { namespace bwd {
template int launch_bwd_gemm<true>(const CUtensorMap&, const CUtensorMap&, const BwdGemmParams&, int, cudaStream_t);
template int launch_bwd_gemm<false>(const CUtensorMap&, const CUtensorMap&, const BwdGemmParams&, int, cudaStream_t);
}}
