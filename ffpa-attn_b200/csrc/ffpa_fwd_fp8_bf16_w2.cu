// explicit instantiations of the FP8 forward (bf16 output, 2 softmax warpgroups): one TU per variant so nvcc runs in parallel
#include "ffpa_fwd_fp8_sm100.cuh"
namespace ffpa {
namespace fp8 {
template int launch_fp8_variant<1, true, 2>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<2, true, 2>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<3, true, 2>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<4, true, 2>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
}  // namespace fp8
}  // namespace ffpa
