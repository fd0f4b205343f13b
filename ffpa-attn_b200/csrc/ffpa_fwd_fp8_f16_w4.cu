// explicit instantiations of the FP8 forward (f16 output, 4 softmax warpgroups): one TU per variant so nvcc runs in parallel
#include "ffpa_fwd_fp8_sm100.cuh"
namespace ffpa {
namespace fp8 {
template int launch_fp8_variant<1, false, 4>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<2, false, 4>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<3, false, 4>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
template int launch_fp8_variant<4, false, 4>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
}  // namespace fp8
}  // namespace ffpa
