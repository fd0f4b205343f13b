// bf16 instantiations of the backward kernel family
#include "ffpa_bwd_sm100.cuh"
namespace ffpa { namespace bwd {
template int dispatch_bwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                      const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const BwdKernelParams&, int, cudaStream_t);
template int launch_preprocess<true>(const ffpa_bwd_params&, float*, float*, int, cudaStream_t);
}}
