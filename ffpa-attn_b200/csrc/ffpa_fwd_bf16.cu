// bf16 instantiations of the forward kernel family (split per dtype to parallelise the build)
#include "ffpa_fwd_sm100.cuh"
namespace ffpa {
template int dispatch_fwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                      const FwdKernelParams&, int, cudaStream_t);
template int launch_merge_splits<true>(const float*, const float*, void*, float*, int64_t, const int64_t*, int, int, int, int, int, cudaStream_t);
}
