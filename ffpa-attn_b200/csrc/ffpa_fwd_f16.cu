// fp16 instantiations of the forward kernel family
#include "ffpa_fwd_sm100.cuh"
namespace ffpa {
template int dispatch_fwd_dtype<false>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                       const FwdKernelParams&, int, cudaStream_t);
template int launch_merge_splits<false>(const float*, const float*, void*, float*, int64_t, const int64_t*, int, int, int, int, int, cudaStream_t);
}
