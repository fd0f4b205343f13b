// instantiations of the second-slab forward GEMM (replay path, head dims > 768)
#include "ffpa_fwd_replay_sm100.cuh"
namespace ffpa { namespace replay {
template int launch_fwd_replay<true>(const CUtensorMap&, const CUtensorMap&, const FwdReplayParams&, int, cudaStream_t);
template int launch_fwd_replay<false>(const CUtensorMap&, const CUtensorMap&, const FwdReplayParams&, int, cudaStream_t);
}}
