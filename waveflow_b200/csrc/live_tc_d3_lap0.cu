// Instantiation of the tensor-core live-path kernel (live_tc.cuh) for D = 3, forward-only variant.
#include "live_tc.cuh"
namespace wf {
int launch_live_tc_d3_lap0(LiveParams& P, const ltc::TcExtra& X, cudaStream_t s) { return ltc::launch_live_tc<3, false>(P, X, s); }
}  // namespace wf
