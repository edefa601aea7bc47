// Instantiation of the tensor-core live-path kernel (live_tc.cuh) for D = 4, forward-only variant.
#include "live_tc.cuh"
namespace wf {
int launch_live_tc_d4_lap0(LiveParams& P, const ltc::TcExtra& X, cudaStream_t s) { return ltc::launch_live_tc<4, false>(P, X, s); }
}  // namespace wf
