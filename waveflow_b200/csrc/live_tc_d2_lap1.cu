// Instantiation of the tensor-core live-path kernel (live_tc.cuh) for D = 2, forward-Laplacian (local energy) variant.
#include "live_tc.cuh"
namespace wf {
int launch_live_tc_d2_lap1(LiveParams& P, const ltc::TcExtra& X, cudaStream_t s) { return ltc::launch_live_tc<2, true>(P, X, s); }
}  // namespace wf
