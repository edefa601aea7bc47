// Instantiation of the fused live-path kernel for D = 3, plain forward variant.
#include "live_kernel.cuh"
namespace wf {
int launch_live_d3_lap0(LiveParams& P, cudaStream_t s) { return launch_live<3, false>(P, s); }
}  // namespace wf
