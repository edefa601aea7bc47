// Instantiation of the fused live-path kernel for D = 2, plain forward variant.
#include "live_kernel.cuh"
namespace wf {
int launch_live_d2_lap0(LiveParams& P, cudaStream_t s) { return launch_live<2, false>(P, s); }
}  // namespace wf
