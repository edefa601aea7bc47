// Instantiation of the fused live-path kernel for D = 4, plain forward variant.
#include "live_kernel.cuh"
namespace wf {
int launch_live_d4_lap0(LiveParams& P, cudaStream_t s) { return launch_live<4, false>(P, s); }
}  // namespace wf
