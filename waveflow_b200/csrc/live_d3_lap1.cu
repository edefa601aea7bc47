// Instantiation of the fused live-path kernel for D = 3, forward-Laplacian (local energy) variant.
#include "live_kernel.cuh"
namespace wf {
int launch_live_d3_lap1(LiveParams& P, cudaStream_t s) { return launch_live<3, true>(P, s); }
}  // namespace wf
