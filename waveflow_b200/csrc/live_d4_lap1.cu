// Instantiation of the fused live-path kernel for D = 4, forward-Laplacian (local energy) variant.
#include "live_kernel.cuh"
namespace wf {
int launch_live_d4_lap1(LiveParams& P, cudaStream_t s) { return launch_live<4, true>(P, s); }
}  // namespace wf
