// Instantiation of the live-flow inverse / sampler kernel for D = 4.
#include "live_inverse.cuh"
namespace wf {
int launch_inverse_d4(InvParams& P, cudaStream_t s) { return launch_inverse<4>(P, s); }
}  // namespace wf
