// Instantiation of the live-flow inverse / sampler kernel for D = 3.
#include "live_inverse.cuh"
namespace wf {
int launch_inverse_d3(InvParams& P, cudaStream_t s) { return launch_inverse<3>(P, s); }
}  // namespace wf
