// Instantiation of the live-flow inverse / sampler kernel for D = 2.
#include "live_inverse.cuh"
namespace wf {
int launch_inverse_d2(InvParams& P, cudaStream_t s) { return launch_inverse<2>(P, s); }
}  // namespace wf
