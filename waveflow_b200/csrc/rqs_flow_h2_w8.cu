// Fused coupling-flow kernels, D = 4, hidden width 8.
#include "rqs_flow.cuh"
namespace wf {
namespace cf {
WF_DEF_CF(2, 8, 8, 0) WF_DEF_CF(2, 8, 8, 1) WF_DEF_CF(2, 8, 32, 0) WF_DEF_CF(2, 8, 32, 1)
}  // namespace cf
}  // namespace wf
