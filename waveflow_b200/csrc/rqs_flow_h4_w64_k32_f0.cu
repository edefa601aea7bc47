// Fused coupling-flow kernel, D = 8, hidden width 64, KP = 32, FULLK = 0.
#include "rqs_flow.cuh"
namespace wf {
namespace cf {
WF_DEF_CF(4, 64, 32, 0)
}  // namespace cf
}  // namespace wf
