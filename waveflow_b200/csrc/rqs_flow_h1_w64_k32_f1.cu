// Fused coupling-flow kernel, D = 2, hidden width 64, KP = 32, FULLK = 1.
#include "rqs_flow.cuh"
namespace wf {
namespace cf {
WF_DEF_CF(1, 64, 32, 1)
}  // namespace cf
}  // namespace wf
