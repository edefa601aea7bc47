// Fused coupling-flow kernel, D = 2, hidden width 64, KP = 8, FULLK = 0.
#include "rqs_flow.cuh"
namespace wf {
namespace cf {
WF_DEF_CF(1, 64, 8, 0)
}  // namespace cf
}  // namespace wf
