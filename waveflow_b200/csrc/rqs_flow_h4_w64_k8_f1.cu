// Fused coupling-flow kernel, D = 8, hidden width 64, KP = 8, FULLK = 1.
#include "rqs_flow.cuh"
namespace wf {
namespace cf {
WF_DEF_CF(4, 64, 8, 1)
}  // namespace cf
}  // namespace wf
