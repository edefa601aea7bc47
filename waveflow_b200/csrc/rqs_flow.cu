// C-ABI entry points of the fused coupling flow (kernels: rqs_flow.cuh, instantiated in rqs_flow_h*_w*.cu).
#include "rqs_flow.cuh"

using namespace wf;
using namespace wf::cf;

extern "C" int64_t wf_rqs_coupling_net_floats(int D, int K, int Hd) {
  if (D < 2 || (D & 1) || K < 2 || Hd < 1) return -1;
  const int KP = K <= 8 ? 8 : 32;
  return (int64_t)cf_head_floats(D / 2, Hd) + (int64_t)(D / 2) * cf_block_floats(Hd, KP);
}

extern "C" int wf_rqs_coupling_flow(const float* weights, int n_layers, int D, int K, int Hd, float tail_bound, int inverse,
                                    const float* x, int64_t N, float* y, float* logdet, void* stream) {
  if (N == 0) return WF_OK;
  if (!weights || !x || !y || !logdet || N < 0 || n_layers < 1 || !(tail_bound > 0.f) || K < 2) return WF_ERR_INVALID_ARG;
  if ((D & 1) || D < 2 || D > 8 || K > 32 || (Hd != 8 && Hd != 64)) return WF_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(weights) & 15) return WF_ERR_INVALID_ARG;
  CfParams P{weights, x, y, logdet, N, n_layers, K, inverse ? 1 : 0, tail_bound};
  const int KP = K <= 8 ? 8 : 32;
  cudaStream_t s = (cudaStream_t)stream;
  const int full = K == KP ? 1 : 0;
#define WF_CALL_CF(H, W) \
  (KP == 8 ? (full ? launch_cf_h##H##_w##W##_k8_f1(P, s) : launch_cf_h##H##_w##W##_k8_f0(P, s)) \
           : (full ? launch_cf_h##H##_w##W##_k32_f1(P, s) : launch_cf_h##H##_w##W##_k32_f0(P, s)))
  switch (D / 2) {
    case 1: return Hd == 8 ? WF_CALL_CF(1, 8) : WF_CALL_CF(1, 64);
    case 2: return Hd == 8 ? WF_CALL_CF(2, 8) : WF_CALL_CF(2, 64);
    case 4: return Hd == 8 ? WF_CALL_CF(4, 8) : WF_CALL_CF(4, 64);
    default: return WF_ERR_UNSUPPORTED;
  }
}
