// XLA FFI handlers over the C ABI of include/waveflow_b200.h -- the custom calls a JAX build of the reference would register
// (environment.yml pins jax==0.4.30).  OPTIONAL and NOT part of libwaveflow_b200.so: it needs jaxlib's headers, which this
// image does not have, so `python -m waveflow_b200.build` never compiles it.  Build where jaxlib is installed:
//
//   JAXLIB_INC=$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))")
//   g++ -O2 -fPIC -shared -std=c++17 -I$JAXLIB_INC -I../../../include -I/usr/local/cuda/include \
//       xla_ffi_shim.cc -L../.. -lwaveflow_b200 -Wl,-rpath,'$ORIGIN' -o ../../libwaveflow_b200_xla.so
//
// Every handler forwards device pointers, shapes and the stream of the FFI call frame to one wf_* entry point; the POD model
// description (wf_live_model) travels as scalar attributes, the table pointers (wf_live_tables) as operands.
// Reference functions replaced: see the table in DESIGN.md section 1 / the comments in include/waveflow_b200.h.
#include <cstdint>
#include <cstring>

#include "waveflow_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using F64 = ffi::Buffer<ffi::F64>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;

static ffi::Error Status(int st) {
  return st == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, wf_status_string(st));
}

// ---- apply_fun_vec + apply_fun_vec_grad + log(. + 1e-7) of ISpline_fun / MSpline_fun (isplines_jax.py:139-149, made.py:79)
static ffi::Error SplineApplyLocal(cudaStream_t stream, F32 rec, S32 lo, F32 dense, F32 c, F32 x, RF32 val, RF32 grad, RF32 logd,
                                   int32_t kind) {
  const int64_t M = c.dimensions()[0];
  const int P = (int)c.dimensions()[1], T = (int)lo.dimensions()[0];
  return Status(wf_spline_apply_local(rec.typed_data(), lo.typed_data(), dense.typed_data(), kind, T, P, c.typed_data(), x.typed_data(), M,
                                      val->typed_data(), grad->typed_data(), logd->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(kWfSplineApplyLocal, SplineApplyLocal,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Attr<int32_t>("kind"));

// ---- unconstrained_RQS (flows/bijections/neural_splines.py:16-71); flags: WF_RQS_INVERSE | WF_RQS_EXACT_BINS
static ffi::Error RqsApply(cudaStream_t stream, F32 inputs, F32 uw, F32 uh, F32 ud, RF32 out, RF32 logabsdet, RS32 bins, float tail_bound,
                           int32_t flags) {
  const int K = (int)uw.dimensions().back();
  const int64_t M = (int64_t)inputs.element_count();
  return Status(wf_rqs_apply(inputs.typed_data(), uw.typed_data(), uh.typed_data(), ud.typed_data(), M, K, tail_bound, flags,
                             out->typed_data(), logabsdet->typed_data(), bins->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(kWfRqsApply, RqsApply,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<S32>().Attr<float>("tail_bound").Attr<int32_t>("flags"));

// ---- the fused live path.  The model description arrives as attributes (same fields as struct wf_live_model).
struct ModelAttrs {
  int32_t D, n_layers, T, P_I, k_I, prior_kind, P_P, k_P, has_box, coord_mean, bc_I, bc_P, n_knots_P, weight_layout;
  float box, reg, tol;
};
static wf_live_model ToModel(const ModelAttrs& a) {
  wf_live_model m;
  std::memset(&m, 0, sizeof(m));
  m.D = a.D; m.n_layers = a.n_layers; m.T = a.T; m.P_I = a.P_I; m.k_I = a.k_I; m.prior_kind = a.prior_kind; m.P_P = a.P_P; m.k_P = a.k_P;
  m.has_box = a.has_box; m.coord_mean = a.coord_mean; m.bc_I = a.bc_I; m.bc_P = a.bc_P; m.box = a.box; m.reg = a.reg; m.tol = a.tol;
  m.n_knots_P = a.n_knots_P; m.weight_layout = a.weight_layout;
  return m;
}
XLA_FFI_REGISTER_STRUCT_ATTR_DECODING(ModelAttrs, ffi::StructMember<int32_t>("D"), ffi::StructMember<int32_t>("n_layers"),
                                      ffi::StructMember<int32_t>("T"), ffi::StructMember<int32_t>("P_I"), ffi::StructMember<int32_t>("k_I"),
                                      ffi::StructMember<int32_t>("prior_kind"), ffi::StructMember<int32_t>("P_P"),
                                      ffi::StructMember<int32_t>("k_P"), ffi::StructMember<int32_t>("has_box"),
                                      ffi::StructMember<int32_t>("coord_mean"), ffi::StructMember<int32_t>("bc_I"),
                                      ffi::StructMember<int32_t>("bc_P"), ffi::StructMember<int32_t>("n_knots_P"),
                                      ffi::StructMember<int32_t>("weight_layout"), ffi::StructMember<float>("box"),
                                      ffi::StructMember<float>("reg"), ffi::StructMember<float>("tol"));

// tables: dense_I, rec_I, lo_I, rec_I_t, dense_P (OB tables for the B prior), ob_to_b -- the Waveflow configuration
static wf_live_tables ToTables(F32 dense_I, F32 rec_I, S32 lo_I, F32 rec_I_t, F32 dense_P, F32 ob_to_b) {
  wf_live_tables t;
  std::memset(&t, 0, sizeof(t));
  t.dense_I = dense_I.typed_data(); t.rec_I = rec_I.typed_data(); t.lo_I = lo_I.typed_data(); t.rec_I_t = rec_I_t.typed_data();
  t.dense_P = dense_P.typed_data(); t.ob_to_b = ob_to_b.typed_data();
  return t;
}

// Waveflow.psi / log_pdf and Serial.direct_fun (wavefunctions.py:33-71, bijections.py:452-460)
static ffi::Error LiveForward(cudaStream_t stream, F32 dense_I, F32 rec_I, S32 lo_I, F32 rec_I_t, F32 dense_P, F32 ob_to_b, F32 weights, F32 x,
                              RF32 u, RF32 logdet, RF32 logpdf, RF32 psi, ModelAttrs attrs) {
  const wf_live_model m = ToModel(attrs);
  const wf_live_tables t = ToTables(dense_I, rec_I, lo_I, rec_I_t, dense_P, ob_to_b);
  return Status(wf_live_forward(&m, &t, weights.typed_data(), x.typed_data(), (int64_t)x.dimensions()[0], u->typed_data(),
                                logdet->typed_data(), logpdf->typed_data(), psi->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(kWfLiveForward, LiveForward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Attr<ModelAttrs>("model"));

// construct_hamiltonian_function + E_loc (utils/physics.py:79-93, vqmc.py:198-200): psi, H psi, E_loc and the block sums
static ffi::Error LocalEnergy(cudaStream_t stream, F32 dense_I, F32 rec_I, S32 lo_I, F32 rec_I_t, F32 dense_P, F32 ob_to_b, F32 weights, F32 x,
                              RF32 psi, RF32 hpsi, RF32 eloc, RF64 sums, ModelAttrs attrs, ffi::Span<const float> protons) {
  const wf_live_model m = ToModel(attrs);
  const wf_live_tables t = ToTables(dense_I, rec_I, lo_I, rec_I_t, dense_P, ob_to_b);
  if (cudaMemsetAsync(sums->typed_data(), 0, 4 * sizeof(double), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  return Status(wf_local_energy(&m, &t, weights.typed_data(), protons.begin(), (int)protons.size(), x.typed_data(),
                                (int64_t)x.dimensions()[0], psi->typed_data(), hpsi->typed_data(), eloc->typed_data(), nullptr, nullptr,
                                sums->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(kWfLocalEnergy, LocalEnergy,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F64>().Attr<ModelAttrs>("model")
                                  .Attr<ffi::Span<const float>>("protons"));

// weights in the reference's parameter layout -> packed -> tensor-core image happens on the Python side (one call per
// parameter set): wf_live_pack_tc is exposed as well so that the whole path stays inside jit.
static ffi::Error LivePackTc(cudaStream_t stream, F32 packed, RF32 image, int32_t D, int32_t n_nets) {
  return Status(wf_live_pack_tc(D, n_nets, packed.typed_data(), image->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(kWfLivePackTc, LivePackTc,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Ret<F32>().Attr<int32_t>("D")
                                  .Attr<int32_t>("n_nets"));
