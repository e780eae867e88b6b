// XLA-FFI binding of the qgrad_b200 C ABI (include/qgrad_b200.h) -- the "thin XLA-FFI layer" the
// north star names, so that a JAX host registers each op as a primitive with a custom_vjp.
//
// One handler per compute entry point of the header and per dtype:
//     qg_<op>(dtype, ...)   ->   qg_xla_<op>_c64 , qg_xla_<op>_c128
// (QG_XLA_DEFINE below; tests/test_cpu_host.py checks, at string level, that the set of <op> here equals the set
// of compute symbols the header declares).  Conventions of every handler:
//   * buffers in the order of the C prototype: inputs are Arg<>, outputs Ret<>, scalars Attr<>;
//   * the expm workspace is an XLA-allocated U8 scratch RESULT (size from qg_expm_workspace_bytes, computed on the
//     Python side), so the library owns nothing; the two-pass pair hands it from qg_xla_expm_fwd_save (result) to
//     qg_xla_expm_bwd_saved (operand, aliased to a result: input_output_aliases={2: 1});
//   * in-place C entry points (Adam) take operands aliased to results; if XLA did not honour the alias the handler
//     copies operand -> result first, then updates the result in place;
//   * batch and dimension come from the buffer shapes ((..., n, n) matrices, (..., n) kets, (..., k) angle rows);
//   * a non-zero status becomes an ffi::Error (negative: kInvalidArgument, positive cudaError_t: kInternal).
//
// STATUS: written to the public XLA FFI C++ API (xla/ffi/api/ffi.h, as shipped by `jax.ffi.include_dir()` in
// JAX >= 0.4.31 / 0.5).  JAX/jaxlib are absent from this environment and not installable, so
// it cannot be compiled against the real header or run here; it IS type-checked on every CPU test run against a
// minimal mock of that header (tests/mock_xla/xla/ffi/api/ffi.h: same names, and a static_assert that every handler
// is invocable with exactly the decoded types of its binding) together with the real include/qgrad_b200.h, so
// argument order and types against the C prototypes are verified.  The file is guarded by __has_include and is not
// part of qgrad_b200/build.py's unit list.  Build, next to a JAX install:
//
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())")
//       -I/usr/local/cuda/include qgrad_b200/csrc/xla_ffi_shim.cc -L qgrad_b200 -lqgrad_b200
//       -o qgrad_b200/libqgrad_b200_xla.so
//
// Python side (register_ffi_target / ffi_call / custom_vjp for every op): INTEGRATION.md section 2.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define QG_HAVE_XLA_FFI 1
#endif
#endif

#ifdef QG_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <utility>

#include "xla/ffi/api/ffi.h"
#include "../../include/qgrad_b200.h"

namespace ffi = xla::ffi;

namespace {

using Stream = ffi::PlatformStream<cudaStream_t>;

template <int D> struct Ty;
template <> struct Ty<QG_C64> { static constexpr ffi::DataType C = ffi::C64; static constexpr ffi::DataType R = ffi::F32; };
template <> struct Ty<QG_C128> { static constexpr ffi::DataType C = ffi::C128; static constexpr ffi::DataType R = ffi::F64; };
template <int D> using CBuf = ffi::Buffer<Ty<D>::C>;       // complex data of the op's dtype
template <int D> using RBuf = ffi::Buffer<Ty<D>::R>;       // the matching real type (angles, losses)
using WsBuf = ffi::Buffer<ffi::U8>;
using I32Buf = ffi::Buffer<ffi::S32>;
using I64Buf = ffi::Buffer<ffi::S64>;
template <typename B> using Out = ffi::Result<B>;

ffi::Error Status(int rc, const char* what) {
  if (rc == QG_OK) return ffi::Error::Success();
  return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + qg_status_string(rc));
}
ffi::Error Bad(const char* what, const char* why) {
  return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(what) + ": " + why);
}

// product of all dimensions but the last `core`
template <typename Buf>
int64_t Lead(const Buf& a, size_t core) {
  auto dims = a.dimensions();
  int64_t batch = 1;
  for (size_t i = 0; i + core < dims.size(); ++i) batch *= dims[i];
  return batch;
}
template <typename Buf> int Last(const Buf& a) { return static_cast<int>(a.dimensions().back()); }
template <typename Buf> size_t Rank(const Buf& a) { return a.dimensions().size(); }

// operand -> result copy for the in-place entry points when XLA did not alias the pair
template <typename In, typename OutT>
cudaError_t Carry(cudaStream_t st, const In& in, OutT& out) {
  if (in.untyped_data() == out->untyped_data()) return cudaSuccess;
  return cudaMemcpyAsync(out->untyped_data(), in.untyped_data(), in.size_bytes(), cudaMemcpyDeviceToDevice, st);
}
#define QG_CARRY(in, out) do { cudaError_t e__ = Carry(st, in, out); if (e__ != cudaSuccess) return Status((int)e__, "operand->result copy"); } while (0)

// ------------------------------------------------------------------------------------------------ expm family
template <int D>
ffi::Error Expm(cudaStream_t st, CBuf<D> a, Out<CBuf<D>> y, Out<WsBuf> ws, int64_t max_squarings) {
  return Status(qg_expm(D, a.untyped_data(), y->untyped_data(), Lead(a, 2), Last(a), ws->untyped_data(), ws->size_bytes(),
                        static_cast<int>(max_squarings), st), "qg_expm");
}
#define BindExpm(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Ret<CBuf<D>>().Ret<WsBuf>().Attr<int64_t>("max_squarings")

// (Y, Abar) fused; adjoint: QG_ADJ_TRANS (1) is JAX's VJP L(A^T, ct) of the holomorphic map, QG_ADJ_CONJ (0) torch's
template <int D>
ffi::Error ExpmFwdVjp(cudaStream_t st, CBuf<D> a, CBuf<D> ct, Out<CBuf<D>> y, Out<CBuf<D>> abar, Out<WsBuf> ws,
                      int64_t max_squarings, int64_t adjoint) {
  return Status(qg_expm_fwd_vjp(D, a.untyped_data(), ct.untyped_data(), y->untyped_data(), abar->untyped_data(), Lead(a, 2),
                                Last(a), static_cast<int>(adjoint), ws->untyped_data(), ws->size_bytes(),
                                static_cast<int>(max_squarings), st), "qg_expm_fwd_vjp");
}
#define BindExpmFwdVjp(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>().Ret<CBuf<D>>().Ret<WsBuf>() \
      .Attr<int64_t>("max_squarings").Attr<int64_t>("adjoint")

// two-pass form: the forward's workspace result is the residual the backward consumes
template <int D>
ffi::Error ExpmFwdSave(cudaStream_t st, CBuf<D> a, Out<CBuf<D>> y, Out<WsBuf> ws, int64_t max_squarings) {
  return Status(qg_expm_fwd_save(D, a.untyped_data(), y->untyped_data(), Lead(a, 2), Last(a), ws->untyped_data(),
                                 ws->size_bytes(), static_cast<int>(max_squarings), st), "qg_expm_fwd_save");
}
template <int D>
ffi::Error ExpmBwdSaved(cudaStream_t st, CBuf<D> a, CBuf<D> ct, WsBuf ws_in, Out<CBuf<D>> abar, Out<WsBuf> ws,
                        int64_t max_squarings, int64_t adjoint) {
  QG_CARRY(ws_in, ws);                                   // the dual recurrences run inside the saved workspace
  return Status(qg_expm_bwd_saved(D, a.untyped_data(), ct.untyped_data(), abar->untyped_data(), Lead(a, 2), Last(a),
                                  static_cast<int>(adjoint), ws->untyped_data(), ws->size_bytes(),
                                  static_cast<int>(max_squarings), st), "qg_expm_bwd_saved");
}
#define BindExpmBwdSaved(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Arg<WsBuf>().Ret<CBuf<D>>().Ret<WsBuf>() \
      .Attr<int64_t>("max_squarings").Attr<int64_t>("adjoint")

// dst (..., cols, rows) = src (..., rows, cols)^H
template <int D>
ffi::Error ConjTranspose(cudaStream_t st, CBuf<D> src, Out<CBuf<D>> dst) {
  auto d = src.dimensions();
  if (d.size() < 2) return Bad("qg_conj_transpose", "rank < 2");
  return Status(qg_conj_transpose(D, src.untyped_data(), dst->untyped_data(), Lead(src, 2), static_cast<int>(d[d.size() - 2]),
                                  static_cast<int>(d.back()), st), "qg_conj_transpose");
}
#define BindUnaryC(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Ret<CBuf<D>>()

// ------------------------------------------------------------------------------------------------ expect / fidelity
// oper (n, n) shared or (batch, n, n); a mismatched operator batch is an error (it would read out of bounds)
template <typename OB>
int OperBatched(const OB& oper, int64_t batch, bool* ok) {
  const int64_t ob = Lead(oper, 2);
  *ok = (ob == 1 || ob == batch);
  return (ob == batch && batch > 1) ? 1 : 0;
}
template <int D>
ffi::Error ExpectKetFwd(cudaStream_t st, CBuf<D> oper, CBuf<D> ket, Out<CBuf<D>> out) {
  const int64_t batch = Lead(ket, 1);
  bool ok; const int ob = OperBatched(oper, batch, &ok);
  if (!ok) return Bad("qg_expect_ket_fwd", "operator batch must be 1 or equal to the state batch");
  return Status(qg_expect_ket_fwd(D, oper.untyped_data(), ket.untyped_data(), out->untyped_data(), batch, Last(ket), ob, st),
                "qg_expect_ket_fwd");
}
template <int D>
ffi::Error ExpectKetBwd(cudaStream_t st, CBuf<D> oper, CBuf<D> ket, CBuf<D> outbar, Out<CBuf<D>> ketbar, Out<CBuf<D>> operbar) {
  const int64_t batch = Lead(ket, 1);
  bool ok; const int ob = OperBatched(oper, batch, &ok);
  if (!ok) return Bad("qg_expect_ket_bwd", "operator batch must be 1 or equal to the state batch");
  return Status(qg_expect_ket_bwd(D, oper.untyped_data(), ket.untyped_data(), outbar.untyped_data(), ketbar->untyped_data(),
                                  operbar->untyped_data(), batch, Last(ket), ob, st), "qg_expect_ket_bwd");
}
template <int D>
ffi::Error ExpectDmFwd(cudaStream_t st, CBuf<D> oper, CBuf<D> rho, Out<CBuf<D>> out) {
  const int64_t batch = Lead(rho, 2);
  bool ok; const int ob = OperBatched(oper, batch, &ok);
  if (!ok) return Bad("qg_expect_dm_fwd", "operator batch must be 1 or equal to the state batch");
  return Status(qg_expect_dm_fwd(D, oper.untyped_data(), rho.untyped_data(), out->untyped_data(), batch, Last(rho), ob, st),
                "qg_expect_dm_fwd");
}
template <int D>
ffi::Error ExpectDmBwd(cudaStream_t st, CBuf<D> oper, CBuf<D> rho, CBuf<D> outbar, Out<CBuf<D>> rhobar, Out<CBuf<D>> operbar) {
  const int64_t batch = Lead(rho, 2);
  bool ok; const int ob = OperBatched(oper, batch, &ok);
  if (!ok) return Bad("qg_expect_dm_bwd", "operator batch must be 1 or equal to the state batch");
  return Status(qg_expect_dm_bwd(D, oper.untyped_data(), rho.untyped_data(), outbar.untyped_data(), rhobar->untyped_data(),
                                 operbar->untyped_data(), batch, Last(rho), ob, st), "qg_expect_dm_bwd");
}
#define BindExpectFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>()
#define BindExpectBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>().Ret<CBuf<D>>()

template <int D>
ffi::Error FidelityKetFwd(cudaStream_t st, CBuf<D> a, CBuf<D> b, Out<RBuf<D>> out) {
  return Status(qg_fidelity_ket_fwd(D, a.untyped_data(), b.untyped_data(), out->untyped_data(), Lead(a, 1), Last(a), st),
                "qg_fidelity_ket_fwd");
}
template <int D>
ffi::Error FidelityKetBwd(cudaStream_t st, CBuf<D> a, CBuf<D> b, RBuf<D> outbar, Out<CBuf<D>> abar, Out<CBuf<D>> bbar) {
  return Status(qg_fidelity_ket_bwd(D, a.untyped_data(), b.untyped_data(), outbar.untyped_data(), abar->untyped_data(),
                                    bbar->untyped_data(), Lead(a, 1), Last(a), st), "qg_fidelity_ket_bwd");
}
template <int D>
ffi::Error FidelityDmFwd(cudaStream_t st, CBuf<D> r1, CBuf<D> r2, Out<RBuf<D>> out) {
  return Status(qg_fidelity_dm_fwd(D, r1.untyped_data(), r2.untyped_data(), out->untyped_data(), Lead(r1, 2), Last(r1), st),
                "qg_fidelity_dm_fwd");
}
template <int D>
ffi::Error FidelityDmBwd(cudaStream_t st, CBuf<D> r1, CBuf<D> r2, RBuf<D> outbar, Out<CBuf<D>> r1bar, Out<CBuf<D>> r2bar) {
  return Status(qg_fidelity_dm_bwd(D, r1.untyped_data(), r2.untyped_data(), outbar.untyped_data(), r1bar->untyped_data(),
                                   r2bar->untyped_data(), Lead(r1, 2), Last(r1), st), "qg_fidelity_dm_bwd");
}
#define BindFidelityFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>()
#define BindFidelityBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Arg<RBuf<D>>().Ret<CBuf<D>>().Ret<CBuf<D>>()

// ------------------------------------------------------------------------------------------------ Displace
// evecs (n, n), evals (n): the real eigensystem Displace.__init__ computes once (jnp.linalg.eigh); alpha (...,) complex
template <int D>
ffi::Error DisplaceFwd(cudaStream_t st, RBuf<D> evecs, RBuf<D> evals, CBuf<D> alpha, Out<CBuf<D>> Dm) {
  return Status(qg_displace_fwd(D, evecs.untyped_data(), evals.untyped_data(), alpha.untyped_data(), Dm->untyped_data(),
                                Lead(alpha, 0), Last(evals), st), "qg_displace_fwd");
}
template <int D>
ffi::Error DisplaceBwd(cudaStream_t st, RBuf<D> evecs, RBuf<D> evals, CBuf<D> alpha, CBuf<D> Dbar, Out<CBuf<D>> alphabar) {
  return Status(qg_displace_bwd(D, evecs.untyped_data(), evals.untyped_data(), alpha.untyped_data(), Dbar.untyped_data(),
                                alphabar->untyped_data(), Lead(alpha, 0), Last(evals), st), "qg_displace_bwd");
}
template <int D>
ffi::Error DisplaceApplyFwd(cudaStream_t st, RBuf<D> evecs, RBuf<D> evals, CBuf<D> alpha, CBuf<D> x, Out<CBuf<D>> y) {
  return Status(qg_displace_apply_fwd(D, evecs.untyped_data(), evals.untyped_data(), alpha.untyped_data(), x.untyped_data(),
                                      y->untyped_data(), Lead(x, 1), Last(x), st), "qg_displace_apply_fwd");
}
template <int D>
ffi::Error DisplaceApplyBwd(cudaStream_t st, RBuf<D> evecs, RBuf<D> evals, CBuf<D> alpha, CBuf<D> x, CBuf<D> ybar,
                            Out<CBuf<D>> alphabar, Out<CBuf<D>> xbar) {
  return Status(qg_displace_apply_bwd(D, evecs.untyped_data(), evals.untyped_data(), alpha.untyped_data(), x.untyped_data(),
                                      ybar.untyped_data(), alphabar->untyped_data(), xbar->untyped_data(), Lead(x, 1), Last(x), st),
                "qg_displace_apply_bwd");
}
#define BindDisplaceFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>()
#define BindDisplaceBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>()
#define BindDisplaceApplyBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>() \
      .Ret<CBuf<D>>().Ret<CBuf<D>>()

// ------------------------------------------------------------------------------------------------ rotations / Unitary
template <int D>
ffi::Error GivensUnitaryFwd(cudaStream_t st, RBuf<D> thetas, RBuf<D> phis, RBuf<D> omegas, Out<CBuf<D>> U) {
  return Status(qg_givens_unitary_fwd(D, thetas.untyped_data(), phis.untyped_data(), omegas.untyped_data(), U->untyped_data(),
                                      Lead(omegas, 1), Last(omegas), st), "qg_givens_unitary_fwd");
}
template <int D>
ffi::Error GivensUnitaryBwd(cudaStream_t st, RBuf<D> thetas, RBuf<D> phis, RBuf<D> omegas, CBuf<D> Ubar, Out<RBuf<D>> tb,
                            Out<RBuf<D>> pb, Out<RBuf<D>> ob) {
  return Status(qg_givens_unitary_bwd(D, thetas.untyped_data(), phis.untyped_data(), omegas.untyped_data(), Ubar.untyped_data(),
                                      tb->untyped_data(), pb->untyped_data(), ob->untyped_data(), Lead(omegas, 1), Last(omegas), st),
                "qg_givens_unitary_bwd");
}
#define BindGivensFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Ret<CBuf<D>>()
#define BindGivensBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>().Ret<RBuf<D>>() \
      .Ret<RBuf<D>>()

// params (..., 2) = (theta, phi); N, i, j static
template <int D>
ffi::Error MakeRotFwd(cudaStream_t st, RBuf<D> params, Out<CBuf<D>> R, int64_t N, int64_t i, int64_t j) {
  return Status(qg_make_rot_fwd(D, params.untyped_data(), R->untyped_data(), Lead(params, 1), static_cast<int>(N),
                                static_cast<int>(i), static_cast<int>(j), st), "qg_make_rot_fwd");
}
template <int D>
ffi::Error MakeRotBwd(cudaStream_t st, RBuf<D> params, CBuf<D> Rbar, Out<RBuf<D>> paramsbar, int64_t N, int64_t i, int64_t j) {
  return Status(qg_make_rot_bwd(D, params.untyped_data(), Rbar.untyped_data(), paramsbar->untyped_data(), Lead(params, 1),
                                static_cast<int>(N), static_cast<int>(i), static_cast<int>(j), st), "qg_make_rot_bwd");
}
#define BindMakeRotFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Ret<CBuf<D>>().Attr<int64_t>("N").Attr<int64_t>("i").Attr<int64_t>("j")
#define BindMakeRotBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>().Attr<int64_t>("N").Attr<int64_t>("i") \
      .Attr<int64_t>("j")

// angles (..., 3) = (phi, theta, omega)
template <int D>
ffi::Error RotZyzFwd(cudaStream_t st, RBuf<D> angles, Out<CBuf<D>> R) {
  return Status(qg_rot_zyz_fwd(D, angles.untyped_data(), R->untyped_data(), Lead(angles, 1), st), "qg_rot_zyz_fwd");
}
template <int D>
ffi::Error RotZyzBwd(cudaStream_t st, RBuf<D> angles, CBuf<D> Rbar, Out<RBuf<D>> anglesbar) {
  return Status(qg_rot_zyz_bwd(D, angles.untyped_data(), Rbar.untyped_data(), anglesbar->untyped_data(), Lead(angles, 1), st),
                "qg_rot_zyz_bwd");
}
#define BindRotZyzFwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Ret<CBuf<D>>()
#define BindRotZyzBwd(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>()

// fused Qubit_Rotation cost: ket (2,) shared or (batch, 2)
template <int D>
ffi::Error RotFidelityFwdGrad(cudaStream_t st, RBuf<D> angles, CBuf<D> ket, Out<RBuf<D>> fid, Out<RBuf<D>> grad) {
  const int64_t batch = Lead(angles, 1), kb = Lead(ket, 1);
  if (kb != 1 && kb != batch) return Bad("qg_rot_fidelity_fwd_grad", "ket batch must be 1 or equal to the angle batch");
  return Status(qg_rot_fidelity_fwd_grad(D, angles.untyped_data(), ket.untyped_data(), (kb == batch && batch > 1) ? 1 : 0,
                                         fid->untyped_data(), grad->untyped_data(), batch, st), "qg_rot_fidelity_fwd_grad");
}
#define BindRotFidelity(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>().Ret<RBuf<D>>()

// fused Unitary-learning cost: kets_in / kets_out (n_kets, N) shared by the batch of parameter sets
template <int D>
ffi::Error UnitaryCostFwdGrad(cudaStream_t st, RBuf<D> thetas, RBuf<D> phis, RBuf<D> omegas, CBuf<D> kets_in, CBuf<D> kets_out,
                              Out<RBuf<D>> loss, Out<RBuf<D>> tb, Out<RBuf<D>> pb, Out<RBuf<D>> ob) {
  return Status(qg_unitary_cost_fwd_grad(D, thetas.untyped_data(), phis.untyped_data(), omegas.untyped_data(),
                                         kets_in.untyped_data(), kets_out.untyped_data(), static_cast<int>(Lead(kets_in, 1)),
                                         loss->untyped_data(), tb->untyped_data(), pb->untyped_data(), ob->untyped_data(),
                                         Lead(omegas, 1), Last(omegas), st), "qg_unitary_cost_fwd_grad");
}
#define BindUnitaryCost(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<RBuf<D>>() \
      .Ret<RBuf<D>>().Ret<RBuf<D>>().Ret<RBuf<D>>()

// fused SNAP cost: alphas (batch, T), thetas (batch, T, n), init / target (n) shared -> cost (batch), alphabar, thetabar
template <int D>
ffi::Error SnapCostFwdGrad(cudaStream_t st, RBuf<D> evecs, RBuf<D> evals, CBuf<D> alphas, RBuf<D> thetas, CBuf<D> init, CBuf<D> target,
                           Out<RBuf<D>> cost, Out<CBuf<D>> alphabar, Out<RBuf<D>> thetabar) {
  if (Last(thetas) != Last(evals)) return Bad("qg_snap_cost_fwd_grad", "theta rows must have the Hilbert-space length (pad_thetas)");
  return Status(qg_snap_cost_fwd_grad(D, evecs.untyped_data(), evals.untyped_data(), alphas.untyped_data(), thetas.untyped_data(),
                                      init.untyped_data(), target.untyped_data(), Last(alphas), cost->untyped_data(),
                                      alphabar->untyped_data(), thetabar->untyped_data(), Lead(alphas, 1), Last(evals), st),
                "qg_snap_cost_fwd_grad");
}
#define BindSnapCost(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Arg<RBuf<D>>().Arg<CBuf<D>>().Arg<CBuf<D>>() \
      .Ret<RBuf<D>>().Ret<CBuf<D>>().Ret<RBuf<D>>()

// ------------------------------------------------------------------------------------------------ learning step
// theta (P,), ctrl (n_set, P), xmask / zmask (P,) int32  ->  A (n_set, 2^nq, 2^nq)
template <int D>
ffi::Error PauliHamiltonian(cudaStream_t st, RBuf<D> theta, RBuf<D> ctrl, I32Buf xmask, I32Buf zmask, Out<CBuf<D>> A,
                            int64_t n_qubits, double t) {
  return Status(qg_pauli_hamiltonian(D, theta.untyped_data(), ctrl.untyped_data(), static_cast<const int32_t*>(xmask.untyped_data()),
                                     static_cast<const int32_t*>(zmask.untyped_data()), Last(ctrl), static_cast<int>(Lead(ctrl, 1)),
                                     static_cast<int>(n_qubits), t, A->untyped_data(), st), "qg_pauli_hamiltonian");
}
template <int D>
ffi::Error PauliProject(cudaStream_t st, CBuf<D> Abar, RBuf<D> ctrl, I32Buf xmask, I32Buf zmask, Out<RBuf<D>> thetabar,
                        int64_t n_qubits, double t) {
  return Status(qg_pauli_project(D, Abar.untyped_data(), ctrl.untyped_data(), static_cast<const int32_t*>(xmask.untyped_data()),
                                 static_cast<const int32_t*>(zmask.untyped_data()), Last(ctrl), static_cast<int>(Lead(ctrl, 1)),
                                 static_cast<int>(n_qubits), t, thetabar->untyped_data(), /*accumulate=*/0, st), "qg_pauli_project");
}
#define BindPauliHamiltonian(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<I32Buf>().Arg<I32Buf>().Ret<CBuf<D>>() \
      .Attr<int64_t>("n_qubits").Attr<double>("t")
#define BindPauliProject(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<RBuf<D>>().Arg<I32Buf>().Arg<I32Buf>().Ret<RBuf<D>>() \
      .Attr<int64_t>("n_qubits").Attr<double>("t")

// C (..., m, n) = A (..., m, k) . B (..., k, n), contiguous
template <int D>
ffi::Error Gemm(cudaStream_t st, CBuf<D> A, CBuf<D> B, Out<CBuf<D>> Cm) {
  auto da = A.dimensions(); auto db = B.dimensions();
  if (da.size() < 2 || db.size() < 2) return Bad("qg_gemm", "rank < 2");
  const int64_t m = da[da.size() - 2], k = da.back(), n = db.back();
  if (db[db.size() - 2] != k || Lead(A, 2) != Lead(B, 2)) return Bad("qg_gemm", "shape mismatch");
  return Status(qg_gemm(D, A.untyped_data(), B.untyped_data(), Cm->untyped_data(), Lead(A, 2), static_cast<int>(m),
                        static_cast<int>(n), static_cast<int>(k), k, n, n, m * k, k * n, m * n, st), "qg_gemm");
}
#define BindGemm(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>()

// phi, out (n_set, n, m_local) with kets as columns -> G same shape, loss_partial () = sum_k |Re o_k|
template <int D>
ffi::Error OverlapLoss(cudaStream_t st, CBuf<D> phi, CBuf<D> out, Out<CBuf<D>> G, Out<RBuf<D>> loss_partial, double inv_m_total) {
  auto d = phi.dimensions();
  if (d.size() < 2) return Bad("qg_overlap_loss", "rank < 2");
  cudaError_t e = cudaMemsetAsync(loss_partial->untyped_data(), 0, loss_partial->size_bytes(), st);   // the C entry accumulates
  if (e != cudaSuccess) return Status((int)e, "qg_overlap_loss");
  return Status(qg_overlap_loss(D, phi.untyped_data(), out.untyped_data(), G->untyped_data(), loss_partial->untyped_data(),
                                Lead(phi, 2), static_cast<int>(d[d.size() - 2]), static_cast<int>(d.back()), inv_m_total, st),
                "qg_overlap_loss");
}
#define BindOverlapLoss(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Arg<CBuf<D>>().Ret<CBuf<D>>().Ret<RBuf<D>>().Attr<double>("inv_m_total")

// Adam: (x, m, v) operands aliased to the three results (input_output_aliases={0: 0, 1: 1, 2: 2})
template <int D>
ffi::Error AdamStep(cudaStream_t st, RBuf<D> x, RBuf<D> m, RBuf<D> v, RBuf<D> grad, Out<RBuf<D>> x_out, Out<RBuf<D>> m_out,
                    Out<RBuf<D>> v_out, double lr, double b1, double b2, double eps, int64_t i) {
  QG_CARRY(x, x_out); QG_CARRY(m, m_out); QG_CARRY(v, v_out);
  return Status(qg_adam_step(D, x_out->untyped_data(), m_out->untyped_data(), v_out->untyped_data(), grad.untyped_data(),
                             static_cast<int64_t>(x.element_count()), lr, b1, b2, eps, i, st), "qg_adam_step");
}
#define BindAdamStep(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Ret<RBuf<D>>().Ret<RBuf<D>>() \
      .Ret<RBuf<D>>().Attr<double>("lr").Attr<double>("b1").Attr<double>("b2").Attr<double>("eps").Attr<int64_t>("i")
// the same with the step index in device memory (aliased operand/result) and the loss written next to it
template <int D>
ffi::Error AdamStepDevice(cudaStream_t st, RBuf<D> x, RBuf<D> m, RBuf<D> v, RBuf<D> grad, I64Buf step, RBuf<D> loss_partial,
                          Out<RBuf<D>> x_out, Out<RBuf<D>> m_out, Out<RBuf<D>> v_out, Out<I64Buf> step_out, Out<RBuf<D>> loss_out,
                          double lr, double b1, double b2, double eps, double inv_m_total) {
  QG_CARRY(x, x_out); QG_CARRY(m, m_out); QG_CARRY(v, v_out); QG_CARRY(step, step_out);
  return Status(qg_adam_step_device(D, x_out->untyped_data(), m_out->untyped_data(), v_out->untyped_data(), grad.untyped_data(),
                                    static_cast<int64_t>(x.element_count()), lr, b1, b2, eps,
                                    static_cast<int64_t*>(step_out->untyped_data()), loss_partial.untyped_data(), inv_m_total,
                                    loss_out->untyped_data(), st), "qg_adam_step_device");
}
#define BindAdamStepDevice(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<RBuf<D>>().Arg<I64Buf>().Arg<RBuf<D>>() \
      .Ret<RBuf<D>>().Ret<RBuf<D>>().Ret<RBuf<D>>().Ret<I64Buf>().Ret<RBuf<D>>() \
      .Attr<double>("lr").Attr<double>("b1").Attr<double>("b2").Attr<double>("eps").Attr<double>("inv_m_total")

// ------------------------------------------------------------------------------------------------ Hermitian eigen-problems
template <int D>
ffi::Error Eigh(cudaStream_t st, CBuf<D> A, Out<RBuf<D>> evals, Out<CBuf<D>> evecs) {
  return Status(qg_eigh(D, A.untyped_data(), evals->untyped_data(), evecs->untyped_data(), Lead(A, 2), Last(A), st), "qg_eigh");
}
#define BindEigh(D) ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Ret<RBuf<D>>().Ret<CBuf<D>>()
template <int D>
ffi::Error SqrtmPsd(cudaStream_t st, CBuf<D> A, Out<CBuf<D>> out) {
  return Status(qg_sqrtm_psd(D, A.untyped_data(), out->untyped_data(), Lead(A, 2), Last(A), st), "qg_sqrtm_psd");
}
template <int D>
ffi::Error FidelityUhlmann(cudaStream_t st, CBuf<D> rho, CBuf<D> sigma, Out<RBuf<D>> out) {
  return Status(qg_fidelity_uhlmann(D, rho.untyped_data(), sigma.untyped_data(), out->untyped_data(), Lead(rho, 2), Last(rho), st),
                "qg_fidelity_uhlmann");
}
template <int D>
ffi::Error HermStats(cudaStream_t st, CBuf<D> A, Out<RBuf<D>> stats) {
  return Status(qg_herm_stats(D, A.untyped_data(), stats->untyped_data(), Lead(A, 2), Last(A), st), "qg_herm_stats");
}
#define BindHermStats(D) ffi::Ffi::Bind().Ctx<Stream>().Arg<CBuf<D>>().Ret<RBuf<D>>()

// ------------------------------------------------------------------------------------------------ generators
// outputs only; seed / first / kind / scale / range are attributes
template <int D>
ffi::Error RandKet(cudaStream_t st, Out<CBuf<D>> out, int64_t seed, int64_t first_ket, int64_t kind) {
  return Status(qg_rand_ket(D, out->untyped_data(), Lead(*out, 1), Last(*out), static_cast<uint64_t>(seed),
                            static_cast<uint64_t>(first_ket), static_cast<int>(kind), st), "qg_rand_ket");
}
#define BindRandKet(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Ret<CBuf<D>>().Attr<int64_t>("seed").Attr<int64_t>("first_ket").Attr<int64_t>("kind")
template <int D>
ffi::Error RandHermitian(cudaStream_t st, Out<CBuf<D>> H, int64_t seed, int64_t first_matrix, double scale) {
  return Status(qg_rand_hermitian(D, H->untyped_data(), Lead(*H, 2), Last(*H), static_cast<uint64_t>(seed),
                                  static_cast<uint64_t>(first_matrix), scale, st), "qg_rand_hermitian");
}
#define BindRandHermitian(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Ret<CBuf<D>>().Attr<int64_t>("seed").Attr<int64_t>("first_matrix").Attr<double>("scale")
template <int D>
ffi::Error RandUniform(cudaStream_t st, Out<RBuf<D>> out, int64_t seed, double lo, double hi) {
  return Status(qg_rand_uniform(D, out->untyped_data(), static_cast<int64_t>(out->element_count()), static_cast<uint64_t>(seed),
                                lo, hi, st), "qg_rand_uniform");
}
#define BindRandUniform(D) \
  ffi::Ffi::Bind().Ctx<Stream>().Ret<RBuf<D>>().Attr<int64_t>("seed").Attr<double>("lo").Attr<double>("hi")

}  // namespace

// qg_<op>  ->  qg_xla_<op>_c64, qg_xla_<op>_c128   (the Bind* macros above spell each binding for a CONCRETE dtype, so the
// fluent Ctx/Arg/Ret/Attr chain never sits in a dependent context)
#define QG_XLA_DEFINE(op, Impl, Binding)                                               \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_##op##_c64, (Impl<QG_C64>), Binding(QG_C64));   \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_##op##_c128, (Impl<QG_C128>), Binding(QG_C128))

QG_XLA_DEFINE(expm, Expm, BindExpm);
QG_XLA_DEFINE(expm_fwd_vjp, ExpmFwdVjp, BindExpmFwdVjp);
QG_XLA_DEFINE(expm_fwd_save, ExpmFwdSave, BindExpm);
QG_XLA_DEFINE(expm_bwd_saved, ExpmBwdSaved, BindExpmBwdSaved);
QG_XLA_DEFINE(conj_transpose, ConjTranspose, BindUnaryC);
QG_XLA_DEFINE(expect_ket_fwd, ExpectKetFwd, BindExpectFwd);
QG_XLA_DEFINE(expect_ket_bwd, ExpectKetBwd, BindExpectBwd);
QG_XLA_DEFINE(expect_dm_fwd, ExpectDmFwd, BindExpectFwd);
QG_XLA_DEFINE(expect_dm_bwd, ExpectDmBwd, BindExpectBwd);
QG_XLA_DEFINE(fidelity_ket_fwd, FidelityKetFwd, BindFidelityFwd);
QG_XLA_DEFINE(fidelity_ket_bwd, FidelityKetBwd, BindFidelityBwd);
QG_XLA_DEFINE(fidelity_dm_fwd, FidelityDmFwd, BindFidelityFwd);
QG_XLA_DEFINE(fidelity_dm_bwd, FidelityDmBwd, BindFidelityBwd);
QG_XLA_DEFINE(displace_fwd, DisplaceFwd, BindDisplaceFwd);
QG_XLA_DEFINE(displace_bwd, DisplaceBwd, BindDisplaceBwd);
QG_XLA_DEFINE(displace_apply_fwd, DisplaceApplyFwd, BindDisplaceBwd);      // same operand list: (evecs, evals, alpha, x) -> y
QG_XLA_DEFINE(displace_apply_bwd, DisplaceApplyBwd, BindDisplaceApplyBwd);
QG_XLA_DEFINE(givens_unitary_fwd, GivensUnitaryFwd, BindGivensFwd);
QG_XLA_DEFINE(givens_unitary_bwd, GivensUnitaryBwd, BindGivensBwd);
QG_XLA_DEFINE(make_rot_fwd, MakeRotFwd, BindMakeRotFwd);
QG_XLA_DEFINE(make_rot_bwd, MakeRotBwd, BindMakeRotBwd);
QG_XLA_DEFINE(rot_zyz_fwd, RotZyzFwd, BindRotZyzFwd);
QG_XLA_DEFINE(rot_zyz_bwd, RotZyzBwd, BindRotZyzBwd);
QG_XLA_DEFINE(rot_fidelity_fwd_grad, RotFidelityFwdGrad, BindRotFidelity);
QG_XLA_DEFINE(unitary_cost_fwd_grad, UnitaryCostFwdGrad, BindUnitaryCost);
QG_XLA_DEFINE(snap_cost_fwd_grad, SnapCostFwdGrad, BindSnapCost);
QG_XLA_DEFINE(pauli_hamiltonian, PauliHamiltonian, BindPauliHamiltonian);
QG_XLA_DEFINE(pauli_project, PauliProject, BindPauliProject);
QG_XLA_DEFINE(gemm, Gemm, BindGemm);
QG_XLA_DEFINE(overlap_loss, OverlapLoss, BindOverlapLoss);
QG_XLA_DEFINE(adam_step, AdamStep, BindAdamStep);
QG_XLA_DEFINE(adam_step_device, AdamStepDevice, BindAdamStepDevice);
QG_XLA_DEFINE(eigh, Eigh, BindEigh);
QG_XLA_DEFINE(sqrtm_psd, SqrtmPsd, BindUnaryC);
QG_XLA_DEFINE(fidelity_uhlmann, FidelityUhlmann, BindFidelityFwd);
QG_XLA_DEFINE(herm_stats, HermStats, BindHermStats);
QG_XLA_DEFINE(rand_ket, RandKet, BindRandKet);
QG_XLA_DEFINE(rand_hermitian, RandHermitian, BindRandHermitian);
QG_XLA_DEFINE(rand_uniform, RandUniform, BindRandUniform);

#endif  // QG_HAVE_XLA_FFI
