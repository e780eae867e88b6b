// XLA-FFI binding of the qgrad_b200 C ABI (include/qgrad_b200.h) -- the "thin XLA-FFI layer" the
// north star names, so that a JAX host registers each op as a primitive with a custom_vjp.
//
// STATUS: written to the public XLA FFI C++ API (xla/ffi/api/ffi.h, as shipped by
// `jax.ffi.include_dir()` in JAX >= 0.4.31 / 0.5); it can be neither compiled nor run in this
// environment -- JAX/jaxlib are absent and not installable, and no
// xla/ffi headers exist on disk.  The whole file is therefore guarded by __has_include and is
// not part of qgrad_b200/build.py's unit list.  Build, next to a JAX install:
//
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include qgrad_b200/csrc/xla_ffi_shim.cc -L qgrad_b200 -lqgrad_b200 \
//       -o qgrad_b200/libqgrad_b200_xla.so
//
// Python side: INTEGRATION.md section 2.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define QG_HAVE_XLA_FFI 1
#endif
#endif

#ifdef QG_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
#include "../../include/qgrad_b200.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc, const char* what) {
  if (rc == QG_OK) return ffi::Error::Success();
  return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + qg_status_string(rc));
}

template <typename Buf>
std::pair<int64_t, int> BatchAndDim(const Buf& a) {       // (..., n, n) -> (prod(...), n)
  auto dims = a.dimensions();
  int64_t batch = 1;
  for (size_t i = 0; i + 2 < dims.size(); ++i) batch *= dims[i];
  return {batch, static_cast<int>(dims.back())};
}

// Y = expm(A).  Workspace is an XLA-allocated scratch result (size from qg_expm_workspace_bytes,
// computed on the Python side and passed as the `ws` result buffer), so the library owns nothing.
template <int kDtype, ffi::DataType kXla>
ffi::Error ExpmImpl(cudaStream_t stream, ffi::Buffer<kXla> a, ffi::Result<ffi::Buffer<kXla>> y,
                    ffi::Result<ffi::Buffer<ffi::U8>> ws, int64_t max_squarings) {
  auto [batch, n] = BatchAndDim(a);
  return Status(qg_expm(kDtype, a.untyped_data(), y->untyped_data(), batch, n, ws->untyped_data(),
                        ws->element_count(), static_cast<int>(max_squarings), stream), "qg_expm");
}

// (Y, Abar) = fused value + pull-back; adjoint = QG_ADJ_TRANS gives JAX's VJP L(A^T, ct) directly.
template <int kDtype, ffi::DataType kXla>
ffi::Error ExpmFwdVjpImpl(cudaStream_t stream, ffi::Buffer<kXla> a, ffi::Buffer<kXla> ct,
                          ffi::Result<ffi::Buffer<kXla>> y, ffi::Result<ffi::Buffer<kXla>> abar,
                          ffi::Result<ffi::Buffer<ffi::U8>> ws, int64_t max_squarings) {
  auto [batch, n] = BatchAndDim(a);
  return Status(qg_expm_fwd_vjp(kDtype, a.untyped_data(), ct.untyped_data(), y->untyped_data(),
                                abar->untyped_data(), batch, n, QG_ADJ_TRANS, ws->untyped_data(),
                                ws->element_count(), static_cast<int>(max_squarings), stream),
                "qg_expm_fwd_vjp");
}

template <int kDtype, ffi::DataType kXla>
ffi::Error ExpectKetFwdImpl(cudaStream_t stream, ffi::Buffer<kXla> oper, ffi::Buffer<kXla> ket,
                            ffi::Result<ffi::Buffer<kXla>> out) {
  auto kd = ket.dimensions();                                 // (batch, n)
  const int64_t batch = kd.size() > 1 ? kd[0] : 1;
  const int n = static_cast<int>(kd.back());
  const int oper_batched = oper.dimensions().size() == 3 ? 1 : 0;
  return Status(qg_expect_ket_fwd(kDtype, oper.untyped_data(), ket.untyped_data(), out->untyped_data(),
                                  batch, n, oper_batched, stream), "qg_expect_ket_fwd");
}

template <int kDtype, ffi::DataType kXla, ffi::DataType kReal>
ffi::Error FidelityKetFwdImpl(cudaStream_t stream, ffi::Buffer<kXla> a, ffi::Buffer<kXla> b,
                              ffi::Result<ffi::Buffer<kReal>> out) {
  auto d = a.dimensions();
  const int64_t batch = d.size() > 1 ? d[0] : 1;
  return Status(qg_fidelity_ket_fwd(kDtype, a.untyped_data(), b.untyped_data(), out->untyped_data(), batch,
                                    static_cast<int>(d.back()), stream), "qg_fidelity_ket_fwd");
}

}  // namespace

#define QG_EXPM_BINDING(T)                                                        \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<T>>() \
      .Ret<ffi::Buffer<T>>().Ret<ffi::Buffer<ffi::U8>>().Attr<int64_t>("max_squarings")
#define QG_EXPM_VJP_BINDING(T)                                                    \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<T>>() \
      .Arg<ffi::Buffer<T>>().Ret<ffi::Buffer<T>>().Ret<ffi::Buffer<T>>()          \
      .Ret<ffi::Buffer<ffi::U8>>().Attr<int64_t>("max_squarings")

XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_expm_c128, (ExpmImpl<QG_C128, ffi::C128>), QG_EXPM_BINDING(ffi::C128));
XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_expm_c64, (ExpmImpl<QG_C64, ffi::C64>), QG_EXPM_BINDING(ffi::C64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_expm_fwd_vjp_c128, (ExpmFwdVjpImpl<QG_C128, ffi::C128>),
                              QG_EXPM_VJP_BINDING(ffi::C128));
XLA_FFI_DEFINE_HANDLER_SYMBOL(qg_xla_expm_fwd_vjp_c64, (ExpmFwdVjpImpl<QG_C64, ffi::C64>),
                              QG_EXPM_VJP_BINDING(ffi::C64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    qg_xla_expect_ket_fwd_c128, (ExpectKetFwdImpl<QG_C128, ffi::C128>),
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::C128>>()
        .Arg<ffi::Buffer<ffi::C128>>().Ret<ffi::Buffer<ffi::C128>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    qg_xla_fidelity_ket_fwd_c128, (FidelityKetFwdImpl<QG_C128, ffi::C128, ffi::F64>),
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::C128>>()
        .Arg<ffi::Buffer<ffi::C128>>().Ret<ffi::Buffer<ffi::F64>>());
// The remaining ops (expect_dm, fidelity_dm, displace, givens_unitary, make_rot, rot_zyz and every
// *_bwd) bind identically: one handler per (op, dtype), buffers in the order of the C prototype.

#endif  // QG_HAVE_XLA_FFI
