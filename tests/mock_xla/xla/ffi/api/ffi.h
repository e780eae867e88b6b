// MOCK of the public XLA FFI C++ API header (xla/ffi/api/ffi.h) -- TEST INFRASTRUCTURE ONLY.
//
// JAX / jaxlib are not installable in this environment, so the real header is not on disk.  This mock declares just
// the names qgrad_b200/csrc/xla_ffi_shim.cc uses, with the shapes the public API gives them (written from the public
// documentation of jax.ffi / XLA FFI; not copied from XLA), so that the shim can be TYPE-CHECKED here
// (g++ -fsyntax-only, tests/test_cpu_host.py) against the real include/qgrad_b200.h: every handler must be
// invocable with exactly the types its binding decodes (Ctx<PlatformStream<S>> -> S, Arg<Buffer<T>> -> Buffer<T>,
// Ret<Buffer<T>> -> Result<Buffer<T>>, Attr<T> -> T), and every call into the C ABI must match its prototype.
// It proves nothing about XLA's runtime behaviour.
#ifndef QG_MOCK_XLA_FFI_API_FFI_H_
#define QG_MOCK_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum class DataType : uint8_t { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, BF16, F32, F64, C64, C128 };
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType U8 = DataType::U8;
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType C64 = DataType::C64;
inline constexpr DataType C128 = DataType::C128;

template <typename T>
class Span {
 public:
  size_t size() const { return v_.size(); }
  const T& operator[](size_t i) const { return v_[i]; }
  const T& back() const { return v_.back(); }
  const T* begin() const { return v_.data(); }
  const T* end() const { return v_.data() + v_.size(); }
 private:
  std::vector<T> v_;
};

template <DataType dtype>
class Buffer {
 public:
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
 private:
  void* data_ = nullptr;
  Span<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};

enum class ErrorCode : uint8_t { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T> struct Decoded { using type = T; };
template <typename S> struct Decoded<PlatformStream<S>> { using type = S; };
template <typename T> struct RetTag {};
template <typename T> struct Decoded<RetTag<T>> { using type = Result<T>; };
}  // namespace internal

template <typename... Ts>
class Binding {
 public:
  template <typename T> Binding<Ts..., T> Ctx() && { return {}; }
  template <typename T> Binding<Ts..., T> Arg() && { return {}; }
  template <typename T> Binding<Ts..., internal::RetTag<T>> Ret() && { return {}; }
  template <typename T> Binding<Ts..., T> Attr(std::string) && { return {}; }
  // the real Binding::To() builds a Handler; the mock only checks the call signature
  template <typename Fn>
  static constexpr bool Accepts() {
    return std::is_invocable_r<Error, Fn, typename internal::Decoded<Ts>::type...>::value;
  }
};

class Ffi {
 public:
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame*)`; the mock defines a symbol of that name
// and statically asserts that `impl` takes exactly what `binding` decodes
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                                   \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(), #fn ": handler signature does not match its binding"); \
  extern "C" void* fn(void*) { return nullptr; }

#endif  // QG_MOCK_XLA_FFI_API_FFI_H_
