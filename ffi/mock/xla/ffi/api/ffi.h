// MOCK of the subset of jaxlib's "xla/ffi/api/ffi.h" that ffi/larnd_ffi.cc uses — NOT the real header.
//
// jax / jaxlib are absent from the build image and from the GPU box (profiles/r2_probe_jax.txt), so the real header
// (jax.ffi.include_dir()) cannot be used to compile the shim here.  This file reproduces the public SHAPE of the API the
// shim relies on (Buffer / Result / Error / Ffi::Bind().Ctx().Arg().Attr().Ret() / XLA_FFI_DEFINE_HANDLER_SYMBOL) so that a
// CPU test can at least prove that larnd_ffi.cc is well-formed C++ against that shape and that every handler's signature
// matches its binding (the .To() below static_asserts invocability).  __graft_entry__.build() compiles the shim against the
// REAL header the moment `import jax` works and never uses this directory then.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum class DataType { U8, S32, S64, F32 };
inline constexpr DataType U8 = DataType::U8, S32 = DataType::S32, S64 = DataType::S64, F32 = DataType::F32;
template <DataType> struct NativeTypeOf;
template <> struct NativeTypeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };

template <typename T> struct Span {
  const T* ptr = nullptr; size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
  const T* begin() const { return ptr; }
  const T* end() const { return ptr + n; }
};

template <DataType dtype> class Buffer {
 public:
  using T = typename NativeTypeOf<dtype>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return {dims_.data(), dims_.size()}; }
  size_t element_count() const { size_t n = 1; for (auto d : dims_) n *= (size_t)d; return n; }
  size_t size_bytes() const { return element_count() * sizeof(T); }
 private:
  T* data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <typename T> class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T> struct PlatformStream {};

namespace internal {
template <typename T> struct CtxType;
template <typename T> struct CtxType<PlatformStream<T>> { using type = T; };
template <typename T> struct RetType { using type = Result<T>; };
struct HandlerBase { virtual ~HandlerBase() = default; XLA_FFI_Error* Call(XLA_FFI_CallFrame*) { return nullptr; } };
template <typename... Ts> struct Handler : HandlerBase {};
}  // namespace internal

template <typename... Ts> class Binding {
 public:
  template <typename T> Binding<Ts..., typename internal::CtxType<T>::type> Ctx() && { return {}; }
  template <typename T> Binding<Ts..., T> Arg() && { return {}; }
  template <typename T> Binding<Ts..., typename internal::RetType<T>::type> Ret() && { return {}; }
  template <typename T> Binding<Ts..., T> Attr(std::string) && { return {}; }
  template <typename Fn> internal::Handler<Ts...>* To(Fn&&) && {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "handler signature does not match its binding");
    return new internal::Handler<Ts...>();
  }
};

class Ffi {
 public:
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                      \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) {               \
    static auto* handler = (binding).To(impl);                               \
    return handler->Call(call_frame);                                        \
  }
