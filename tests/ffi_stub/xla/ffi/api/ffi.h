// TEST INFRASTRUCTURE -- a declaration-only stand-in for jaxlib's xla/ffi/api/ffi.h (absent from this image), just
// large enough to TYPE-CHECK vbsky_b200/csrc/ffi/vbsky_ffi.cc: buffer / span / error types with the members the handlers
// use, and a binding builder that records the decoded C++ type of every Ctx/Attr/Arg/Ret so that
// XLA_FFI_DEFINE_HANDLER_SYMBOL can static_assert -- as the real header does -- that the handler is invocable with exactly
// those types.  Written from the documented API (jax.ffi / XLA FFI); not a copy of the real header.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
template <DataType>
struct NativeOf;
template <>
struct NativeOf<F64> { using type = double; };
template <>
struct NativeOf<F32> { using type = float; };
template <>
struct NativeOf<S32> { using type = int32_t; };
template <>
struct NativeOf<S64> { using type = int64_t; };
template <>
struct NativeOf<U8> { using type = uint8_t; };

template <class T>
struct Span {
  const T* begin() const;
  const T* end() const;
  size_t size() const;
  const T& operator[](size_t) const;
};

template <DataType dt>
struct Buffer {
  using T = typename NativeOf<dt>::type;
  T* typed_data() const;
  void* untyped_data() const;
  size_t element_count() const;
  size_t size_bytes() const;
  Span<const int64_t> dimensions() const;
};

template <class B>
struct Result {
  B* operator->() const;
  B& operator*() const;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal };
struct Error {
  Error(ErrorCode, std::string);
  static Error Success();
};

template <class S>
struct PlatformStream {};
template <class C>
struct CtxDecoded { using type = C; };
template <class S>
struct CtxDecoded<PlatformStream<S>> { using type = S; };

template <class... Ts>
struct Binding {
  template <class C>
  Binding<Ts..., typename CtxDecoded<C>::type> Ctx();
  template <class A>
  Binding<Ts..., A> Attr(const char*);
  template <class A>
  Binding<Ts..., A> Arg();
  template <class R>
  Binding<Ts..., Result<R>> Ret();
};
struct Ffi {
  static Binding<> Bind();
};

template <class F, class B>
struct HandlerMatches;
template <class F, class... Ts>
struct HandlerMatches<F, Binding<Ts...>> : std::is_invocable_r<Error, F, Ts...> {};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                      \
  static_assert(::xla::ffi::HandlerMatches<decltype(&impl), decltype(binding)>::value,                          \
                #name ": the handler's parameters do not match its binding");                                  \
  extern "C" void* name(void*)
