// STUB of the subset of XLA's typed-FFI C++ API (xla/ffi/api/ffi.h, shipped with jaxlib - absent from this image) that
// netket_b200/csrc/ffi/nkb200_jax_ffi.cc uses.  TEST INFRASTRUCTURE: it lets tests/test_ffi_shim.py type-check the shim without
// jaxlib - every handler must be callable with exactly the context / argument / result / attribute types its binding declares,
// in that order, and return ffi::Error.  It is written from the public documentation of the API (names and shapes of
// Ffi::Bind().Ctx/Arg/Ret/Attr().To, AnyBuffer, Buffer<dtype>, Result<T>, Error, XLA_FFI_DEFINE_HANDLER_SYMBOL); it registers
// nothing and cannot run a call.  A build against the real header remains the authority (build line in the shim's header).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum class DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType PRED = DataType::PRED, S8 = DataType::S8, S16 = DataType::S16, S32 = DataType::S32, S64 = DataType::S64,
                          U8 = DataType::U8, U16 = DataType::U16, U32 = DataType::U32, U64 = DataType::U64, F16 = DataType::F16,
                          F32 = DataType::F32, F64 = DataType::F64, BF16 = DataType::BF16;

template <DataType dt>
struct NativeTypeOf;
template <> struct NativeTypeOf<DataType::S8> { using type = int8_t; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeTypeOf<DataType::U64> { using type = uint64_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::F64> { using type = double; };

enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal, kUnimplemented };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  const std::string &message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span(const T *d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T &operator[](size_t i) const { return d_[i]; }
  const T *begin() const { return d_; }
  const T *end() const { return d_ + n_; }

 private:
  const T *d_;
  size_t n_;
};

class AnyBuffer {
 public:
  using Dimensions = Span<int64_t>;
  DataType element_type() const { return dt_; }
  Dimensions dimensions() const { return Dimensions(dims_, rank_); }
  void *untyped_data() const { return data_; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }

 private:
  DataType dt_ = DataType::F32;
  void *data_ = nullptr;
  const int64_t *dims_ = nullptr;
  size_t rank_ = 0;
};

template <DataType dt>
class Buffer {
 public:
  using Dimensions = Span<int64_t>;
  using T = typename NativeTypeOf<dt>::type;
  DataType element_type() const { return dt; }
  Dimensions dimensions() const { return Dimensions(dims_, rank_); }
  void *untyped_data() const { return data_; }
  T *typed_data() const { return static_cast<T *>(data_); }
  size_t element_count() const { return 0; }

 private:
  void *data_ = nullptr;
  const int64_t *dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T *operator->() { return &v_; }
  T &operator*() { return v_; }

 private:
  T v_;
};

template <typename T>
struct PlatformStream {};

namespace stub {
template <typename T> struct CtxOf;
template <typename T> struct CtxOf<PlatformStream<T>> { using type = T; };
template <typename... Ts> struct List {};

template <typename Fn, typename L> struct InvocableWith;
template <typename Fn, typename... Ts>
struct InvocableWith<Fn, List<Ts...>> {
  static constexpr bool value = std::is_invocable_r<Error, Fn, Ts...>::value;
};
}  // namespace stub

template <typename... Ps>
class Binding {
 public:
  template <typename T> Binding<Ps..., typename stub::CtxOf<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ps..., T> Arg() const { return {}; }
  template <typename T> Binding<Ps..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ps..., T> Attr(const char * /*name*/) const { return {}; }
  template <typename Fn>
  int To(Fn /*fn*/) const {
    static_assert(stub::InvocableWith<Fn, stub::List<Ps...>>::value,
                  "the handler's parameters do not match the binding (context, arguments, results, attributes: in that order)");
    return 0;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)   \
  extern "C" int name() {                                    \
    static const int handler = (binding).To(impl);           \
    return handler;                                          \
  }
