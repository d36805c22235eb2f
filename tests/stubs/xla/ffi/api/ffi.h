// Minimal STAND-IN for jaxlib's xla/ffi/api/ffi.h, for tests/test_ffi_sources.py only: JAX is not installable in the
// development image, so ffi/zenflow_b200_xla.cc is type-checked against this declaration-only subset of the public
// XLA FFI C++ API (the names and signatures it uses).  It is NOT the XLA header and nothing ships with it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>

typedef struct CUstream_st* cudaStream_t;
extern "C" int cudaMemsetAsync(void* ptr, int value, size_t count, cudaStream_t stream);

namespace xla::ffi {

enum DataType { F32, F64, S32 };
enum class ErrorCode { kInvalidArgument, kInternal, kResourceExhausted };

template <typename T>
class Span {
public:
    size_t size() const;
    const T& operator[](size_t i) const;
    const T* begin() const;
    const T* end() const;
};

class Error {
public:
    Error(ErrorCode code, std::string message);
    static Error Success();
    bool failure() const;
    bool success() const;
};

template <typename T>
class ErrorOr {
public:
    bool has_value() const;
    T& value();
    T& operator*();
};

template <DataType dtype> struct NativeTypeOf;
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };

template <DataType dtype>
class Buffer {
public:
    using T = typename NativeTypeOf<dtype>::type;
    T* typed_data() const;
    Span<const int64_t> dimensions() const;
    size_t element_count() const;
};

template <typename T>
class Result {
public:
    T* operator->() const;
    T& operator*() const;
};

class RemainingArgs {
public:
    size_t size() const;
    template <typename T> ErrorOr<T> get(size_t index) const;
};
class RemainingRets {
public:
    size_t size() const;
    template <typename T> ErrorOr<Result<T>> get(size_t index) const;
};

class ScratchAllocator {
public:
    std::optional<void*> Allocate(size_t size, size_t alignment = 1);
};
template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
    template <typename T> Binding<Ts..., T> Ctx() const;
    template <typename T> Binding<Ts..., T> Arg() const;
    template <typename T> Binding<Ts..., T> Ret() const;
    template <typename T> Binding<Ts..., T> Attr(const char* name) const;
    Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const;
    Binding<Ts..., ::xla::ffi::RemainingRets> RemainingRets() const;
};
struct Ffi { static Binding<> Bind(); };

// maps a binding's slot types to the implementation's parameter types
template <typename T> struct ParamOf { using type = T; };
template <typename T> struct ParamOf<PlatformStream<T>> { using type = T; };
template <typename... Ts, typename Fn>
constexpr bool CheckCallable(const Binding<Ts...>&, Fn*) { return true; }

}  // namespace xla::ffi

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
// the real macro instantiates the handler; here it only checks that the binding expression is well-formed
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                       \
    extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame* frame) {                   \
        (void)frame;                                                            \
        (void)::xla::ffi::CheckCallable((binding), &impl);                      \
        return nullptr;                                                         \
    }
