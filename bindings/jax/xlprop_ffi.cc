// XLA-FFI handlers over the C ABI of include/xlprop.h -- the JAX side of the drop-in boundary (BASELINE.json north star:
// "a thin C-ABI layer registered as JAX FFI custom calls").
//
// STATUS: written against the public XLA FFI API (xla/ffi/api/ffi.h as shipped with jaxlib >= 0.4.31) but NEVER COMPILED
// OR RUN in the build image of this repository: jax/jaxlib and the FFI headers are not installable there (no network, no
// wheel).  Everything these handlers call is built and tested (tests/, ctypes + torch skin); the handlers themselves are a
// mechanical wrapping: operands/results -> device pointers, static arguments -> attributes, XLA's stream and scratch
// allocator -> the `stream` / `ws` parameters.  Build on a JAX box:
//
//   g++ -O2 -std=c++17 -shared -fPIC xlprop_ffi.cc -o libxlprop_jax.so \
//       -I$(python -c "from jax import ffi; print(ffi.include_dir())") -I../../include -I/usr/local/cuda/include \
//       -L../../xlumina_b200 -lxlprop -Wl,-rpath,'$ORIGIN/../../xlumina_b200'
//
// Conventions: complex64 buffers, z = one float64 on the device, flags = 0 (JAX cotangent convention: plain transposes,
// no conjugation), leading dimensions of `field` are a batch sharing z (nfields).  No allocation, no synchronisation, no
// host callbacks inside a handler => kCmdBufferCompatible (CUDA-graph capturable), after one warm-up call per device has
// uploaded the twiddle table.
#include <cstddef>
#include <cstdint>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"
#include "xlprop.h"

namespace ffi = xla::ffi;
using C64 = ffi::Buffer<ffi::C64>;
using F64 = ffi::Buffer<ffi::F64>;
using U8 = ffi::Buffer<ffi::U8>;
template <class B> using Res = ffi::Result<B>;

static ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc == XL_E_BAD_ARG || rc == XL_E_UNSUPPORTED ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, xl_last_error());
}
static int Side(const C64& b) { return static_cast<int>(b.dimensions().back()); }
static int Fields(const C64& b, int n) { return static_cast<int>(b.element_count() / (static_cast<size_t>(n) * n)); }
#define XL_SCRATCH(ptr, bytes)                                                              \
  auto ptr##_opt = scratch.Allocate(bytes);                                                 \
  if (!ptr##_opt) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "xlprop: workspace"); \
  void* ptr = *ptr##_opt

// ------------------------------------------------------------------------------------------------ scalar RS (wave_optics.py:281-289)
static ffi::Error RsFwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 field, F64 z, double dx, double dy, double k,
                        Res<C64> out, Res<U8> H) {
  const int n = Side(field), f = Fields(field, n);
  const size_t wsb = xl_rs_workspace_bytes(n, f, 0);
  XL_SCRATCH(ws, wsb);
  return Status(xl_rs_fwd(field.typed_data(), out->typed_data(), H->typed_data(), z.typed_data(), n, f, dx, dy, k, 0, ws, wsb, stream));
}
// cotangents of (field, z); needs the primal input and output for the exact i*k*out part of d out/dz (xlprop.h)
static ffi::Error RsBwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 field, C64 primal_out, C64 ct_out, U8 H, F64 z,
                        double dx, double dy, double k, Res<C64> ct_field, Res<F64> ct_z) {
  const int n = Side(field), f = Fields(field, n);
  const size_t wsb = xl_rs_workspace_bytes(n, f, 1);
  XL_SCRATCH(ws, wsb);
  if (cudaMemsetAsync(ct_z->typed_data(), 0, sizeof(double), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "xlprop: cudaMemsetAsync");
  return Status(xl_rs_bwd(field.typed_data(), primal_out.typed_data(), ct_out.typed_data(), ct_field->typed_data(),
                          ct_z->typed_data(), H.typed_data(), z.typed_data(), n, f, dx, dy, k, 0, ws, wsb, stream));
}

// ------------------------------------------------------------------------------------------------ vectorial RS, Ez formed in the kernel
// (vectorized_optics.py:244-284 + 364-373): exy (2,N,N) -> (3,N,N)
static ffi::Error VrsFwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 exy, F64 z, double x0, double y0, double dx,
                         double dy, double k, Res<C64> out, Res<U8> H) {
  const int n = Side(exy);
  const size_t wsb = xl_rs_workspace_bytes(n, 3, 0);
  XL_SCRATCH(ws, wsb);
  return Status(xl_vrs_fwd(exy.typed_data(), nullptr, out->typed_data(), H->typed_data(), z.typed_data(), n, x0, y0, dx, dy, k, 0, ws, wsb, stream));
}
static ffi::Error VrsBwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 exy, C64 primal_out, C64 ct_out, U8 H, F64 z,
                         double x0, double y0, double dx, double dy, double k, Res<C64> ct_exy, Res<F64> ct_z) {
  const int n = Side(exy);
  const size_t wsb = xl_rs_workspace_bytes(n, 3, 1);
  XL_SCRATCH(ws, wsb);
  if (cudaMemsetAsync(ct_z->typed_data(), 0, sizeof(double), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "xlprop: cudaMemsetAsync");
  return Status(xl_vrs_bwd(exy.typed_data(), nullptr, primal_out.typed_data(), ct_out.typed_data(), ct_exy->typed_data(), ct_z->typed_data(),
                           H.typed_data(), z.typed_data(), n, x0, y0, dx, dy, k, 0, ws, wsb, stream));
}

// ------------------------------------------------------------------------------------------------ CZT / VCZT (wave_optics.py:333-357,
// vectorized_optics.py:321-361, 375-384).  `vectorial` = 0: field (N,N) -> (My,Mx); 1: exy (2,N,N) -> (3,My,Mx).
// z is a device scalar like everywhere else, but it has no cotangent: no reference caller differentiates a CZT distance
// (SURVEY.md 8f-4); the Python side raises if asked to.
static ffi::Error CztFwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 in, F64 z, double wavelength, int64_t vectorial,
                         double x0, double dx, double y0, double dy, double xo0, double xol, double yo0, double yol, Res<C64> out) {
  const int n = Side(in);
  const auto od = out->dimensions();
  const int my = static_cast<int>(od[od.size() - 2]), mx = static_cast<int>(od[od.size() - 1]);
  const size_t wsb = xl_czt_workspace_bytes(n, mx, my, static_cast<int>(vectorial));
  if (wsb == 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "xlprop: CZT sizes unsupported (m+M-1 a power of two, or padded length > 4096)");
  XL_SCRATCH(ws, wsb);
  return Status(xl_czt_fwd(in.typed_data(), nullptr, out->typed_data(), z.typed_data(), wavelength, n, mx, my, static_cast<int>(vectorial),
                           x0, dx, y0, dy, xo0, xol, yo0, yol, 0, ws, wsb, stream));
}
static ffi::Error CztBwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 ct_out, F64 z, double wavelength, int64_t vectorial,
                         double x0, double dx, double y0, double dy, double xo0, double xol, double yo0, double yol, Res<C64> ct_in) {
  const int n = static_cast<int>(ct_in->dimensions().back());
  const auto od = ct_out.dimensions();
  const int my = static_cast<int>(od[od.size() - 2]), mx = static_cast<int>(od[od.size() - 1]);
  const size_t wsb = xl_czt_workspace_bytes(n, mx, my, static_cast<int>(vectorial));
  XL_SCRATCH(ws, wsb);
  return Status(xl_czt_bwd(ct_out.typed_data(), ct_in->typed_data(), z.typed_data(), wavelength, n, mx, my, static_cast<int>(vectorial),
                           x0, dx, y0, dy, xo0, xol, yo0, yol, 0, ws, wsb, stream));
}

// ------------------------------------------------------------------------------------------------ high-NA objective focusing
// (optical_elements.py:515-638): exy (2,N,N) -> (3,My,Mx); lens matrix, apodisation and the constant fused
static ffi::Error HighnaFwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 exy, double radius, double f, double wavelength,
                            double x0, double dx, double y0, double dy, double xo0, double xol, double yo0, double yol, Res<C64> out) {
  const int n = Side(exy);
  const auto od = out->dimensions();
  const int my = static_cast<int>(od[od.size() - 2]), mx = static_cast<int>(od[od.size() - 1]);
  const size_t wsb = xl_highna_workspace_bytes(n, mx, my);
  if (wsb == 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "xlprop: high-NA sizes unsupported");
  XL_SCRATCH(ws, wsb);
  return Status(xl_highna_fwd(exy.typed_data(), nullptr, out->typed_data(), n, mx, my, radius, f, wavelength, x0, dx, y0, dy, xo0, xol, yo0, yol,
                              0, ws, wsb, stream));
}
static ffi::Error HighnaBwd(cudaStream_t stream, ffi::ScratchAllocator scratch, C64 ct_out, double radius, double f, double wavelength,
                            double x0, double dx, double y0, double dy, double xo0, double xol, double yo0, double yol, Res<C64> ct_exy) {
  const int n = static_cast<int>(ct_exy->dimensions().back());
  const auto od = ct_out.dimensions();
  const int my = static_cast<int>(od[od.size() - 2]), mx = static_cast<int>(od[od.size() - 1]);
  const size_t wsb = xl_highna_workspace_bytes(n, mx, my);
  XL_SCRATCH(ws, wsb);
  return Status(xl_highna_bwd(ct_out.typed_data(), ct_exy->typed_data(), n, mx, my, radius, f, wavelength, x0, dx, y0, dy, xo0, xol, yo0,
                              yol, 0, ws, wsb, stream));
}

// ------------------------------------------------------------------------------------------------ bindings
#define XL_CTX ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
#define XL_GEOM4 .Attr<double>("x0").Attr<double>("dx").Attr<double>("y0").Attr<double>("dy")
#define XL_GEOM_OUT .Attr<double>("xo0").Attr<double>("xol").Attr<double>("yo0").Attr<double>("yol")
#define XL_TRAITS {ffi::Traits::kCmdBufferCompatible}

XLA_FFI_DEFINE_HANDLER_SYMBOL(XlRsFwd, RsFwd,
    XL_CTX.Arg<C64>().Arg<F64>().Attr<double>("dx").Attr<double>("dy").Attr<double>("k").Ret<C64>().Ret<U8>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlRsBwd, RsBwd,
    XL_CTX.Arg<C64>().Arg<C64>().Arg<C64>().Arg<U8>().Arg<F64>().Attr<double>("dx").Attr<double>("dy").Attr<double>("k")
        .Ret<C64>().Ret<F64>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlVrsFwd, VrsFwd,
    XL_CTX.Arg<C64>().Arg<F64>().Attr<double>("x0").Attr<double>("y0").Attr<double>("dx").Attr<double>("dy").Attr<double>("k")
        .Ret<C64>().Ret<U8>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlVrsBwd, VrsBwd,
    XL_CTX.Arg<C64>().Arg<C64>().Arg<C64>().Arg<U8>().Arg<F64>().Attr<double>("x0").Attr<double>("y0").Attr<double>("dx")
        .Attr<double>("dy").Attr<double>("k").Ret<C64>().Ret<F64>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlCztFwd, CztFwd,
    XL_CTX.Arg<C64>().Arg<F64>().Attr<double>("wavelength").Attr<int64_t>("vectorial") XL_GEOM4 XL_GEOM_OUT.Ret<C64>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlCztBwd, CztBwd,
    XL_CTX.Arg<C64>().Arg<F64>().Attr<double>("wavelength").Attr<int64_t>("vectorial") XL_GEOM4 XL_GEOM_OUT.Ret<C64>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlHighnaFwd, HighnaFwd,
    XL_CTX.Arg<C64>().Attr<double>("radius").Attr<double>("f").Attr<double>("wavelength") XL_GEOM4 XL_GEOM_OUT.Ret<C64>(), XL_TRAITS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(XlHighnaBwd, HighnaBwd,
    XL_CTX.Arg<C64>().Attr<double>("radius").Attr<double>("f").Attr<double>("wavelength") XL_GEOM4 XL_GEOM_OUT.Ret<C64>(), XL_TRAITS);
