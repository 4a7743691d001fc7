// TEST INFRASTRUCTURE ONLY.  Force-included (-include) when oracle/build_ref.py compiles the reference's
// VSLAM/backend/src/matching_kernels.cu: that file dispatches on the deprecated `tensor.type()` (matching_kernels.cu:103), which
// the AT_DISPATCH macros of the installed PyTorch no longer accept.  This restores the overload older PyTorch shipped; the
// reference source itself is compiled unmodified from where it lies.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
