// Stand-in for finite_element/matrix_utils.h (TEST INFRASTRUCTURE): declarations of the two helpers fem_utils.cu mentions in
// functions the pin tests never call (invariant2 / invariant4); defined as traps in oracle/ref_sym.cpp.
#pragma once
#include <type_define.h>
namespace uipc::backend::cuda {
Float ddot(const Matrix3x3& A, const Matrix3x3& B);
void svd(const Matrix3x3& F, Matrix3x3& U, Vector3& Sigma, Matrix3x3& V) noexcept;
} // namespace uipc::backend::cuda
