// Stand-in for the reference's backends/cuda/type_define.h (which pulls muda + Eigen + uipc/common/type_define.h) -- TEST
// INFRASTRUCTURE, see mini_eigen.h. Same names, float64 like the reference (uipc::Float = double).
#pragma once
#include "mini_eigen.h"
#ifndef __CUDACC__
#define __host__
#define __device__
#endif
#define UIPC_GENERIC
#define UIPC_DEVICE
#define UIPC_HOST
namespace uipc {
using Float = double;
using IndexT = int;
using Vector2i = Eigen::Matrix<IndexT, 2, 1>;
using Vector3i = Eigen::Matrix<IndexT, 3, 1>;
using Vector4i = Eigen::Matrix<IndexT, 4, 1>;
using Vector2 = Eigen::Matrix<Float, 2, 1>;
using Vector3 = Eigen::Matrix<Float, 3, 1>;
using Matrix2x2 = Eigen::Matrix<Float, 2, 2>;
using Matrix3x3 = Eigen::Matrix<Float, 3, 3>;
using Vector9 = Eigen::Matrix<Float, 9, 1>;
using Vector12 = Eigen::Matrix<Float, 12, 1>;
using Matrix9x9 = Eigen::Matrix<Float, 9, 9>;
using Matrix9x12 = Eigen::Matrix<Float, 9, 12>;
using Matrix12x12 = Eigen::Matrix<Float, 12, 12>;
template <class T, int M, int N>
using Matrix = Eigen::Matrix<T, M, N>;
template <class T, int N>
using Vector = Eigen::Matrix<T, N, 1>;
namespace backend::cuda {
    using namespace uipc;
    namespace distance {}
}
} // namespace uipc
