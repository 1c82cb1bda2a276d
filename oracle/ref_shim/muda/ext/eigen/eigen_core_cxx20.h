// Stand-in for muda/ext/eigen/eigen_core_cxx20.h (TEST INFRASTRUCTURE): Eigen is not in this image, mini_eigen.h supplies the subset.
#pragma once
#include "../../../mini_eigen.h"
