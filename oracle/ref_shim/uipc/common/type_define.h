// Stand-in for the reference's include/uipc/common/type_define.h (TEST INFRASTRUCTURE): same scalar / small-matrix names on the
// minimal Eigen subset of ../../mini_eigen.h.
#pragma once
#include "../../type_define.h"
