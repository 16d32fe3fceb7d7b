// Drop-in include: put include/ceres_b200/compat first on the include path and the reference's
// `#include <ceres/ceres.h>` / `#include "ceres/autodiff_cost_function.h"` resolve to the B200 mirror.
#ifndef CERES_B200_AS_CERES
#define CERES_B200_AS_CERES
#endif
#include "../../ceres.h"
