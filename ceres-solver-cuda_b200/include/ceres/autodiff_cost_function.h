#include "ceres/cost_function.h"
