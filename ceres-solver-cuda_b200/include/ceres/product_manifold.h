#include "ceres/manifold.h"
