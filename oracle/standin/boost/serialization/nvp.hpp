#include "l3d_standin_boost.h"
