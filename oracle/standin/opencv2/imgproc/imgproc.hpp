#include "l3d_standin_opencv.h"
