// placeholder translation unit until the NVLink comm kernels land
#include "launchers.h"
