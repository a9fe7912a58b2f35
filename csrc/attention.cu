// placeholder translation unit until the tcgen05 attention kernel lands
#include "launchers.h"
