// placeholder translation unit (Florence-2 kernels land here)
#include "b2p_internal.h"
