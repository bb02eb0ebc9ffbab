// forwarding header: the reference keeps MRT in its own file (src/core/collision/MRT/MRT.cuh)
#include "core/collision/collision.cuh"
