// forwarding header: the reference keeps BGK in its own file (src/core/collision/BGK/BGK.cuh)
#include "core/collision/collision.cuh"
