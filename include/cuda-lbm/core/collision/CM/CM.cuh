// forwarding header: the reference keeps CM in its own file (src/core/collision/CM/CM.cuh)
#include "core/collision/collision.cuh"
