// core/streaming/streaming.cuh — the reference selects periodic axes with the PERIODIC_X / PERIODIC_Y macros of this
// file (src/core/streaming/streaming.cuh:8-11, edited by hand per scenario).  With the shim they are ordinary -D flags,
// or a scenario may state them itself with `static constexpr bool periodic_x / periodic_y` members.
#ifndef STREAMING_H
#define STREAMING_H
#ifdef PERIODIC_X
constexpr bool lbm_b200_periodic_x_default = true;
#else
constexpr bool lbm_b200_periodic_x_default = false;
#endif
#ifdef PERIODIC_Y
constexpr bool lbm_b200_periodic_y_default = true;
#else
constexpr bool lbm_b200_periodic_y_default = false;
#endif
#endif  // STREAMING_H
