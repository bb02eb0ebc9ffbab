// defines.hpp — compile-time configuration of the header shim (stands in for the reference's src/defines.hpp:4-66).
//
// The reference selects lattice, scenario and grid by editing this file.  Here everything can come from the
// command line (-DNX= -DNY= -DSCALE=, -DUSE_TAYLOR_GREEN ...); when only a USE_* macro is given, the grid the
// reference ships for that scenario is used (src/defines.hpp:20-35).  Only the D2Q9 lattice exists on this path.
#ifndef DEFINES_H
#define DEFINES_H

#ifdef D3Q27
#error "the B200 engine covers the reference's D2Q9 path only (SURVEY.md 8f: D3Q27 is out of scope)"
#endif
#ifndef D2Q9
#define D2Q9
#endif

#if !defined(NX) || !defined(NY)
#  if defined(NX) || defined(NY)
#    error "give both -DNX= and -DNY= (or neither, with a USE_* scenario macro)"
#  endif
#  if defined(USE_TAYLOR_GREEN)
#    ifndef SCALE
#      define SCALE 2
#    endif
#    define NX (128 * SCALE)
#    define NY (128 * SCALE)
#  elif defined(USE_POISEUILLE)
#    ifndef SCALE
#      define SCALE 1
#    endif
#    define NX (150 * SCALE)
#    define NY (100 * SCALE)
#  elif defined(USE_LID_DRIVEN)
#    ifndef SCALE
#      define SCALE 1
#    endif
#    define NX (129 * SCALE)
#    define NY (129 * SCALE)
#  elif defined(USE_FLOW_PAST_CYLINDER)
#    ifndef SCALE
#      define SCALE 1
#    endif
#    define NX (256 * SCALE)
#    define NY (128 * SCALE)
#  else
#    error "no grid: pass -DNX= -DNY= or one of -DUSE_TAYLOR_GREEN / -DUSE_POISEUILLE / -DUSE_LID_DRIVEN / -DUSE_FLOW_PAST_CYLINDER"
#  endif
#endif
#ifndef SCALE
#define SCALE 1
#endif
#ifndef NZ
#define NZ 1
#endif

// thread-block edge of the reference's own kernels; kept because scenario code may name it
#define BLOCK_SIZE 16

#endif  // DEFINES_H
