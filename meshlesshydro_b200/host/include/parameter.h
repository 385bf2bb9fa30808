// Compile-time switches of the demonstrator (same macro names and meaning as
// /root/reference/demonstrator/include/parameter.h:9-74).  As in the reference (README.md:120), a test case is
// selected by swapping this header: `make PARAM=params/sedov3d.h` pre-includes another file with the same
// macros (the ones under params/ carry the values of the reference's testcases/*/parameter.h), or pass the
// reference's own testcase header.  The GPU-backed Particles class turns these macros into the run-time
// mlh_config of the C ABI (include/mlh_gpu.h), so ONE libmlh_gpu.so serves every build.
#ifndef DEMONSTRATOR_PARAMETER_H
#define DEMONSTRATOR_PARAMETER_H

/// possible values 2 or 3 for 2D or 3D simulations
#define DIM 2
/// define if periodic boundaries should be employed
#define PERIODIC_BOUNDARIES 1
/// define if timestep is adaptive
#define ADAPTIVE_TIMESTEP 1
/// Courant-Friedrichs-Levy number
#define CFL .4
/// maximum number of interactions for each particle (regular / with ghost particles)
#define MAX_NUM_INTERACTIONS 400
#define MAX_NUM_GHOST_INTERACTIONS 300
/// slope limiting and its parameter
#define SLOPE_LIMITING 1
#define BETA 4.
/// pairwise limiter
#define PAIRWISE_LIMITER 0
#define PSI_1 .5
#define PSI_2 .25
/// meshless finite mass instead of meshless finite volume
#define MESHLESS_FINITE_MASS 0
/// one-sided evaluation of the fluxes (always on in the GPU path: each face is solved once)
#define ENFORCE_FLUX_SYM 1
/// particles move with the fluid
#define MOVE_PARTICLES 1
/// 0: no additions, 1: additional checks
#define DEBUG_LVL 1
/// first order quadrature point for the Riemann problems (the only mode of the GPU path)
#define FIRST_ORDER_QUAD_POINT 1
/// SPH mode of the demonstrator: not part of the MFV hot path
#define RUNSPH 0

#endif // DEMONSTRATOR_PARAMETER_H
