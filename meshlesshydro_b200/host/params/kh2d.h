// values of /root/reference/testcases/kelvin-helmholtz/parameter_long_run.h (2D periodic Kelvin-Helmholtz)
#ifndef DEMONSTRATOR_PARAMETER_H
#define DEMONSTRATOR_PARAMETER_H
#define DIM 2
#define PERIODIC_BOUNDARIES 1
#define ADAPTIVE_TIMESTEP 1
#define CFL .4
#define MAX_NUM_INTERACTIONS 400
#define MAX_NUM_GHOST_INTERACTIONS 300
#define SLOPE_LIMITING 1
#define BETA 4.
#define PAIRWISE_LIMITER 0
#define PSI_1 .5
#define PSI_2 .25
#define MESHLESS_FINITE_MASS 0
#define ENFORCE_FLUX_SYM 1
#define MOVE_PARTICLES 1
#define DEBUG_LVL 1
#define FIRST_ORDER_QUAD_POINT 1
#define RUNSPH 0
#endif
