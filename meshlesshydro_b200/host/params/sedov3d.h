// values of /root/reference/testcases/sedov/parameter.h (3D Sedov blast wave, open boundaries)
#ifndef DEMONSTRATOR_PARAMETER_H
#define DEMONSTRATOR_PARAMETER_H
#define DIM 3
#define PERIODIC_BOUNDARIES 0
#define ADAPTIVE_TIMESTEP 1
#define CFL .25
#define MAX_NUM_INTERACTIONS 400
#define MAX_NUM_GHOST_INTERACTIONS 300
#define SLOPE_LIMITING 1
#define BETA 1.
#define PAIRWISE_LIMITER 1
#define PSI_1 .5
#define PSI_2 .25
#define MESHLESS_FINITE_MASS 0
#define ENFORCE_FLUX_SYM 1
#define MOVE_PARTICLES 1
#define DEBUG_LVL 1
#define FIRST_ORDER_QUAD_POINT 1
#define RUNSPH 0
#endif
