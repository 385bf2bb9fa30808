// fluid-block test case in 2D, open boundaries (/root/reference/testcases/fluid-block: its parameter.h lacks CFL /
// PSI_1 / PSI_2 and says DIM 3 although generateIC.py writes 2D files; limiter settings of the
// Kelvin-Helmholtz long run, fixed time step as in its config.info)
#ifndef DEMONSTRATOR_PARAMETER_H
#define DEMONSTRATOR_PARAMETER_H
#define DIM 2
#define PERIODIC_BOUNDARIES 0
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
