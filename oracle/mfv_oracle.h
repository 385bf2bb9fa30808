/*
 * oracle/mfv_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  The shipped GPU path (meshlesshydro_b200/) never links or calls it.
 *
 * Serial CPU restatement of one time step of the reference's MFV hot path
 * (MeshlessScheme::run body, /root/reference/demonstrator/src/MeshlessScheme.cpp:39-253),
 * following /root/reference/demonstrator/src/Particles.cpp, Domain.cpp, Riemann.cpp and
 * Helper.cpp function by function (each function in mfv_oracle.c cites its lines).
 * DIM, PERIODIC_BOUNDARIES and the other parameter.h switches are run-time fields here.
 *
 * Pinning: the reference ships no tests or golden vectors.  This
 * restatement is pinned against the reference's own sources compiled in this container
 * (oracle/_ref, built by oracle/ref_build/Makefile) by tests/test_oracle_vs_ref.py and
 * through the fixtures under tests/golden/ generated from oracle/_ref.  The Riemann
 * arithmetic itself is PARITY UNPINNED (third-party, un-vendored; see riemann_exact.h).
 */
#ifndef MLH_MFV_ORACLE_H
#define MLH_MFV_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int dim;            /* DIM: 2 or 3                                   (parameter.h:9)  */
    int periodic;       /* PERIODIC_BOUNDARIES (2D only, as the reference) (parameter.h:12) */
    int max_ni;         /* MAX_NUM_INTERACTIONS                           (parameter.h:21) */
    int max_gi;         /* MAX_NUM_GHOST_INTERACTIONS                     (parameter.h:25) */
    int slope_limiting; /* SLOPE_LIMITING                                 (parameter.h:28) */
    int pairwise;       /* PAIRWISE_LIMITER                               (parameter.h:34) */
    int mfm;            /* MESHLESS_FINITE_MASS                           (parameter.h:39) */
    int move_particles; /* MOVE_PARTICLES                                 (parameter.h:45) */
    int abs_mode;       /* quirk Q1: 0 = INT_TRUNC (int abs(int), g++/libstdc++), 1 = FABS */
    int q13_mode;       /* quirk Q13: 0 = ZERO_Z (xjxi[2] reads as 0), 1 = GEOMETRIC       */
    int q3_mode;        /* quirk Q3: 0 = as the reference (vz[i] on the j side), 1 = fixed */
    int quad_point_h4;  /* 1 = FIRST_ORDER_QUAD_POINT 0: quadrature point x_i + h/4 (x_j - x_i) (parameter.h:55,
                           Particles.cpp:1358-1359,1514-1537,1555-1563); 0 = the midpoint (all shipped parameter files) */
    double cfl;         /* CFL   (parameter.h:18) */
    double beta;        /* BETA  (parameter.h:31) */
    double psi1, psi2;  /* PSI_1, PSI_2 (parameter.h:35-36) */
    double h;           /* kernelSize (config.info) */
    double gamma;       /* gamma      (config.info) */
    double box[6];      /* periodic box [minX,minY(,minZ),maxX,maxY(,maxZ)] (main.cpp:71-77) */
} orc_config;

void *orc_create(const orc_config *cfg, int N, const double *x, const double *y, const double *z,
                 const double *vx, const double *vy, const double *vz, const double *m, const double *u);
void orc_destroy(void *ctx);
/* same contract as ref_step in oracle/ref_build/ref_driver.cpp */
double orc_step(void *ctx, double dtFixed, double dtMax, int stopAfter);
double orc_last_dt_cfl(void *ctx);
void orc_sums(void *ctx, double *out6);
void orc_grid(void *ctx, int *cells3, double *cellSize3, double *bounds6);
/* same names as ref_fetch, plus "Binv" (N*dim*dim, row-major as used at Particles.cpp:1249),
 * "one_sided" (int[N], 1 if the particle has a ghost pair whose reverse slot is missing, quirk Q9) */
long orc_fetch(void *ctx, const char *name, void *dst);
void orc_phase_seconds(void *ctx, double *out11);

/* stand-alone pieces exported for unit tests */
void orc_inverse(double *A, int n); /* Helper.cpp:7-18 (LAPACK dgetrf_+dgetri_ restated) */
double orc_cubic_spline(double r, double h, int dim); /* Particles.cpp:7-24 */
int orc_riemann(double gamma, double rhoL, double uL, double PL, double rhoR, double uR, double PR,
                double *sol3, int *iters); /* riemann_exact.h */
/* one face: Riemann.cpp:7-229.  WR/WL = [rho,P,vx,vy(,vz)] (modified in place like the reference) */
void orc_face_flux(int dim, int mfm, double gamma, double *WR, double *WL, const double *vFrame,
                   const double *Aij, double *Fij);

#ifdef __cplusplus
}
#endif
#endif
