/*
 * include/mlh_gpu.h -- the drop-in boundary: C ABI of the B200 MFV hot path.
 *
 * The reference (jammartin/meshlessHydro, CPU demonstrator) has no FFI layer: its hot path sits
 * behind the C++ methods of `Particles` / `Domain` / `Riemann`, called in a fixed order by
 * MeshlessScheme::run() (demonstrator/src/MeshlessScheme.cpp:39-253).  Each entry point below
 * replaces the reference methods named in its comment (file:line under
 * /root/reference/demonstrator/); host code (C++ mirror classes in meshlesshydro_b200/host/,
 * the ctypes binding in meshlesshydro_b200/capi.py, bench.py, tests/) calls ONLY these.
 * Plain pointers and sizes, no CUDA / torch types.  All arrays are host memory unless a
 * function says "device".  Every function returns MLH_OK (0) or a negative MLH_E_* code and
 * records a message retrievable with mlh_last_error().  There is no CPU fallback: without a
 * CUDA device mlh_create fails with MLH_E_NO_DEVICE.
 */
#ifndef MLH_GPU_H
#define MLH_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define MLH_ABI_VERSION 2

/* status codes */
#define MLH_OK 0
#define MLH_E_INVALID (-1)    /* bad argument / unsupported configuration */
#define MLH_E_NO_DEVICE (-2)  /* no CUDA device: the product never falls back to the CPU */
#define MLH_E_CUDA (-3)       /* CUDA runtime error (message in mlh_last_error) */
#define MLH_E_STATE (-4)      /* call order violated (e.g. advance before prepare) */
#define MLH_E_DEVICE_FLAG (-5)/* a kernel raised an error flag, see mlh_error_flags */
#define MLH_E_COMM (-6)       /* NCCL error */

/* device error flags (bit set); the host wrapper maps them to the reference's exit codes */
#define MLH_F_MAX_INTERACTIONS 1u /* Particles.cpp:348-352 exit(1) / :2249-2253 exit(3) */
#define MLH_F_OUT_OF_GRID 2u      /* particle outside the search grid (reference: UB)    */
#define MLH_F_NEG_GHOST_PRESSURE 4u /* Particles.cpp:1873-1881 exit(6) when DEBUG_LVL     */
#define MLH_F_HALO_OVERFLOW 8u    /* multi-GPU: halo / migration buffer too small         */
#define MLH_F_MIGRATION 16u       /* multi-GPU: a particle moved more than one cell layer */
#define MLH_F_VACUUM 32u          /* Riemann.cpp:113,126 "Vacuum state sampled"           */

/* quirk switches */
#define MLH_ABS_INT_TRUNC 0 /* Q1: unqualified abs() resolves to int abs(int) (g++/libstdc++) */
#define MLH_ABS_FABS 1      /* Q1: fabs (libc++ builds; the evident intent)                   */
#define MLH_Q13_ZERO_Z 0    /* Q13: xjxi[2] never written -> reads ~0 (observed)              */
#define MLH_Q13_GEOMETRIC 1 /* Q13: z_j - z_i                                                  */
#define MLH_Q3_REFERENCE 0  /* Q3: vz[i] used on the j side (Particles.cpp:1717,1719)         */
#define MLH_Q3_FIXED 1

typedef struct mlh_ctx mlh_ctx;

/*
 * Run configuration = the reference's compile-time parameter.h switches
 * (demonstrator/include/parameter.h:9-74) plus the run-time keys of config.info that the path
 * reads (MeshlessScheme.h:19-28).  Zero-initialise, then set fields.
 */
typedef struct {
    int dim;                  /* DIM (2|3)                                    parameter.h:9  */
    int periodic;             /* PERIODIC_BOUNDARIES (2D only, as the reference) parameter.h:12 */
    int max_interactions;     /* neighbour-list capacity per particle, regular + ghost entries
                                 (MAX_NUM_INTERACTIONS + MAX_NUM_GHOST_INTERACTIONS, :21,:25); 0 -> 128 */
    int slope_limiting;       /* SLOPE_LIMITING        parameter.h:28 */
    int pairwise_limiter;     /* PAIRWISE_LIMITER      parameter.h:34 */
    int meshless_finite_mass; /* MESHLESS_FINITE_MASS  parameter.h:39 */
    int move_particles;       /* MOVE_PARTICLES        parameter.h:45 */
    int abs_mode;             /* MLH_ABS_*  */
    int q13_mode;             /* MLH_Q13_*  */
    int q3_mode;              /* MLH_Q3_*   */
    int symmetric_seam;       /* 1: a periodic-seam pair is kept iff EITHER side passes the cutoff
                                 test (conservative; deviates from the reference's sets, quirk Q9) */
    int debug_capture;        /* 1: also store pre-limiter gradients and per-particle flux sums */
    int first_order_quad_point; /* FIRST_ORDER_QUAD_POINT parameter.h:55: 1 = face at the midpoint (every shipped parameter
                                 file), 0 = at x_i + kernelSize/4 (x_j - x_i) (Particles.cpp:1358-1359,1514-1537,1555-1563) */
    int reserved0;            /* keeps the doubles 8-byte aligned; must be 0 */
    double cfl;               /* CFL    parameter.h:18 */
    double beta;              /* BETA   parameter.h:31 */
    double psi1, psi2;        /* PSI_1, PSI_2 parameter.h:35-36 */
    double kernel_size;       /* kernelSize (config.info) */
    double gamma;             /* gamma      (config.info) */
    double box[6];            /* periodicBoxLimits [minX,minY(,minZ),maxX,maxY(,maxZ)] main.cpp:71-77 */
    int device;               /* CUDA device ordinal */
    int rank, nranks;         /* slab decomposition over GPUs of one node (nranks<=1: single GPU) */
    long capacity;            /* particle capacity of this rank incl. halo; 0 -> derived from N */
    long stage_bytes;         /* budget of the per-face staging buffer of the flux pass (it replaces the reference's
                                 per-slot WijL/WijR/Aij/Fij arrays, Particles.h:201-229); 0 -> 1/8 of the device memory.  Faces are
                                 processed in chunks that fit the budget. */
} mlh_config;

int mlh_abi_version(void);
void mlh_default_config(mlh_config *cfg); /* values of demonstrator/include/parameter.h, MFV mode */

/* Particles::Particles + MeshlessScheme ctor (Particles.cpp:67-149, MeshlessScheme.cpp:7-19) */
int mlh_create(const mlh_config *cfg, mlh_ctx **out);
int mlh_destroy(mlh_ctx *ctx); /* Particles::~Particles, Particles.cpp:151-225 */
const char *mlh_last_error(const mlh_ctx *ctx); /* ctx may be NULL: error of the last failed mlh_create */
unsigned mlh_error_flags(mlh_ctx *ctx);         /* synchronises; MLH_F_* bits raised so far */

/*
 * InitialDistribution::getAllParticles (InitialDistribution.cpp:32-60): host SoA -> HBM.
 * z/vz NULL in 2D.  `global_ids` NULL -> 0..N-1 (file order; decides the canonical face
 * orientation, quirk Q4).  With nranks>1 every rank passes the particles it owns.
 */
int mlh_upload(mlh_ctx *ctx, long N, const double *x, const double *y, const double *z,
               const double *vx, const double *vy, const double *vz, const double *m, const double *u,
               const int *global_ids);

/* ---- one time step = MeshlessScheme::run() loop body, MeshlessScheme.cpp:39-253 ---- */

/* K0+K1: Particles::getDomainLimits (Particles.cpp:228-267, non-periodic), Domain::createGrid
 * (Domain.cpp:9-54), Particles::assignParticlesAndCells (Particles.cpp:270-322): cell keys,
 * counting sort, reorder of the SoA state (within a cell ascending original index). */
int mlh_build_grid(mlh_ctx *ctx);
/* K2: Particles::gridNNS (:324-365), createGhostParticles (:2113-2191), ghostNNS (:2237-2260),
 * Domain::getNeighborCells (Domain.cpp:83-118). */
int mlh_neighbours(mlh_ctx *ctx);
/* K3: compDensity/compOmega (:1151-1184, :2262-2290), compPressure (:1272-1288), the matrix part of
 * compPsijTilde (:1186-1228, :2292-2381) incl. Helper::inverseMatrix (Helper.cpp:7-18),
 * updateGhostState (:2193-2206; a halo exchange when nranks>1). */
int mlh_density_matrix(mlh_ctx *ctx);
/* K3b: psi-tilde weights + gradient (:1230-1270, :2400-2502), slopeLimiter (:1313-1444),
 * compGlobalTimestep (:1446-1485), updateGhostGradients (:2208-2222; halo exchange when nranks>1). */
int mlh_gradients_limit(mlh_ctx *ctx);
/* dt of compGlobalTimestep (min over all ranks); synchronises the stream */
int mlh_timestep(mlh_ctx *ctx, double *dt_cfl);
/* K4+K5: compEffectiveFace (:1290-1311, :2504-2531), compRiemannStatesLR (:1488-1733, :2533-2683),
 * pairwiseLimiter (:1735-1785), solveRiemannProblems (:1787-1911) with Riemann::* (Riemann.cpp:7-229)
 * and the exact solver, collectFluxes (:1913-2011), updateStateAndPosition (:2013-2110). */
int mlh_flux_update(mlh_ctx *ctx, double dt);

/* phases 0-8 (grid .. limiter + CFL); afterwards the pre-update state, rho, P, gradients and noi
 * are what Particles::dump2file would write (MeshlessScheme.cpp:165-195). */
int mlh_prepare(mlh_ctx *ctx, double *dt_cfl);
/* phases 9-16 with the given dt */
int mlh_advance(mlh_ctx *ctx, double dt);
/* prepare + dt policy + advance without a host round trip for dt: dt_fixed>=0 -> that value
 * (ADAPTIVE_TIMESTEP 0); negative: the CFL dt, clipped to dt_max if dt_max>0 (dump-time clipping,
 * MeshlessScheme.cpp:94-101).  dt_used may be NULL (no synchronisation then). */
int mlh_step(mlh_ctx *ctx, double dt_fixed, double dt_max, double *dt_used);

/*
 * Stand-alone face solver = the reference's Riemann class for a batch of n faces (Riemann.h:19-26):
 * for each face `Riemann{WR, WL, vFrame, Aij, i}.exact(Fij, gamma)` with gamma / MESHLESS_FINITE_MASS of the context
 * (Riemann.cpp:7-229 incl. the exact solver behind RiemannSolver::solve, Riemann.cpp:93-94).  Same argument roles as
 * the class: WL = left state of the solver (the caller's WijR, quirk Q5), WR = right state; W = [rho, P, vx, vy(, vz)]
 * per face (n x (DIM+2)), vFrame and Aij n x DIM, Fij n x (DIM+2) = [mass, energy, px, py(, pz)].  The inputs are not
 * modified (the class rotates WR/WL in place; the host mirror does that itself).  Works on a context without particles.
 */
int mlh_riemann_faces(mlh_ctx *ctx, long n, const double *WR, const double *WL, const double *vFrame, const double *Aij,
                      double *Fij);

/* ---- results ---- */
/* current state in ORIGINAL particle order (of this rank's uploaded/owned ids; ids_out optional) */
int mlh_download_state(mlh_ctx *ctx, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                       double *m, double *u, int *ids_out);
/* what dump2file needs besides the state (Particles.cpp:2999-3006): rho, P, rhoGrad[N*DIM], noi;
 * valid after mlh_prepare; original order; any pointer may be NULL */
int mlh_download_diag(mlh_ctx *ctx, double *rho, double *P, double *rhoGrad, int *noi);
/* sumVolume, sumMass, sumEnergy, sumMomentumX/Y/Z (Particles.cpp:2830-2886); all ranks; out[6] */
int mlh_sums(mlh_ctx *ctx, double *out6);
long mlh_num_particles(mlh_ctx *ctx);      /* owned by this rank */
int mlh_grid_info(mlh_ctx *ctx, int *cells3, double *cell_size3, double *bounds6);

/*
 * Parity harness: copy one named per-particle quantity to the host in ORIGINAL order.
 * doubles: "x","y","z","vx","vy","vz","m","u","rho","P","omega","Binv"(N*D*D),
 *          "rhoGrad","vxGrad","vyGrad","vzGrad","PGrad"(N*D), "gradPre"((D+2)*N*D, debug_capture),
 *          "mF","eF"(N),"vF"(N*D) (debug_capture)
 * ints:    "cell","noi"(regular neighbours),"noiGhosts","sorted_index","num_faces"(1 value)
 * neighbour lists: "nnl"/"nnlGhosts" -> int[N*max_interactions], row i = ORIGINAL ids of the
 *          regular neighbours (resp. parent ids of the ghost neighbours) of particle i in list
 *          order; "nnlGhostCodes" the image code of each ghost entry (2 bits/dim: 1=+L, 2=-L).
 * per face (debug_capture, after mlh_advance/mlh_step; single GPU): "face_pairs" -> int[3*F]: original ids of the
 *          canonical endpoint a (the lower id, which solves the face in the reference, Particles.cpp:1841,1889) and of
 *          its partner b, and the image code of b in a's list; "face_rec" -> double[F*(4*DIM+4)]: the reference's
 *          per-slot WijR[a-slot] (state of a), WijL (state of b), vFrame, Aij (Particles.cpp:1290-1311,1488-1733);
 *          "face_F" -> double[F*(DIM+2)]: Fij of that slot (Particles.cpp:1787-1911).
 * "flux_symmetry" -> int[4] = Particles::checkFluxSymmetry (Particles.cpp:2888-2976) in structural form: slots with a
 *          face, violations (must be 0), faces used from both sides, faces only their owner uses (partner on another
 *          rank / one-sided seam pair, quirk Q9).  Valid after mlh_neighbours.
 * Returns the element count written, or <0.  dst NULL -> just the count.
 */
long mlh_debug_fetch(mlh_ctx *ctx, const char *field, void *dst, long dst_elems);

/* ---- measurement hooks (bench.py) ---- */
void *mlh_stream(mlh_ctx *ctx); /* the cudaStream_t every kernel of this context is launched on */
int mlh_synchronize(mlh_ctx *ctx);
/* per-kernel CUDA-event timing on that stream: enable, run steps, read accumulated ms + launch counts.
 * names: array of const char* (static storage) */
int mlh_profile_enable(mlh_ctx *ctx, int on);
int mlh_profile_read(mlh_ctx *ctx, int max_entries, const char **names, double *total_ms, long *launches);
long mlh_launch_count(mlh_ctx *ctx); /* kernels launched by this context so far */
/* ms between two points on the stream, measured with CUDA events recorded on mlh_stream */
int mlh_timer_start(mlh_ctx *ctx);
int mlh_timer_stop(mlh_ctx *ctx, double *ms);

/* pinned (page-locked) host memory for the arrays handed to mlh_upload / mlh_download_state: the
 * reference's Particles owns plain new[] arrays (Particles.cpp:67-149); a GPU-backed Particles allocates
 * them here so that the per-step host<->device copies run at full PCIe rate and asynchronously */
int mlh_host_alloc(unsigned long bytes, void **ptr);
int mlh_host_free(void *ptr);
/* measured FP64 (DFMA) peak of a device in TFLOP/s (2 flop per DFMA): the roofline denominator of the
 * Riemann/flux kernel, which MEASURED_PEAKS.json does not carry */
int mlh_measure_fp64_peak(int device, double *tflops);

/* ---- multi-GPU (slab decomposition along the slowest cell axis, NCCL halo exchange) ---- */
#define MLH_NCCL_ID_BYTES 128
int mlh_comm_unique_id(char *id128);                        /* rank 0: ncclGetUniqueId */
int mlh_comm_init(mlh_ctx *ctx, const char *id128);         /* all ranks: ncclCommInitRank */
/* host-side slab planning (no GPU needed): cell-layer range [lo,hi) owned by `rank` */
int mlh_slab_range(int n_layers, int nranks, int rank, int *lo, int *hi);

#ifdef __cplusplus
}
#endif
#endif /* MLH_GPU_H */
