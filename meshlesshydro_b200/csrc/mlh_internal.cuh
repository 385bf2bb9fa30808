// meshlesshydro_b200/csrc/mlh_internal.cuh -- shared by the kernels and the C-ABI layer (not public).
//
// Data layout in HBM (all FP64 SoA, one array per scalar component, 32-bit indices):
//   SRT set: particles sorted by search-grid cell (x fastest), ascending ORIGINAL index inside a
//            cell == the order of the reference's Cell::prtcls vectors (Particles.cpp:319).
//            Read by K2..K4.  pos, vel, m, u, id + derived rho, P, omega, cs, Binv(D*D), grad((D+2)*D).
//   CUR set: the state the next step starts from (output of K4/K5, input of K1), same order as
//            the SRT set of the step that produced it.
//   nnl:     neighbour lists, slot-major (entry of slot s of particle i at nnl[s*ncap + i]) so that
//            one-thread-per-particle kernels read them coalesced.  Slots [0,noi) are regular
//            neighbours in the reference's list order, slots [noi, noi+noig) periodic images
//            ("ghosts") ordered by ascending parent original index (== ascending ghost index of
//            Particles::ghostNNS).  Entry = sorted index j | image code << 26.
//   faces:   every pair (i,j) of the lists is ONE face, owned by exactly one of its endpoints (the one
//            with the lower ORIGINAL index; always self when the partner is a halo particle of
//            another rank or does not list the pair, quirk Q9).  fmap (slot-major like nnl) maps
//            each list slot to its face: (global face index << 2) | bit 1 = owned by this particle | bit 0 = this
//            endpoint adds -F; MLH_FMAP_SKIP = no face.  Between K2 and k_face_index an OWNED slot holds the
//            MLH_K2_* word below (rank among the owner's slots + where the partner keeps the pair); the owner then
//            numbers the face and writes the index into the partner's slot as well.  fa/fe = owner index and list
//            entry of face f; F = fluxes in canonical orientation, AoS MLH_FREC(D) doubles per face.
//   grp:     grp[c*ncap + i] = first slot of stencil cell c (reference order, Domain.cpp:83-118) in the list of i.
//            The list of j is ordered stencil cell by stencil cell and ascending inside a cell, so particle i of
//            cell C sits in j's list at grp[mirror(c)][j] + (number of particles of C below i that list j) -- the
//            second term is counted by the warp that searches cell C (k2_neighbours.cu): no reverse search.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../include/mlh_gpu.h"

#define MLH_MAX_D 3
// 3D: 10 doubles.  [MLH_PK1_3D = 12 pads the record to 96 bytes = three 256-bit loads (load_packed) instead of five
// 128-bit ones that straddle sectors: no gain (A/B r2q, Sedov 61^3 / 128^3: K3b 0.179 -> 0.178 / 1.531 -> 1.518 ms,
// K4a 0.261 -> 0.263 / 2.386 -> 2.411); MLH_FREC_3D = 8 (64-byte flux records) LOST: update 0.087 -> 0.097 ms.]
#ifndef MLH_PK1_3D
#define MLH_PK1_3D 10
#endif
#define MLH_PK1(D) ((D) == 3 ? MLH_PK1_3D : 2 * (D) + 4)
#define MLH_PK2(D) ((D) * (D) + ((D) + 2) * (D))
#define MLH_NNL_IDX_BITS 26
#define MLH_NNL_IDX_MASK ((1 << MLH_NNL_IDX_BITS) - 1)
#ifndef MLH_FREC_3D
#define MLH_FREC_3D 6
#endif
#define MLH_FREC(D) ((D) == 2 ? 4 : MLH_FREC_3D)      // doubles per face in the flux array (D+2 used)
#define MLH_FMAP_SKIP 0xFFFFFFFCu           // slot without a face (partner's list overflowed)
// K2 -> k_face_index word of an owned slot
#define MLH_K2_OWNED 2u
#define MLH_K2_RANK_SHIFT 2           // 10 bits: rank of the slot among the owner's owned slots
#define MLH_K2_RANK_MASK 0x3FFu
#define MLH_K2_R_SHIFT 12             // 5 bits: particles of the owner's cell below the owner that list the partner
#define MLH_K2_R_OVER (1u << 17)      // not counted (cell with more than 32 particles): the partner's list is searched
#define MLH_K2_SC_SHIFT 18            // 5 bits: stencil cell of the owner as the partner sees it
#define MLH_K2_NOPARTNER (1u << 23)   // the partner keeps no slot for the pair (other rank's halo / one-sided pair)
#define MLH_K2_GHOST (1u << 24)       // periodic-image slot: the partner's slot is found among its image entries

// constants of the restated exact Riemann solver (same expressions as oracle/riemann_exact.h:rs_init)
struct RsConsts {
    double gamma, gp1d2g, gm1d2g, gm1dgp1, tdgp1, tdgm1, gm1d2, tgdgm1, ginv;
    double sqrt_tdgp1; // sqrt(2/(gamma+1))
    int root_n;        // n if (gamma-1)/(2 gamma) == 1/n for n in {5 (gamma=5/3), 7 (gamma=7/5)}, else 0
};

struct Grid {
    int cells[3];   // global cells per dim (Domain::cellsX/Y/Z); cells[2]=1 in 2D
    int lcells[3];  // cells of the local grid: == cells except along the slab axis when nranks>1
    int slab_dim;   // D-1 (slowest-varying axis of the cell id, Particles.cpp:298-302)
    int layer0;     // global layer index of local layer 0 along slab_dim (may be -1); 0 when nranks==1
    int sliced;     // 1 when nranks>1 (local grid = owned layers + one halo layer each side)
    int ncells;     // local cell count
    double bmin[3], bmax[3], cell_size[3];
};

struct DevPtrs {
    // SRT set
    double *x[3], *v[3], *m, *u, *rho, *P, *omega, *cs;
    double *B[9];   // Binv row-major as used by the reference (Particles.cpp:1249)
    double *g[15];  // gradients: field f in {0 rho,1 vx,2 vy,3 vz,4 P}; component a -> g[f*3+a]
    int *id, *cell, *noi, *noig, *nnl;
    unsigned short *grp; // grp[c*ncap + i] = first slot of stencil cell c (reference order) in the list of i
    // faces (see header comment)
    unsigned *fmap;
    int *nown, *face_start, *face_scan_tmp, *fa, *fe;
    double *F;
    // packed (AoS) gather records of the SRT set -- what a neighbour visit needs, contiguous per particle so that a
    // gather costs 3 (pk1) + 6 (pk2) 32-byte sectors instead of one sector per scalar array:
    //   pk1[i*PK1 ..] = x[D], v[D], rho, P, cs, omega          (written by K3;  PK1 = 2D+4)
    //   pk2[i*PK2 ..] = Binv[D*D] row-major, grad[(D+2)][D] in W order rho,P,vx,vy(,vz)   (written by K3b; PK2 = D*D+(D+2)*D)
    double *pk1, *pk2;
    // CUR set
    double *cx[3], *cv[3], *cm, *cu;
    int *cid;
    // sort scratch
    int *ckey, *crank, *perm, *cell_count, *cell_start, *scan_tmp;
    // debug capture (may be null)
    double *gpre[15];
    double *flux[5]; // mF, eF, vF[3]
    double *dbg_face; // debug_capture: per face (AoS, 4D+4 doubles) WijR (canonical endpoint), WijL, vFrame, Aij as K4a formed them
    // reductions
    unsigned long long *dt_bits; // min CFL dt as ordered bit pattern
    double *bbox;                // [0..2] min, [3..5] max over i>=1 (quirk Q8) as ORDERED KEYS (dbl_key), [6..8] x[0], [9..11] x[1] as doubles
    Grid *grid;                  // the search grid the kernels of K1 / K2 / the halo exchange read (Domain::createGrid on
                                 // the device for non-periodic single-GPU runs, else a copy of Params::grid)
    double *sums;                // 6 doubles
    unsigned *flags;             // MLH_F_* bits
    unsigned *counters;          // [0] one-sided seam pairs, [3] longest neighbour list of this step (K2)
    double *dt_used;             // dt chosen on device by k_select_dt
};

struct Params {
    int D, periodic;
    int n;          // particles in the SRT set (owned + halo)
    int own_begin, own_end; // owned range inside the SRT set ([0,n) when nranks==1)
    int ncur;       // particles in the CUR set
    int ncap;       // capacity (stride of nnl)
    int max_ni;
    int fcap;       // face capacity (fa, fe, F)
    int slope_limiting, pairwise, mfm, move_particles, abs_mode, q13_mode, q3_mode, symmetric_seam, debug_capture;
    int quad_h4;    // FIRST_ORDER_QUAD_POINT 0: face at x_i + h4 (x_j - x_i), h4 = kernelSize/4 (a FACTOR, as in the reference)
    double h4;
    double h, hSqr, gamma, cfl, beta, psi1, psi2;
    double h2, sigma, sigma4; // cubic spline: h/2, normalisation, normalisation/4 (Particles.cpp:10-15)
    double inv_h2;            // RN(1 / h2), for mlh_div_known
    RsConsts rs;
    Grid grid;
    DevPtrs d;
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// Order-preserving 64-bit key of a (non-NaN) double: unsigned comparison of keys == comparison of the doubles, so a
// bounding box is reduced with native 64-bit atomicMin/atomicMax (and ncclMin/ncclMax on ncclUint64) instead of CAS loops.
__host__ __device__ __forceinline__ unsigned long long dbl_key(double v) {
    union { double d; unsigned long long u; } c;
    c.d = v;
    return (c.u >> 63) ? ~c.u : (c.u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_dbl(unsigned long long k) {
    union { double d; unsigned long long u; } c;
    c.u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return c.d;
}

// x*x + y*y (+ z*z) exactly as `pow(dx,2) + pow(dy,2); dSqr += pow(dz,2)` evaluates on the CPU
// (no FMA contraction): Particles.cpp:342-346.  Bit-exactness of the neighbour sets hangs on this.
template <int D>
__device__ __forceinline__ double dist_sqr_exact(const double *d) {
    double s = __dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1]));
    if (D == 3) s = __dadd_rn(s, __dmul_rn(d[2], d[2]));
    return s;
}

// near-correctly-rounded t^3 (error-compensated) standing in for glibc's pow(t, 3.) at Particles.cpp:20
__device__ __forceinline__ double cube_cr(double t) {
    double t2 = __dmul_rn(t, t);
    double e = __fma_rn(t, t, -t2);
    return __fma_rn(t2, t, __dmul_rn(e, t));
}

// a / b, correctly rounded, for a divisor whose correctly rounded reciprocal y = RN(1/b) is already known (a run
// constant like h/2, or omega_i for all neighbours of particle i): two FMA correction steps instead of the ~15
// instructions of the IEEE division.  After the first step q is within half an ulp (+ a rounding) of a/b, hence a
// faithful approximation, and Markstein's theorem (Handbook of Floating-Point Arithmetic, section 4.7: y = RN(1/b), q
// faithful, r = a - b q exact by FMA  =>  RN(q + r y) = RN(a/b)) makes the second step exact.  Operands here are far
// from the overflow / underflow ranges the theorem excludes; zero, infinity and NaN behave as in a division.
// CPU check of the same sequence against a / b: tests/test_div_known.py (12 M quotients per run; 7e8 once, 0 mismatches).
__device__ __forceinline__ double mlh_div_known(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    q = __fma_rn(__fma_rn(-b, q, a), y, q);
    return __fma_rn(__fma_rn(-b, q, a), y, q);
}

// Kernel::cubicSpline, Particles.cpp:7-24 (support radius h; h2 = h/2)
__device__ __forceinline__ double cubic_spline(double r, const Params &p) {
    double q = mlh_div_known(r, p.h2, p.inv_h2);
    if (q <= 1.) {
        // sigma*(1.-3./2.*q*q*(1.-q/2.))
        double a = __dmul_rn(__dmul_rn(1.5, q), q);
        double b = __dsub_rn(1., __dmul_rn(q, 0.5));
        return __dmul_rn(p.sigma, __dsub_rn(1., __dmul_rn(a, b)));
    } else if (q < 2.) {
        return __dmul_rn(p.sigma4, cube_cr(__dsub_rn(2., q)));
    }
    return 0.;
}

// periodic image of coordinate x with code c (1: parent near the low side, image above max;
// 2: parent near the high side, image below min) -- Particles.cpp:2123-2142 formulas
__device__ __forceinline__ double image_coord(double x, int c, double bmin, double bmax) {
    if (c == 1) return __dadd_rn(bmax, __dsub_rn(x, bmin));
    if (c == 2) return __dsub_rn(bmin, __dsub_rn(bmax, x));
    return x;
}
// does that image exist?  `x <= min + h` / else-if `max - h < x` (Particles.cpp:2123,2126)
__device__ __forceinline__ bool image_exists(double x, int c, double bmin, double bmax, double h) {
    bool low = x <= __dadd_rn(bmin, h);
    if (c == 1) return low;
    if (c == 2) return !low && (__dsub_rn(bmax, h) < x);
    return true;
}
__device__ __forceinline__ int reverse_code1(int c) { return c == 1 ? 2 : (c == 2 ? 1 : 0); }
__device__ __forceinline__ int reverse_code(int code) {
    return reverse_code1(code & 3) | (reverse_code1((code >> 2) & 3) << 2) | (reverse_code1((code >> 4) & 3) << 4);
}

// quirk Q1: `abs` at Particles.cpp:1416-1417,1748-1759
__device__ __forceinline__ double q1_abs(double v, int mode) {
    if (mode == MLH_ABS_FABS) return fabs(v);
    // x86-64 cvttsd2si: truncation toward zero; NaN and anything outside int32 give the "integer indefinite" INT_MIN.
    // cvt.rzi.s32.f64 saturates instead (NaN -> 0, v >= 2^31 -> INT_MAX, v <= -2^31 -> INT_MIN): patch the first two.
    int k = __double2int_rz(v);
    if (!(v < 2147483648.0)) k = INT_MIN;
    if (k < 0) k = (int)(0u - (unsigned)k);
    return (double)k;
}

// 16-byte vector loads/stores of a packed record (PKn is even and the arrays are 256-byte aligned)
// Packed records (pk1, pk2, F; bases 256-byte aligned by the carve).  A record of a multiple of four doubles is
// 32-byte aligned and moves with 256-bit accesses (sm_100: LDG/STG.E.ENL2.256): one L1 sector look-up per 32 bytes
// instead of two -- the gathers of K4a ran at 85 % of the L1 look-up rate (profiles/r2t_k_face_states_kh1000j_*).
#ifndef MLH_WIDE_LDST
#define MLH_WIDE_LDST 1
#endif
template <int N>
__device__ __forceinline__ void load_packed(const double *__restrict__ src, double *dst) {
    static_assert(N % 2 == 0, "packed records hold an even number of doubles");
    constexpr int NQ = (MLH_WIDE_LDST && N % 4 == 0) ? N / 4 : 0;
#pragma unroll
    for (int k = 0; k < NQ; ++k)
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(dst[4 * k]), "=d"(dst[4 * k + 1]), "=d"(dst[4 * k + 2]), "=d"(dst[4 * k + 3])
            : "l"(src + 4 * k));
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int k = 2 * NQ; k < N / 2; ++k) {
        const double2 v = __ldg(s2 + k);
        dst[2 * k] = v.x;
        dst[2 * k + 1] = v.y;
    }
}
template <int N>
__device__ __forceinline__ void store_packed(double *dst, const double *src) {
    static_assert(N % 2 == 0, "packed records hold an even number of doubles");
    constexpr int NQ = (MLH_WIDE_LDST && N % 4 == 0) ? N / 4 : 0;
#pragma unroll
    for (int k = 0; k < NQ; ++k)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                     :: "l"(dst + 4 * k), "d"(src[4 * k]), "d"(src[4 * k + 1]), "d"(src[4 * k + 2]), "d"(src[4 * k + 3])
                     : "memory");
    double2 *d2 = reinterpret_cast<double2 *>(dst);
#pragma unroll
    for (int k = 2 * NQ; k < N / 2; ++k) d2[k] = make_double2(src[2 * k], src[2 * k + 1]);
}

// displacement (neighbour - self) and distance for list entry e of particle i, the way the reference
// evaluates it for a regular neighbour (Particles.cpp:1170-1175) or a ghost (:2275-2280)
template <int D, bool PER>
__device__ __forceinline__ void neighbour_geometry(const Params &p, const double *xi, int e, double *d, double *r) {
    const int j = e & MLH_NNL_IDX_MASK;
    double s[3];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double xj = p.d.x[k][j]; // SoA on purpose: the neighbours of one stencil cell are consecutive j, i.e. 4 per sector
        if (PER && ((unsigned)e >> MLH_NNL_IDX_BITS) != 0u) { // few entries are images: the branch is warp-uniform almost always
            int ck = (e >> (MLH_NNL_IDX_BITS + 2 * k)) & 3;
            xj = image_coord(xj, ck, p.grid.bmin[k], p.grid.bmax[k]);
        }
        d[k] = __dsub_rn(xj, xi[k]);
        s[k] = __dsub_rn(xi[k], xj);
    }
    *r = sqrt(dist_sqr_exact<D>(s));
}
// the same in two halves, so that a caller can request the coordinates of the NEXT neighbour before it computes with
// the current one (software pipeline of the gather)
template <int D>
__device__ __forceinline__ void neighbour_position(const Params &p, int e, double *xj) {
    const int j = e & MLH_NNL_IDX_MASK;
#pragma unroll
    for (int k = 0; k < D; ++k) xj[k] = p.d.x[k][j];
}
template <int D, bool PER>
__device__ __forceinline__ void neighbour_geometry_from(const Params &p, const double *xi, int e, const double *xjraw, double *d,
                                                        double *r) {
    double s[3];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double xj = xjraw[k];
        if (PER && ((unsigned)e >> MLH_NNL_IDX_BITS) != 0u) { // few entries are images: the branch is warp-uniform almost always
            int ck = (e >> (MLH_NNL_IDX_BITS + 2 * k)) & 3;
            xj = image_coord(xj, ck, p.grid.bmin[k], p.grid.bmax[k]);
        }
        d[k] = __dsub_rn(xj, xi[k]);
        s[k] = __dsub_rn(xi[k], xj);
    }
    *r = sqrt(dist_sqr_exact<D>(s));
}

// Bounding box of Particles::getDomainLimits (Particles.cpp:228-267): per-thread running min / max (max already
// excludes original particle 0, quirk Q8; NaN never passes `x < mn` / `x > mx`) -> warp shuffle -> shared memory ->
// one native atomic per block and bound.  All threads of the block must call it.
template <int D>
__device__ __forceinline__ void mlh_bbox_block_reduce(const Params &p, double *mn, double *mx) {
    __shared__ double red[2 * D][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double a = __shfl_xor_sync(0xffffffffu, mn[k], o);
            const double b = __shfl_xor_sync(0xffffffffu, mx[k], o);
            mn[k] = a < mn[k] ? a : mn[k];
            mx[k] = b > mx[k] ? b : mx[k];
        }
        if (lane == 0) {
            red[k][w] = mn[k];
            red[D + k][w] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * D) {
        const bool is_max = threadIdx.x >= D;
        double v = red[threadIdx.x][0];
        for (int q = 1; q < nw; ++q) {
            const double o = red[threadIdx.x][q];
            v = is_max ? (o > v ? o : v) : (o < v ? o : v);
        }
        unsigned long long *keys = (unsigned long long *)p.d.bbox;
        if (is_max)
            atomicMax(keys + 3 + (threadIdx.x - D), dbl_key(v));
        else
            atomicMin(keys + threadIdx.x, dbl_key(v));
    }
}

// ---------------------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------------------
enum KernelId {
    KID_BBOX = 0, KID_KEY, KID_SCAN, KID_SCATTER, KID_CELLSORT, KID_GATHER, KID_NEIGHBOURS, KID_DENSITY,
    KID_GRADIENT, KID_FACE_INDEX, KID_SELECT_DT, KID_FACES, KID_FLUX_SETUP, KID_FLUX, KID_FLUX_FINISH, KID_UPDATE, KID_SUMS, KID_UNPERMUTE, KID_HALO, KID_COUNT
};

struct mlh_ctx {
    mlh_config cfg;
    Params p;
    cudaStream_t stream;
    long n_owned;        // particles owned by this rank
    long capacity;
    bool have_state;     // mlh_upload done
    bool bbox_valid;     // d.bbox describes the CUR set (reduced by the update kernel of the last step; non-periodic runs)
    bool grid_host_current; // Params::grid (host) equals *d.grid (device); false while the device builds the grid itself
    bool grid_mirror_pending; // an asynchronous copy of *d.grid into h_grid is in flight (ev_grid)
    Grid *h_grid;        // pinned
    cudaEvent_t ev_grid;
    int phase;           // 0 = CUR valid (start of step); 1..4 after grid/neighbours/density/gradients
    void *pool;          // single device allocation backing all arrays
    size_t pool_bytes;
    double *stage;       // face staging buffer of K4 (k4_flux.cu): record, P*, solver queue x stage_chunk
    int stage_chunk;     // faces per K4 chunk (multiple of 128)
    size_t stage_budget; // bytes the staging buffer may take (fixed at the first flux pass)
    int num_sms;
    double *dl_scratch;  // un-permutation staging of mlh_download_state / mlh_download_diag (8 x ncap doubles, lazily allocated)
    char *rf_buf;        // device scratch of mlh_riemann_faces (the reference's per-face Riemann class: kept between calls)
    size_t rf_bytes;
    int max_cells;       // allocated cell-array size
    // pinned host mirror for small readbacks
    double *h_small;     // pinned
    unsigned *h_flags;   // pinned
    char err[512];
    // profiling
    bool profiling;
    cudaEvent_t ev[2 * 64];
    int ev_kernel[64];
    int ev_used;
    double prof_ms[KID_COUNT];
    long prof_launches[KID_COUNT];
    long launches;
    cudaEvent_t timer[2];
    // multi-GPU
    void *nccl_comm;
    int layer_lo, layer_hi; // owned global cell layers along the slab axis
    int n_layers_global;
    double *halo_buf;       // exchange-1 send staging: [2 directions][2D+2 fields][halo_cap]
    int *halo_ids;          // [2][halo_cap]
    int *halo_counts;       // device: sent dn/up, received dn/up
    int *h_counts;          // pinned mirror (8 ints)
    int halo_cap;
    int lo_layer_end, hi_layer_begin; // owned bottom layer = [own_begin, lo_layer_end), top = [hi_layer_begin, own_end)
};

// kernel launchers (each .cu implements its stage)
int mlh_launch_sort(mlh_ctx *c);        // k1_sort.cu
// exclusive scan of in[0..n) into out[0..n], out[n] = total; tmp holds n/1024+2 ints (k1_sort.cu)
int mlh_exclusive_scan(mlh_ctx *c, const int *in, int *out, int *tmp, int n);
int mlh_launch_face_index(mlh_ctx *c);  // k4_flux.cu (after K2; builds the face list)
int mlh_launch_neighbours(mlh_ctx *c);  // k2_neighbours.cu
int mlh_launch_density(mlh_ctx *c);     // k3_density.cu
int mlh_launch_gradient(mlh_ctx *c);    // k3b_gradient.cu
int mlh_launch_flux(mlh_ctx *c, double dt_fixed, double dt_max); // k4_flux.cu
int mlh_launch_sums(mlh_ctx *c);        // k5_reduce.cu
int mlh_launch_bbox(mlh_ctx *c);        // k5_reduce.cu
int mlh_launch_bbox_q8_replay(mlh_ctx *c); // k5_reduce.cu (rare path of quirk Q8)
int mlh_launch_make_grid(mlh_ctx *c);      // k5_reduce.cu: Domain::createGrid on the device (+ the Q8 replay if needed)
// halo.cu (nranks > 1)
void mlh_comm_destroy(mlh_ctx *c);
int mlh_comm_bbox(mlh_ctx *c);
int mlh_halo_exchange_particles(mlh_ctx *c);
int mlh_halo_read_layout(mlh_ctx *c);
int mlh_halo_refresh(mlh_ctx *c, double *const *arrays, int narrays, int width);
int mlh_comm_min_dt(mlh_ctx *c);
int mlh_comm_sum(mlh_ctx *c, double *dev, int n);
int mlh_launch_unpermute_f64(mlh_ctx *c, const double *src, const int *ids, double *dst, int n, int comps, int stride); // k5_reduce.cu
int mlh_launch_unpermute_i32(mlh_ctx *c, const int *src, const int *ids, int *dst, int n);

void mlh_prof_begin(mlh_ctx *c, int kid);
void mlh_prof_end(mlh_ctx *c, int kid);

#define MLH_CUDA_CHECK(c, expr)                                                                    \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            snprintf((c)->err, sizeof((c)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,     \
                     cudaGetErrorString(_e));                                                      \
            return MLH_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

static inline int mlh_blocks(long n, int threads) { return (int)((n + threads - 1) / threads); }
