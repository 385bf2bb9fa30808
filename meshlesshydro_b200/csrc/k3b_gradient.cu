// K3b -- locally-centred least-squares gradients, slope limiter and the CFL time step.
//
// Replaces (per particle, reference operation order kept):
//   psi-tilde weights of compPsijTilde  /root/reference/demonstrator/src/Particles.cpp:1230-1252 (ghosts :2400-2453)
//   Particles::gradient for rho, vx, vy, (vz), P   :1257-1270 (ghosts :2457-2502)
//   Particles::slopeLimiter             :1313-1444 (quirks Q1 abs, Q2 DBL_MIN "max", Q6 alpha rule)
//   Particles::compGlobalTimestep       :1446-1485 (quirk Q7: regular neighbours only) -- the global
//                                       min is a warp-shuffle + one atomicMin per block on the
//                                       ordered bit pattern of the (positive) double.
// The limiter reads only neighbours' VALUES (not their gradients), so gradient + limiter fuse into
// one kernel with two sweeps over the particle's list; ghost values are the parents' values
// (updateGhostState/updateGhostGradients, :2193-2222, need no copy here).
// Roofline: 18 (2D) / 30 (3D) algorithmic doubles per particle, ~100 FP64 ops per neighbour -> FP64-bound.
#include "mlh_internal.cuh"
#include <cfloat>

namespace {

template <int D>
__device__ __forceinline__ double dot_seq(const double *a, const double *b) { // Helper::dotProduct, Helper.cpp:20-26
    double res = 0.;
#pragma unroll
    for (int k = 0; k < D; ++k) res = __dadd_rn(res, __dmul_rn(a[k], b[k]));
    return res;
}

// neighbour record for list entry e: pk1 of particle j with the periodic image applied to the position
template <int D, bool PER>
__device__ __forceinline__ void load_neighbour(const Params &p, int e, double *nb) {
    const int j = e & MLH_NNL_IDX_MASK;
    load_packed<MLH_PK1(D)>(p.d.pk1 + (size_t)j * MLH_PK1(D), nb);
    if (PER && ((unsigned)e >> MLH_NNL_IDX_BITS) != 0u) { // image entries only (10 % of the kernel's instructions when unconditional)
#pragma unroll
        for (int k = 0; k < D; ++k) nb[k] = image_coord(nb[k], (e >> (MLH_NNL_IDX_BITS + 2 * k)) & 3, p.grid.bmin[k], p.grid.bmax[k]);
    }
}

// resident blocks per SM: the pipelined record (second register set) fits 128 registers in 2D; 3D needs 162 to stay
// spill-free (A/B r02c: 0.190 -> 0.181 ms at 61^3 with 3 blocks, 0.200 ms with 4 and spills; KH 1M 0.99 -> 0.83 ms)
#ifndef MLH_K3B_BLOCKS
#define MLH_K3B_BLOCKS(D) ((D) == 3 ? 3 : 4)
#endif
template <int D, bool PER>
__global__ void __launch_bounds__(128, MLH_K3B_BLOCKS(D)) k_gradient_limit(const Params p) {
    constexpr int NF = D + 2;
    constexpr int PK1 = MLH_PK1(D);
    int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    double dt_i = DBL_MAX;
    if (i < p.own_end) {
        // fields: rho, vx, vy, (vz), P  (order of MeshlessScheme.cpp:114-120 and Particles.cpp:1329-1335);
        // field f of a packed record sits at pk1 index fidx(f)
        int fslot[NF], fidx[NF];
        fslot[0] = 0; fidx[0] = 2 * D;
#pragma unroll
        for (int k = 0; k < D; ++k) { fslot[1 + k] = 1 + k; fidx[1 + k] = D + k; }
        fslot[NF - 1] = 4; fidx[NF - 1] = 2 * D + 1;

        double own[PK1];
        load_packed<PK1>(p.d.pk1 + (size_t)i * PK1, own);
        double xi[3], fi[NF], B[D * D];
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = own[k];
#pragma unroll
        for (int f = 0; f < NF; ++f) fi[f] = own[fidx[f]];
#pragma unroll
        for (int k = 0; k < D * D; ++k) B[k] = p.d.B[k][i];
        const double omg = own[2 * D + 3];
        const double inv_omg = __ddiv_rn(1., omg);
        const int nreg = p.d.noi[i], ntot = nreg + p.d.noig[i];

        double g[NF][D];
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int a = 0; a < D; ++a) g[f][a] = 0.;

        // ---- sweep 1: gradients + signal velocity (compGlobalTimestep, Particles.cpp:1446-1485: regular entries only,
        // quirk Q7; it needs the same |x_i - x_j| as the kernel weight, so it rides along here) ----
        double vSig = DBL_MIN; // quirk Q2
        const double ci = own[2 * D + 2];
        // list entries are fetched one visit ahead of the record gather (an L1 prefetch of the next record on top of
        // that cost more LSU issue than it hid: 0.199 -> 0.216 ms at 61^3, profiles/README.md r01t)
        int e_next = ntot > 0 ? p.d.nnl[i] : 0;
        // software pipeline: the record of visit s+1 is requested before visit s is computed (the gather latency was the
        // largest stall of both sweeps, profiles/r02b); its list entry was read a visit earlier
        double nbn[PK1];
        int e_next2 = ntot > 1 ? p.d.nnl[(size_t)p.ncap + i] : 0;
        if (ntot > 0) load_neighbour<D, PER>(p, e_next, nbn);
#pragma unroll 2
        for (int s = 0; s < ntot; ++s) {
            double nb[PK1];
#pragma unroll
            for (int k = 0; k < PK1; ++k) nb[k] = nbn[k];
            e_next = e_next2;
            if (s + 2 < ntot) e_next2 = p.d.nnl[(size_t)(s + 2) * p.ncap + i];
            if (s + 1 < ntot) load_neighbour<D, PER>(p, e_next, nbn);
            double d[3], sd[3];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                d[k] = __dsub_rn(nb[k], xi[k]);
                sd[k] = __dsub_rn(xi[k], nb[k]);
            }
            const double r = sqrt(dist_sqr_exact<D>(sd)); // Particles.cpp:1170-1175 / :2275-2280
            const double psij = mlh_div_known(cubic_spline(r, p), omg, inv_omg);
            double pt[D];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double acc = 0.;
#pragma unroll
                for (int b = 0; b < D; ++b) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(B[D * a + b], d[b]), psij));
                pt[a] = acc;
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double df = __dsub_rn(nb[fidx[f]], fi[f]);
#pragma unroll
                for (int a = 0; a < D; ++a) g[f][a] = __dadd_rn(g[f][a], __dmul_rn(df, pt[a]));
            }
            if (s < nreg) {
                // xij = x_i - x_j = sd, and sqrt(dotProduct(xij, xij)) is r bit for bit (same products, same sum order)
                double vij[D];
#pragma unroll
                for (int k = 0; k < D; ++k) vij[k] = __dsub_rn(own[D + k], nb[D + k]);
                double vijxij = __ddiv_rn(dot_seq<D>(vij, sd), r);
                vijxij = vijxij < 0. ? vijxij : 0.;
                const double vSig_i = __dsub_rn(__dadd_rn(ci, nb[2 * D + 2]), vijxij);
                vSig = vSig_i > vSig ? vSig_i : vSig;
            }
        }
        if (p.debug_capture) {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int a = 0; a < D; ++a) p.d.gpre[fslot[f] * 3 + a][i] = g[f][a];
        }

        // ---- sweep 2: slope limiter extrema (all entries) ----
        double maxNgb[NF], minNgb[NF], maxMid[NF], minMid[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            maxNgb[f] = DBL_MIN; // quirk Q2: smallest positive normal, not -inf
            minNgb[f] = DBL_MAX;
            maxMid[f] = DBL_MIN;
            minMid[f] = DBL_MAX;
        }
        e_next = ntot > 0 ? p.d.nnl[i] : 0;
        e_next2 = ntot > 1 ? p.d.nnl[(size_t)p.ncap + i] : 0;
        if (p.slope_limiting && ntot > 0) load_neighbour<D, PER>(p, e_next, nbn);
#pragma unroll 2
        for (int s = 0; p.slope_limiting && s < ntot; ++s) {
            double nb[PK1];
#pragma unroll
            for (int k = 0; k < PK1; ++k) nb[k] = nbn[k];
            e_next = e_next2;
            if (s + 2 < ntot) e_next2 = p.d.nnl[(size_t)(s + 2) * p.ncap + i];
            if (s + 1 < ntot) load_neighbour<D, PER>(p, e_next, nbn);
            {
                double xijxi[D];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    // FIRST_ORDER_QUAD_POINT: (x_i + x_j)/2 (`*.5` is exact like `/2.`), :1355-1356; else :1358-1359,1369
                    const double xij = p.quad_h4 ? __dadd_rn(xi[k], __dmul_rn(p.h4, __dsub_rn(nb[k], xi[k])))
                                                 : __dmul_rn(__dadd_rn(xi[k], nb[k]), .5);
                    xijxi[k] = __dsub_rn(xij, xi[k]);
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const double fj = nb[fidx[f]];
                    if (maxNgb[f] < fj) maxNgb[f] = fj;
                    if (minNgb[f] > fj) minNgb[f] = fj;
                    const double fij = __dadd_rn(fi[f], dot_seq<D>(g[f], xijxi));
                    if (maxMid[f] < fij) maxMid[f] = fij;
                    if (minMid[f] > fij) minMid[f] = fij;
                }
            }
        }
        if (p.slope_limiting) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double alphaMax = q1_abs(__ddiv_rn(__dsub_rn(maxNgb[f], fi[f]), __dsub_rn(maxMid[f], fi[f])), p.abs_mode);
                const double alphaMin = q1_abs(__ddiv_rn(__dsub_rn(fi[f], minNgb[f]), __dsub_rn(fi[f], minMid[f])), p.abs_mode);
                if (alphaMin <= alphaMax && __dmul_rn(p.beta, alphaMin) < 1.) {
#pragma unroll
                    for (int a = 0; a < D; ++a) g[f][a] = __dmul_rn(g[f][a], alphaMin);
                } else if (alphaMax <= alphaMin && __dmul_rn(p.beta, alphaMax) < 1.) {
#pragma unroll
                    for (int a = 0; a < D; ++a) g[f][a] = __dmul_rn(g[f][a], alphaMax);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int a = 0; a < D; ++a) p.d.g[fslot[f] * 3 + a][i] = g[f][a];
        {   // packed record for K4a: Binv, then gradients in W order rho, P, vx, vy(, vz)
            double rec[MLH_PK2(D)];
#pragma unroll
            for (int k = 0; k < D * D; ++k) rec[k] = B[k];
#pragma unroll
            for (int nu = 0; nu < NF; ++nu) {
                const int f = nu == 0 ? 0 : (nu == 1 ? NF - 1 : nu - 1);
#pragma unroll
                for (int a = 0; a < D; ++a) rec[D * D + nu * D + a] = g[f][a];
            }
            store_packed<MLH_PK2(D)>(p.d.pk2 + (size_t)i * MLH_PK2(D), rec);
        }
        dt_i = __ddiv_rn(__dmul_rn(p.cfl, p.h), vSig);
        if (!(dt_i < DBL_MAX)) dt_i = DBL_MAX; // `dt < dt_` never selects NaN / inf
    }
    // ---- block min -> global atomicMin (positive doubles order like their bit patterns) ----
    unsigned long long bits = (unsigned long long)__double_as_longlong(dt_i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other < bits ? other : bits;
    }
    __shared__ unsigned long long wmin[4];
    if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = wmin[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = wmin[w] < b ? wmin[w] : b;
        atomicMin(p.d.dt_bits, b);
    }
}

} // namespace

int mlh_launch_gradient(mlh_ctx *c) {
    Params &p = c->p;
    int n = p.own_end - p.own_begin;
    // the dt accumulator was set to DBL_MAX (Particles.cpp:1447) by k_density_matrix
    mlh_prof_begin(c, KID_GRADIENT);
    if (p.D == 2 && p.periodic)
        k_gradient_limit<2, true><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else if (p.D == 2)
        k_gradient_limit<2, false><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else if (p.periodic)
        k_gradient_limit<3, true><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else
        k_gradient_limit<3, false><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_GRADIENT);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
