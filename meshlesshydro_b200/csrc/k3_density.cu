// K3 -- density / omega / volume partition and the least-squares gradient matrix, one gather pass.
//
// Replaces (per particle, in the reference's operation order):
//   Particles::compDensity + compOmega   /root/reference/demonstrator/src/Particles.cpp:1151-1184
//   ghost overloads                      :2262-2290  (omega += image terms AFTER the self term)
//   Particles::compPressure              :1272-1288
//   matrix part of compPsijTilde         :1186-1228 (ghosts :2292-2381): E = sum (xj-xi)(xj-xi)^T psi_j(x_i),
//                                        psi_j(x_i) = W(r_ij)/omega_i  (quirk Q10: omega of i)
//   Helper::inverseMatrix                Helper.cpp:7-18 (LAPACK dgetrf_/dgetri_, restated below)
// The per-slot psi-tilde weights the reference stores (Particles.h:204) are NOT materialised; the
// consumers (K3b, K4) recompute them from Binv and omega.
// Roofline: algorithmic bytes 7 (2D) / 11 (3D) doubles per particle; ~35 FP64 ops per neighbour
// visit (two visits) -> FP64-bound on B200.
#include "mlh_internal.cuh"
#include <cfloat>

// LAPACK dgetf2 (partial pivoting) + dtrti2 + unblocked dgetri on an n x n column-major matrix held
// in registers; same operation order as oracle/mfv_oracle.c:orc_inverse, no FMA contraction.
template <int n>
__device__ __forceinline__ void inverse_lu(double *A) {
#define AA(r, c) A[(r) + (c) * n]
    int ipiv[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
        int pv = j;
        double amax = fabs(AA(j, j));
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            double v = fabs(AA(i, j));
            if (v > amax) {
                amax = v;
                pv = i;
            }
        }
        ipiv[j] = pv;
        if (amax != 0.) {
#pragma unroll
            for (int r = j + 1; r < n; ++r)
                if (pv == r) {
#pragma unroll
                    for (int k = 0; k < n; ++k) {
                        double t = AA(j, k);
                        AA(j, k) = AA(r, k);
                        AA(r, k) = t;
                    }
                }
            double rcp = __ddiv_rn(1., AA(j, j));
#pragma unroll
            for (int i = j + 1; i < n; ++i) AA(i, j) = __dmul_rn(AA(i, j), rcp);
        }
#pragma unroll
        for (int k = j + 1; k < n; ++k)
#pragma unroll
            for (int i = j + 1; i < n; ++i) AA(i, k) = __dsub_rn(AA(i, k), __dmul_rn(AA(i, j), AA(j, k)));
    }
#pragma unroll
    for (int j = 0; j < n; ++j) {
        AA(j, j) = __ddiv_rn(1., AA(j, j));
        double ajj = -AA(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) {
            if (AA(k, j) != 0.) {
                double t = AA(k, j);
#pragma unroll
                for (int i = 0; i < k; ++i) AA(i, j) = __dadd_rn(AA(i, j), __dmul_rn(t, AA(i, k)));
                AA(k, j) = __dmul_rn(AA(k, j), AA(k, k));
            }
        }
#pragma unroll
        for (int i = 0; i < j; ++i) AA(i, j) = __dmul_rn(AA(i, j), ajj);
    }
#pragma unroll
    for (int j = n - 2; j >= 0; --j) {
        double work[n];
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            work[i] = AA(i, j);
            AA(i, j) = 0.;
        }
#pragma unroll
        for (int k = j + 1; k < n; ++k)
#pragma unroll
            for (int i = 0; i < n; ++i) AA(i, j) = __dsub_rn(AA(i, j), __dmul_rn(AA(i, k), work[k]));
    }
#pragma unroll
    for (int j = n - 2; j >= 0; --j) {
#pragma unroll
        for (int cc = j + 1; cc < n; ++cc)
            if (ipiv[j] == cc) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    double t = AA(i, j);
                    AA(i, j) = AA(i, cc);
                    AA(i, cc) = t;
                }
            }
    }
#undef AA
}

namespace {

template <int D, bool PER>
__global__ void __launch_bounds__(128) k_density_matrix(const Params p) {
    int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.d.dt_bits = 0x7FEFFFFFFFFFFFFFull; // dt_ = DBL_MAX, Particles.cpp:1447
    if (!PER && blockIdx.x == 0 && threadIdx.x < 3) {
        // bounding box of the positions this step will produce (reduced by k_flux_sum_update); same start values as
        // k_bbox_init (k5_reduce.cu): Particles.cpp:230-231, quirk Q2
        unsigned long long *keys = (unsigned long long *)p.d.bbox;
        keys[threadIdx.x] = dbl_key(DBL_MAX);
        keys[3 + threadIdx.x] = dbl_key(DBL_MIN);
        p.d.bbox[6 + threadIdx.x] = -DBL_MAX;
        p.d.bbox[9 + threadIdx.x] = -DBL_MAX;
    }
    if (i >= p.own_end) return;
    double xi[3];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = p.d.x[k][i];
    const int nreg = p.d.noi[i], ntot = nreg + p.d.noig[i];
    // omega: regular terms, then the self term, then the image terms.  (Keeping the kernel values of this sweep in
    // shared memory for the matrix sweep LOST, 0.100 -> 0.142 ms at 61^3: 32 KB per block is carved out of the L1 that
    // serves the neighbour gathers -- profiles/README.md r02a.)
    // list entries are read two visits ahead, the next neighbour's coordinates one visit ahead (software pipeline: each
    // visit is an entry load followed by a dependent gather)
    double omg = 0.;
    int e_next = nreg > 0 ? p.d.nnl[i] : 0, e_next2 = nreg > 1 ? p.d.nnl[(size_t)p.ncap + i] : 0;
    double xn[3] = {0., 0., 0.}; // coordinates of the next neighbour, requested one visit ahead
    if (nreg > 0) neighbour_position<D>(p, e_next, xn);
#pragma unroll 2
    for (int s = 0; s < nreg; ++s) {
        double d[3], r;
        const int e = e_next;
        const double xc[3] = {xn[0], xn[1], xn[2]};
        e_next = e_next2;
        if (s + 2 < nreg) e_next2 = p.d.nnl[(size_t)(s + 2) * p.ncap + i];
        if (s + 1 < nreg) neighbour_position<D>(p, e_next, xn);
        neighbour_geometry_from<D, false>(p, xi, e, xc, d, &r);
        omg = __dadd_rn(omg, cubic_spline(r, p));
    }
    omg = __dadd_rn(omg, cubic_spline(0., p));
    if (PER)
        for (int s = nreg; s < ntot; ++s) {
            double d[3], r;
            neighbour_geometry<D, true>(p, xi, p.d.nnl[(size_t)s * p.ncap + i], d, &r);
            omg = __dadd_rn(omg, cubic_spline(r, p));
        }
    const double rho = __dmul_rn(p.d.m[i], omg);
    const double P = __dmul_rn(__dmul_rn(__dsub_rn(p.gamma, 1.), rho), p.d.u[i]);
    p.d.omega[i] = omg;
    p.d.rho[i] = rho;
    p.d.P[i] = P;
    const double cs = sqrt(__ddiv_rn(__dmul_rn(p.gamma, P), rho)); // Particles.cpp:1451
    p.d.cs[i] = cs;
    {   // packed gather record for K3b / K4a
        double rec[MLH_PK1(D)];
#pragma unroll
        for (int k = 2 * D + 4; k < MLH_PK1(D); ++k) rec[k] = 0.; // padding (3D)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            rec[k] = xi[k];
            rec[D + k] = p.d.v[k][i];
        }
        rec[2 * D] = rho;
        rec[2 * D + 1] = P;
        rec[2 * D + 2] = cs;
        rec[2 * D + 3] = omg;
        store_packed<MLH_PK1(D)>(p.d.pk1 + (size_t)i * MLH_PK1(D), rec);
    }
    // E matrix
    const double inv_omg = __ddiv_rn(1., omg);
    double E[D * D];
#pragma unroll
    for (int k = 0; k < D * D; ++k) E[k] = 0.;
    e_next = ntot > 0 ? p.d.nnl[i] : 0;
    e_next2 = ntot > 1 ? p.d.nnl[(size_t)p.ncap + i] : 0;
    if (ntot > 0) neighbour_position<D>(p, e_next, xn);
#pragma unroll 2
    for (int s = 0; s < ntot; ++s) {
        double d[3], r;
        const int e = e_next;
        const double xc[3] = {xn[0], xn[1], xn[2]};
        e_next = e_next2;
        if (s + 2 < ntot) e_next2 = p.d.nnl[(size_t)(s + 2) * p.ncap + i];
        if (s + 1 < ntot) neighbour_position<D>(p, e_next, xn);
        neighbour_geometry_from<D, PER>(p, xi, e, xc, d, &r);
        double psij = mlh_div_known(cubic_spline(r, p), omg, inv_omg);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b) E[D * a + b] = __dadd_rn(E[D * a + b], __dmul_rn(__dmul_rn(d[a], d[b]), psij));
    }
    inverse_lu<D>(E);
#pragma unroll
    for (int k = 0; k < D * D; ++k) p.d.B[k][i] = E[k];
}

} // namespace

int mlh_launch_density(mlh_ctx *c) {
    Params &p = c->p;
    int n = p.own_end - p.own_begin;
    mlh_prof_begin(c, KID_DENSITY);
    if (p.D == 2 && p.periodic)
        k_density_matrix<2, true><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else if (p.D == 2)
        k_density_matrix<2, false><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else if (p.periodic)
        k_density_matrix<3, true><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    else
        k_density_matrix<3, false><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_DENSITY);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
