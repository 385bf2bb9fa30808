/*
 * oracle/mfv_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.  See mfv_oracle.h.
 *
 * Every function names the reference lines it restates (paths relative to
 * /root/reference/demonstrator/).  The operation ORDER of the reference is kept (sum
 * order over neighbour slots, left-to-right products) so that the restatement agrees
 * with the reference-source build (oracle/_ref) to round-off of libm's pow/LAPACK only.
 */
#include "mfv_oracle.h"
#include "riemann_exact.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define NF_MAX 5 /* fields rho, P, vx, vy, vz  (W layout: Particles.cpp:1567-1578) */

typedef struct {
    orc_config cfg;
    int N, D;
    /* particle state (Particles.h:33-38) */
    double *x[3], *v[3], *m, *u, *rho, *P, *omega;
    int *cell;
    /* gradients: [field 0=rho,1=vx,2=vy,3=vz,4=P][N*D] */
    double *grad[5], *gradPre[5];
    double *Binv; /* N*D*D */
    /* per-slot (Particles.h:201-208) */
    int *nnl, *noi;
    double *psij, *Aij, *WijL, *WijR, *Fij, *vFrame;
    double *mF, *eF, *vF;
    /* ghosts (Particles.h:222-231) */
    int Ng;
    double *gx[3];
    int *gparent, *ghostMap;
    double *grho, *gP, *gomega, *gv[3], *ggrad[5];
    int *nnlG, *noiG;
    double *psijG, *AijG, *WijLG, *WijRG, *FijG, *vFrameG;
    int *one_sided;
    /* grid (Domain.h:43-69) */
    double bmin[3], bmax[3], cellSize[3];
    int cells[3], numCells;
    int *cellStart, *cellList; /* CSR stand-in for Cell::prtcls (ascending i per cell) */
    double dt, dtCfl;
    double phaseSec[16];
    int errFlags; /* bit0: MAX_NUM_INTERACTIONS exceeded, bit1: ghost overflow, bit2: negative ghost-face pressure */
} Orc;

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }
static int *ialloc(size_t n) { return (int *)calloc(n ? n : 1, sizeof(int)); }

/* ---------- Kernel::cubicSpline, Particles.cpp:7-24 (support radius = h) ---------- */
double orc_cubic_spline(double r, double h, int dim) {
    double h2 = h / 2.;
    double sigma;
    double q;
    if (dim == 2)
        sigma = 10. / (7. * M_PI * h2 * h2);
    else
        sigma = 1. / (M_PI * h2 * h2 * h2);
    q = r / h2;
    if (0. <= q && q <= 1.) {
        return sigma * (1. - 3. / 2. * q * q * (1. - q / 2.));
    } else if (1. < q && q < 2.) {
        return sigma / 4. * pow(2. - q, 3.);
    } else {
        return 0.;
    }
}

/* ---------- quirk Q1: unqualified abs() at Particles.cpp:1416-1417,1748-1759 ---------- */
static double q1_abs(double v, int mode) {
    int k;
    if (mode == 1) return fabs(v);
    /* `int abs(int)`: double->int as x86-64 cvttsd2si (NaN / out of range -> INT_MIN), then
     * two's-complement negate (INT_MIN stays INT_MIN), then back to double */
    if (!(v > -2147483649.0 && v < 2147483648.0))
        k = INT_MIN;
    else
        k = (int)v;
    if (k < 0) k = (int)(0u - (unsigned)k);
    return (double)k;
}

/* ---------- Helper::inverseMatrix, Helper.cpp:7-18: LAPACK dgetrf_ + dgetri_ on a
 * column-major n x n matrix, restated (dgetf2 with partial pivoting; dtrti2; dgetri unblocked) */
void orc_inverse(double *A, int n) {
    int ipiv[3];
    double work[3];
    int i, j, k;
#define AA(r, c) A[(r) + (c) * n]
    for (j = 0; j < n; ++j) { /* dgetf2 */
        int p = j;
        double amax = fabs(AA(j, j));
        for (i = j + 1; i < n; ++i)
            if (fabs(AA(i, j)) > amax) {
                amax = fabs(AA(i, j));
                p = i;
            }
        ipiv[j] = p;
        if (AA(p, j) != 0.) {
            if (p != j)
                for (k = 0; k < n; ++k) {
                    double t = AA(j, k);
                    AA(j, k) = AA(p, k);
                    AA(p, k) = t;
                }
            {
                double r = 1. / AA(j, j);
                for (i = j + 1; i < n; ++i) AA(i, j) *= r;
            }
        }
        for (k = j + 1; k < n; ++k)
            for (i = j + 1; i < n; ++i) AA(i, k) -= AA(i, j) * AA(j, k);
    }
    for (j = 0; j < n; ++j) { /* dtrti2: inverse of upper triangular U in place */
        double ajj;
        AA(j, j) = 1. / AA(j, j);
        ajj = -AA(j, j);
        /* x := U(0:j-1,0:j-1) * x  (dtrmv upper, no-trans, non-unit), x = A(0:j-1, j) */
        for (k = 0; k < j; ++k) {
            if (AA(k, j) != 0.) {
                double t = AA(k, j);
                for (i = 0; i < k; ++i) AA(i, j) += t * AA(i, k);
                AA(k, j) *= AA(k, k);
            }
        }
        for (i = 0; i < j; ++i) AA(i, j) *= ajj;
    }
    for (j = n - 2; j >= 0; --j) { /* dgetri unblocked: solve inv(A)*L = inv(U) */
        for (i = j + 1; i < n; ++i) {
            work[i] = AA(i, j);
            AA(i, j) = 0.;
        }
        for (k = j + 1; k < n; ++k) /* dgemv: A(:,j) -= A(:,k)*work[k] */
            for (i = 0; i < n; ++i) AA(i, j) -= AA(i, k) * work[k];
    }
    for (j = n - 2; j >= 0; --j) { /* undo the row interchanges as column swaps */
        int p = ipiv[j];
        if (p != j)
            for (i = 0; i < n; ++i) {
                double t = AA(i, j);
                AA(i, j) = AA(i, p);
                AA(i, p) = t;
            }
    }
#undef AA
}

int orc_riemann(double gamma, double rhoL, double uL, double PL, double rhoR, double uR, double PR,
                double *sol3, int *iters) {
    rs_consts c;
    rs_init(&c, gamma);
    return rs_solve(&c, rhoL, uL, PL, rhoR, uR, PR, &sol3[0], &sol3[1], &sol3[2], 0., iters);
}

/* ---------- Helper::dotProduct, Helper.cpp:20-26 ---------- */
static double dotp(const double *a, const double *b, int D) {
    double res = 0.;
    int k;
    for (k = 0; k < D; ++k) res += a[k] * b[k];
    return res;
}

/* Helper::rotationMatrix2D, Helper.cpp:39-45 */
static void rot2(const double *a, const double *b, double *L) {
    L[0] = a[0] * b[0] + a[1] * b[1];
    L[1] = -(a[0] * b[1] - a[1] * b[0]);
    L[2] = -L[1];
    L[3] = L[0];
}

/* Helper::crossProduct + rotationMatrix3D, Helper.cpp:28-37,48-77 (singular for cos = -1) */
static void rot3(const double *a, const double *b, double *L) {
    double v[3], cosAB, n;
    v[0] = a[1] * b[2] - a[2] * b[1];
    v[1] = a[2] * b[0] - a[0] * b[2];
    v[2] = a[0] * b[1] - a[1] * b[0];
    cosAB = dotp(a, b, 3);
    n = 1. / (1. + cosAB);
    L[0] = 1. - n * (v[2] * v[2] + v[1] * v[1]);
    L[1] = -v[2] + n * v[0] * v[1];
    L[2] = v[1] + n * v[0] * v[2];
    L[3] = v[2] + n * v[0] * v[1];
    L[4] = 1. - n * (v[2] * v[2] + v[0] * v[0]);
    L[5] = -v[0] + n * v[1] * v[2];
    L[6] = -v[1] + n * v[0] * v[2];
    L[7] = v[0] + n * v[1] * v[2];
    L[8] = 1. - n * (v[1] * v[1] + v[0] * v[0]);
}

/*
 * One face.  Riemann::Riemann (Riemann.cpp:7-81), Riemann::exact (:83-141),
 * rotateAndProjectFluxes2D/3D (:144-229).  Argument naming as INSIDE the class:
 * the caller passes (WijL, WijR) for (WR, WL) (Particles.cpp:1852), i.e. WR = neighbour j,
 * WL = particle i; the solver gets (left = WL, right = WR) (Riemann.cpp:93).
 */
void orc_face_flux(int D, int mfm, double gamma, double *WR, double *WL, const double *vFrame,
                   const double *Aij, double *Fij) {
    double AijNorm = sqrt(dotp(Aij, Aij, D));
    double hatA[3], unitX[3] = {1., 0., 0.}, Lam[9], LamInv[9];
    double rhoSol, PSol, vSol[3] = {0., 0., 0.}, vLab[3], vSolBuf[3];
    rs_consts c;
    int flagLR, k;
    for (k = 0; k < D; ++k) hatA[k] = 1. / AijNorm * Aij[k];
    if (D == 2) {
        double vBufR[2] = {WR[2], WR[3]}, vBufL[2] = {WL[2], WL[3]};
        rot2(hatA, unitX, Lam);
        WR[2] = Lam[0] * vBufR[0] + Lam[1] * vBufR[1];
        WR[3] = Lam[2] * vBufR[0] + Lam[3] * vBufR[1];
        WL[2] = Lam[0] * vBufL[0] + Lam[1] * vBufL[1];
        WL[3] = Lam[2] * vBufL[0] + Lam[3] * vBufL[1];
    } else {
        double vBufR[3] = {WR[2], WR[3], WR[4]}, vBufL[3] = {WL[2], WL[3], WL[4]};
        rot3(hatA, unitX, Lam);
        WR[2] = Lam[0] * vBufR[0] + Lam[1] * vBufR[1] + Lam[2] * vBufR[2];
        WR[3] = Lam[3] * vBufR[0] + Lam[4] * vBufR[1] + Lam[5] * vBufR[2];
        WR[4] = Lam[6] * vBufR[0] + Lam[7] * vBufR[1] + Lam[8] * vBufR[2];
        WL[2] = Lam[0] * vBufL[0] + Lam[1] * vBufL[1] + Lam[2] * vBufL[2];
        WL[3] = Lam[3] * vBufL[0] + Lam[4] * vBufL[1] + Lam[5] * vBufL[2];
        WL[4] = Lam[6] * vBufL[0] + Lam[7] * vBufL[1] + Lam[8] * vBufL[2];
    }
    rs_init(&c, gamma);
    flagLR = rs_solve(&c, WL[0], WL[2], WL[1], WR[0], WR[2], WR[1], &rhoSol, &vSol[0], &PSol, 0., (int *)0);
    if (flagLR == 1) {
        for (k = 1; k < D; ++k) vSol[k] = WR[2 + k];
    } else if (flagLR == -1) {
        for (k = 1; k < D; ++k) vSol[k] = WL[2 + k];
    } /* flag 0: the reference leaves vSol[1..] as it was (uninitialised member); here 0 */
    for (k = 0; k < D; ++k) vSolBuf[k] = vSol[k];
    if (D == 2) {
        rot2(unitX, hatA, LamInv);
        vSol[0] = LamInv[0] * vSolBuf[0] + LamInv[1] * vSolBuf[1];
        vSol[1] = LamInv[2] * vSolBuf[0] + LamInv[3] * vSolBuf[1];
        Fij[0] = Aij[0] * rhoSol * vSol[0] + Aij[1] * rhoSol * vSol[1];
        vLab[0] = vSol[0] + vFrame[0];
        vLab[1] = vSol[1] + vFrame[1];
        if (mfm) {
            vSol[0] = 0.;
            vSol[1] = 0.;
        }
        Fij[2] = Aij[0] * (rhoSol * vLab[0] * vSol[0] + PSol) + Aij[1] * rhoSol * vLab[0] * vSol[1];
        Fij[3] = Aij[0] * rhoSol * vLab[1] * vSol[0] + Aij[1] * (rhoSol * vLab[1] * vSol[1] + PSol);
        Fij[1] = Aij[0] * (vSol[0] * (PSol / (gamma - 1.) + rhoSol * .5 * dotp(vLab, vLab, 2)) + PSol * vLab[0]) +
                 Aij[1] * (vSol[1] * (PSol / (gamma - 1.) + rhoSol * .5 * dotp(vLab, vLab, 2)) + PSol * vLab[1]);
    } else {
        rot3(unitX, hatA, LamInv);
        vSol[0] = LamInv[0] * vSolBuf[0] + LamInv[1] * vSolBuf[1] + LamInv[2] * vSolBuf[2];
        vSol[1] = LamInv[3] * vSolBuf[0] + LamInv[4] * vSolBuf[1] + LamInv[5] * vSolBuf[2];
        vSol[2] = LamInv[6] * vSolBuf[0] + LamInv[7] * vSolBuf[1] + LamInv[8] * vSolBuf[2];
        Fij[0] = Aij[0] * rhoSol * vSol[0] + Aij[1] * rhoSol * vSol[1] + Aij[2] * rhoSol * vSol[2];
        vLab[0] = vSol[0] + vFrame[0];
        vLab[1] = vSol[1] + vFrame[1];
        vLab[2] = vSol[2] + vFrame[2];
        if (mfm) {
            vSol[0] = 0.;
            vSol[1] = 0.;
            vSol[2] = 0.;
        }
        Fij[2] = Aij[0] * (rhoSol * vLab[0] * vSol[0] + PSol) + Aij[1] * rhoSol * vLab[0] * vSol[1] +
                 Aij[2] * rhoSol * vLab[0] * vSol[2];
        Fij[3] = Aij[0] * rhoSol * vLab[1] * vSol[0] + Aij[1] * (rhoSol * vLab[1] * vSol[1] + PSol) +
                 Aij[2] * rhoSol * vLab[1] * vSol[2];
        Fij[4] = Aij[0] * rhoSol * vLab[2] * vSol[0] + Aij[1] * rhoSol * vLab[2] * vSol[1] +
                 Aij[2] * (rhoSol * vLab[2] * vSol[2] + PSol);
        Fij[1] = Aij[0] * (vSol[0] * (PSol / (gamma - 1.) + rhoSol * .5 * dotp(vLab, vLab, 3)) + PSol * vLab[0]) +
                 Aij[1] * (vSol[1] * (PSol / (gamma - 1.) + rhoSol * .5 * dotp(vLab, vLab, 3)) + PSol * vLab[1]) +
                 Aij[2] * (vSol[2] * (PSol / (gamma - 1.) + rhoSol * .5 * dotp(vLab, vLab, 3)) + PSol * vLab[2]);
    }
}

/* ---------- Domain::createGrid, Domain.cpp:9-54 ---------- */
static void create_grid(Orc *o) {
    int k;
    o->numCells = 1;
    for (k = 0; k < o->D; ++k) {
        o->cells[k] = (int)floor((o->bmax[k] - o->bmin[k]) / o->cfg.h);
        o->cellSize[k] = (o->bmax[k] - o->bmin[k]) / (double)o->cells[k];
        o->numCells *= o->cells[k];
    }
    if (o->D == 2) {
        o->cells[2] = 1;
        o->cellSize[2] = 0.;
    }
    free(o->cellStart);
    o->cellStart = ialloc((size_t)o->numCells + 1);
}

/* ---------- Particles::getDomainLimits, Particles.cpp:228-267 (quirks Q2, Q8) ---------- */
static void domain_limits(Orc *o) {
    int i, k;
    for (k = 0; k < o->D; ++k) {
        double mn = DBL_MAX, mx = DBL_MIN;
        for (i = 0; i < o->N; ++i) {
            if (o->x[k][i] < mn) {
                mn = o->x[k][i];
            } else if (o->x[k][i] > mx) {
                mx = o->x[k][i];
            }
        }
        o->bmin[k] = mn;
        o->bmax[k] = mx;
    }
}

/* ---------- Particles::assignParticlesAndCells, Particles.cpp:270-322 ---------- */
static void assign_cells(Orc *o) {
    int i, k, c;
    memset(o->cellStart, 0, sizeof(int) * ((size_t)o->numCells + 1));
    for (i = 0; i < o->N; ++i) {
        int f[3] = {0, 0, 0};
        for (k = 0; k < o->D; ++k) {
            f[k] = (int)floor((o->x[k][i] - o->bmin[k]) / o->cellSize[k]);
            if (f[k] == o->cells[k]) f[k] -= 1;
        }
        c = f[0] + f[1] * o->cells[0];
        if (o->D == 3) c += f[2] * o->cells[0] * o->cells[1];
        o->cell[i] = c;
        if (c >= 0 && c < o->numCells) o->cellStart[c + 1]++;
    }
    for (c = 0; c < o->numCells; ++c) o->cellStart[c + 1] += o->cellStart[c];
    {
        int *fill = ialloc((size_t)o->numCells);
        for (i = 0; i < o->N; ++i) { /* ascending i within a cell == push_back order (:319) */
            c = o->cell[i];
            if (c >= 0 && c < o->numCells) o->cellList[o->cellStart[c] + fill[c]++] = i;
        }
        free(fill);
    }
}

/* ---------- Domain::getNeighborCells, Domain.cpp:83-118 (x outer, y, z inner; -1 outside) ---------- */
static int neighbor_cells(const Orc *o, int iCell, int *out) {
    int iX, iY, iZ = 0, k, l, m, n = 0;
    if (o->D == 2) {
        iX = iCell % o->cells[0];
        iY = iCell / o->cells[0];
        for (k = iX - 1; k <= iX + 1; ++k)
            for (l = iY - 1; l <= iY + 1; ++l) {
                if (k < 0 || k >= o->cells[0])
                    out[n] = -1;
                else if (l < 0 || l >= o->cells[1])
                    out[n] = -1;
                else
                    out[n] = k + l * o->cells[0];
                ++n;
            }
    } else {
        iX = iCell % o->cells[0];
        iY = (iCell / o->cells[0]) % o->cells[1];
        iZ = iCell / (o->cells[0] * o->cells[1]);
        for (k = iX - 1; k <= iX + 1; ++k)
            for (l = iY - 1; l <= iY + 1; ++l)
                for (m = iZ - 1; m <= iZ + 1; ++m) {
                    if (k < 0 || k >= o->cells[0])
                        out[n] = -1;
                    else if (l < 0 || l >= o->cells[1])
                        out[n] = -1;
                    else if (m < 0 || m >= o->cells[2])
                        out[n] = -1;
                    else
                        out[n] = k + l * o->cells[0] + m * o->cells[0] * o->cells[1];
                    ++n;
                }
    }
    return n;
}

/* ---------- Particles::gridNNS, Particles.cpp:324-365 ---------- */
static void grid_nns(Orc *o) {
    const double hSqr = o->cfg.h * o->cfg.h;
    int i, s, q, cells[27];
    for (i = 0; i < o->N; ++i) {
        int ns = neighbor_cells(o, o->cell[i], cells);
        int noiBuf = 0;
        for (s = 0; s < ns; ++s) {
            if (cells[s] < 0) continue;
            for (q = o->cellStart[cells[s]]; q < o->cellStart[cells[s] + 1]; ++q) {
                int ip = o->cellList[q];
                if (ip != i) {
                    double dx = o->x[0][ip] - o->x[0][i], dy = o->x[1][ip] - o->x[1][i];
                    double dSqr = dx * dx + dy * dy; /* pow(.,2) == exact square */
                    if (o->D == 3) {
                        double dz = o->x[2][ip] - o->x[2][i];
                        dSqr += dz * dz;
                    }
                    if (dSqr < hSqr) {
                        if (noiBuf >= o->cfg.max_ni) {
                            o->errFlags |= 1; /* reference: exit(1) */
                            continue;
                        }
                        o->nnl[noiBuf + (size_t)i * o->cfg.max_ni] = ip;
                        ++noiBuf;
                    }
                }
            }
        }
        o->noi[i] = noiBuf;
    }
}

/* ---------- Particles::createGhostParticles, Particles.cpp:2113-2191 (2D only) ---------- */
static void create_ghosts(Orc *o) {
    const double h = o->cfg.h;
    const double minX = o->bmin[0], maxX = o->bmax[0], minY = o->bmin[1], maxY = o->bmax[1];
    int i, g = 0;
    for (i = 0; i < o->N; ++i) {
        int fX = 0, fY = 0;
        const double x = o->x[0][i], y = o->x[1][i];
        o->ghostMap[i * 3] = o->ghostMap[i * 3 + 1] = o->ghostMap[i * 3 + 2] = -1;
        if (x <= minX + h) {
            o->gx[0][g] = maxX + (x - minX);
            fX = 1;
        } else if (maxX - h < x) {
            o->gx[0][g] = minX - (maxX - x);
            fX = 1;
        } else {
            o->gx[0][g] = x;
        }
        if (y <= minY + h) {
            o->gx[1][g] = maxY + (y - minY);
            fY = 1;
        } else if (maxY - h < y) {
            o->gx[1][g] = minY - (maxY - y);
            fY = 1;
        } else {
            o->gx[1][g] = y;
        }
        if (fX || fY) {
            o->ghostMap[i * 3] = g;
            o->gparent[g] = i;
            ++g;
        }
        if (fX && fY) {
            o->gx[0][g] = x;
            if (y <= minY + h)
                o->gx[1][g] = maxY + (y - minY);
            else if (maxY - h < y)
                o->gx[1][g] = minY - (maxY - y);
            o->ghostMap[i * 3 + 1] = g;
            o->gparent[g] = i;
            ++g;
            if (x <= minX + h)
                o->gx[0][g] = maxX + (x - minX);
            else if (maxX - h < x)
                o->gx[0][g] = minX - (maxX - x);
            o->gx[1][g] = y;
            o->ghostMap[i * 3 + 2] = g;
            o->gparent[g] = i;
            ++g;
        }
    }
    o->Ng = g;
}

/* ---------- Particles::ghostNNS, Particles.cpp:2237-2260 (brute force, ascending ghost index) ---------- */
static void ghost_nns(Orc *o) {
    const double hSqr = o->cfg.h * o->cfg.h;
    int i, g;
    for (i = 0; i < o->N; ++i) {
        int n = 0;
        for (g = 0; g < o->Ng; ++g) {
            double dx = o->gx[0][g] - o->x[0][i], dy = o->gx[1][g] - o->x[1][i];
            double dSqr = dx * dx + dy * dy;
            if (dSqr < hSqr) {
                if (n >= o->cfg.max_gi) {
                    o->errFlags |= 2; /* reference: exit(3) */
                    continue;
                }
                o->nnlG[n + (size_t)i * o->cfg.max_gi] = g;
                ++n;
            }
        }
        o->noiG[i] = n;
    }
}

static double dist_ij(const Orc *o, int i, int j) { /* Particles.cpp:1170-1175 */
    double dx = o->x[0][i] - o->x[0][j], dy = o->x[1][i] - o->x[1][j];
    double dSqr = dx * dx + dy * dy;
    if (o->D == 3) {
        double dz = o->x[2][i] - o->x[2][j];
        dSqr += dz * dz;
    }
    return sqrt(dSqr);
}
static double dist_ig(const Orc *o, int i, int g) { /* Particles.cpp:2275-2280 */
    double dx = o->x[0][i] - o->gx[0][g], dy = o->x[1][i] - o->gx[1][g];
    return sqrt(dx * dx + dy * dy);
}

/* ---------- compDensity/compOmega (Particles.cpp:1151-1184; ghosts :2262-2290), compPressure (:1272-1288) ---------- */
static void density_pressure(Orc *o) {
    const double h = o->cfg.h;
    int i, j;
    for (i = 0; i < o->N; ++i) {
        double omg = 0.;
        for (j = 0; j < o->noi[i]; ++j) omg += orc_cubic_spline(dist_ij(o, i, o->nnl[j + (size_t)i * o->cfg.max_ni]), h, o->D);
        o->omega[i] = omg + orc_cubic_spline(0., h, o->D);
        o->rho[i] = o->m[i] * o->omega[i];
    }
    if (o->cfg.periodic)
        for (i = 0; i < o->N; ++i) {
            double omg = o->omega[i];
            for (j = 0; j < o->noiG[i]; ++j) omg += orc_cubic_spline(dist_ig(o, i, o->nnlG[j + (size_t)i * o->cfg.max_gi]), h, o->D);
            o->omega[i] = omg;
            o->rho[i] = o->m[i] * o->omega[i];
        }
    for (i = 0; i < o->N; ++i) o->P[i] = (o->cfg.gamma - 1.) * o->rho[i] * o->u[i];
}

/* ---------- Particles::compGlobalTimestep, Particles.cpp:1446-1485 (quirks Q2, Q7) ---------- */
static double global_timestep(Orc *o) {
    const double gamma = o->cfg.gamma;
    double dt_ = DBL_MAX;
    int i, jn, k;
    for (i = 0; i < o->N; ++i) {
        double vSig = DBL_MIN;
        double ci = sqrt(gamma * o->P[i] / o->rho[i]);
        double dt;
        for (jn = 0; jn < o->noi[i]; ++jn) {
            int j = o->nnl[(size_t)i * o->cfg.max_ni + jn];
            double cj = sqrt(gamma * o->P[j] / o->rho[j]);
            double xij[3], vij[3], vijxij, vSig_i;
            for (k = 0; k < o->D; ++k) {
                xij[k] = o->x[k][i] - o->x[k][j];
                vij[k] = o->v[k][i] - o->v[k][j];
            }
            vijxij = dotp(vij, xij, o->D) / sqrt(dotp(xij, xij, o->D));
            vijxij = vijxij < 0. ? vijxij : 0.;
            vSig_i = ci + cj - vijxij;
            vSig = vSig_i > vSig ? vSig_i : vSig;
        }
        dt = o->cfg.cfl * o->cfg.h / vSig;
        dt_ = dt < dt_ ? dt : dt_;
    }
    return dt_;
}

/* ---------- updateGhostState (:2193-2206) / updateGhostGradients (:2208-2222) ---------- */
static void update_ghost_state(Orc *o) {
    int i, k;
    for (i = 0; i < o->N * 3; ++i)
        if (o->ghostMap[i] >= 0) {
            int g = o->ghostMap[i], p = i / 3;
            o->grho[g] = o->rho[p];
            o->gP[g] = o->P[p];
            o->gomega[g] = o->omega[p];
            for (k = 0; k < o->D; ++k) o->gv[k][g] = o->v[k][p];
        }
}
static void update_ghost_gradients(Orc *o) {
    int i, k, f;
    for (i = 0; i < o->N * 3; ++i)
        if (o->ghostMap[i] >= 0) {
            int g = o->ghostMap[i], p = i / 3;
            for (f = 0; f < 5; ++f) {
                if (f == 3 && o->D == 2) continue;
                for (k = 0; k < o->D; ++k) o->ggrad[f][(size_t)g * o->D + k] = o->grad[f][(size_t)p * o->D + k];
            }
        }
}

/* ---------- Particles::compPsijTilde, Particles.cpp:1186-1254 (ghost overload :2292-2455) ---------- */
static void psij_tilde(Orc *o) {
    const int D = o->D;
    const double h = o->cfg.h;
    int i, j, a, b;
    double B[9], xi[3], xj[3];
    for (i = 0; i < o->N; ++i) {
        for (a = 0; a < D * D; ++a) B[a] = 0.;
        for (a = 0; a < D; ++a) xi[a] = o->x[a][i];
        for (j = 0; j < o->noi[i]; ++j) {
            int ip = o->nnl[j + (size_t)i * o->cfg.max_ni];
            double psij_xi = orc_cubic_spline(dist_ij(o, i, ip), h, D) / o->omega[i];
            for (a = 0; a < D; ++a) xj[a] = o->x[a][ip];
            for (a = 0; a < D; ++a)
                for (b = 0; b < D; ++b) B[D * a + b] += (xj[a] - xi[a]) * (xj[b] - xi[b]) * psij_xi;
        }
        if (o->cfg.periodic)
            for (j = 0; j < o->noiG[i]; ++j) {
                int g = o->nnlG[j + (size_t)i * o->cfg.max_gi];
                double psij_xi = orc_cubic_spline(dist_ig(o, i, g), h, D) / o->omega[i];
                for (a = 0; a < D; ++a) xj[a] = o->gx[a][g];
                for (a = 0; a < D; ++a)
                    for (b = 0; b < D; ++b) B[D * a + b] += (xj[a] - xi[a]) * (xj[b] - xi[b]) * psij_xi;
            }
        orc_inverse(B, D);
        for (a = 0; a < D * D; ++a) o->Binv[(size_t)i * D * D + a] = B[a];
        for (j = 0; j < o->noi[i]; ++j) {
            int ip = o->nnl[j + (size_t)i * o->cfg.max_ni];
            double psij_xi = orc_cubic_spline(dist_ij(o, i, ip), h, D) / o->omega[i];
            double *pt = &o->psij[((size_t)i * o->cfg.max_ni + j) * D];
            for (a = 0; a < D; ++a) xj[a] = o->x[a][ip];
            for (a = 0; a < D; ++a) {
                pt[a] = 0.;
                for (b = 0; b < D; ++b) pt[a] += B[D * a + b] * (xj[b] - xi[b]) * psij_xi;
            }
        }
        if (o->cfg.periodic)
            for (j = 0; j < o->noiG[i]; ++j) {
                int g = o->nnlG[j + (size_t)i * o->cfg.max_gi];
                double psij_xi = orc_cubic_spline(dist_ig(o, i, g), h, D) / o->omega[i];
                double *pt = &o->psijG[((size_t)i * o->cfg.max_gi + j) * D];
                for (a = 0; a < D; ++a) xj[a] = o->gx[a][g];
                for (a = 0; a < D; ++a) {
                    pt[a] = 0.;
                    for (b = 0; b < D; ++b) pt[a] += B[D * a + b] * (xj[b] - xi[b]) * psij_xi;
                }
            }
    }
}

/* ---------- Particles::gradient, Particles.cpp:1257-1270 (ghost overload :2457-2502) ---------- */
static void gradient(Orc *o, const double *f, double *grad, const double *fGhost) {
    const int D = o->D;
    int i, j, a;
    for (i = 0; i < o->N; ++i) {
        double *g = &grad[(size_t)i * D];
        for (a = 0; a < D; ++a) g[a] = 0;
        for (j = 0; j < o->noi[i]; ++j)
            for (a = 0; a < D; ++a)
                g[a] += (f[o->nnl[j + (size_t)i * o->cfg.max_ni]] - f[i]) * o->psij[((size_t)i * o->cfg.max_ni + j) * D + a];
        if (o->cfg.periodic)
            for (j = 0; j < o->noiG[i]; ++j)
                for (a = 0; a < D; ++a)
                    g[a] += (fGhost[o->nnlG[j + (size_t)i * o->cfg.max_gi]] - f[i]) *
                            o->psijG[((size_t)i * o->cfg.max_gi + j) * D + a];
    }
}

/* ---------- Particles::slopeLimiter (per field), Particles.cpp:1338-1444 (quirks Q1, Q2, Q6) ---------- */
static void slope_limiter(Orc *o, const double *f, double *grad, const double *fGhost) {
    const int D = o->D;
    int i, jn, a;
    for (i = 0; i < o->N; ++i) {
        double psiMaxNgb = DBL_MIN, psiMinNgb = DBL_MAX, psiMaxMid = DBL_MIN, psiMinMid = DBL_MAX;
        double xij[3], xijxi[3], alphaMax, alphaMin, fij;
        double *g = &grad[(size_t)i * D];
        for (jn = 0; jn < o->noi[i]; ++jn) {
            int j = o->nnl[(size_t)i * o->cfg.max_ni + jn];
            for (a = 0; a < D; ++a) {
                if (o->cfg.quad_point_h4) /* :1358-1359,1369 */
                    xij[a] = o->x[a][i] + o->cfg.h / 4. * (o->x[a][j] - o->x[a][i]);
                else
                    xij[a] = (o->x[a][i] + o->x[a][j]) / 2.;
                xijxi[a] = xij[a] - o->x[a][i];
            }
            if (psiMaxNgb < f[j]) psiMaxNgb = f[j];
            if (psiMinNgb > f[j]) psiMinNgb = f[j];
            fij = f[i] + dotp(g, xijxi, D);
            if (psiMaxMid < fij) psiMaxMid = fij;
            if (psiMinMid > fij) psiMinMid = fij;
        }
        if (o->cfg.periodic)
            for (jn = 0; jn < o->noiG[i]; ++jn) {
                int j = o->nnlG[(size_t)i * o->cfg.max_gi + jn];
                for (a = 0; a < D; ++a) {
                    if (o->cfg.quad_point_h4) /* :1391-1392,1402 */
                        xij[a] = o->x[a][i] + o->cfg.h / 4. * (o->gx[a][j] - o->x[a][i]);
                    else
                        xij[a] = (o->x[a][i] + o->gx[a][j]) / 2.;
                    xijxi[a] = xij[a] - o->x[a][i];
                }
                if (psiMaxNgb < fGhost[j]) psiMaxNgb = fGhost[j];
                if (psiMinNgb > fGhost[j]) psiMinNgb = fGhost[j];
                fij = f[i] + dotp(g, xijxi, D);
                if (psiMaxMid < fij) psiMaxMid = fij;
                if (psiMinMid > fij) psiMinMid = fij;
            }
        alphaMax = q1_abs((psiMaxNgb - f[i]) / (psiMaxMid - f[i]), o->cfg.abs_mode);
        alphaMin = q1_abs((f[i] - psiMinNgb) / (f[i] - psiMinMid), o->cfg.abs_mode);
        if (alphaMin <= alphaMax && o->cfg.beta * alphaMin < 1.) {
            for (a = 0; a < D; ++a) g[a] *= alphaMin;
        } else if (alphaMax <= alphaMin && o->cfg.beta * alphaMax < 1.) {
            for (a = 0; a < D; ++a) g[a] *= alphaMax;
        }
    }
}

/* ---------- Particles::compEffectiveFace, Particles.cpp:1290-1311 (ghost overload :2504-2531, quirk Q9) ---------- */
static void effective_face(Orc *o) {
    const int D = o->D, M = o->cfg.max_ni, MG = o->cfg.max_gi;
    int i, j, a;
    for (i = 0; i < o->N; ++i)
        for (j = 0; j < o->noi[i]; ++j) {
            int ji = o->nnl[(size_t)i * M + j], ij;
            for (ij = 0; ij < o->noi[ji]; ++ij)
                if (o->nnl[ij + (size_t)ji * M] == i) break;
            for (a = 0; a < D; ++a)
                o->Aij[((size_t)i * M + j) * D + a] = 1. / o->omega[i] * o->psij[((size_t)i * M + j) * D + a] -
                                                      1. / o->omega[ji] * o->psij[((size_t)ij + (size_t)ji * M) * D + a];
        }
    if (!o->cfg.periodic) return;
    memset(o->one_sided, 0, sizeof(int) * o->N);
    for (i = 0; i < o->N; ++i)
        for (j = 0; j < o->noiG[i]; ++j) {
            size_t ii = (size_t)i * MG + j;
            int ji = o->nnlG[ii], par = o->gparent[ji], ij;
            size_t rev;
            for (ij = 0; ij < o->noiG[par]; ++ij)
                if (o->gparent[o->nnlG[ij + (size_t)par * MG]] == i) break;
            if (ij == o->noiG[par]) { /* reverse slot missing: the reference reads a stale slot */
                o->one_sided[i] = 1;
                o->one_sided[par] = 1;
            }
            rev = (size_t)ij + (size_t)par * MG;
            if (rev >= (size_t)o->N * MG) rev = (size_t)o->N * MG - 1; /* keep the stale read in bounds */
            for (a = 0; a < D; ++a)
                o->AijG[ii * D + a] = 1. / o->omega[i] * o->psijG[ii * D + a] - 1. / o->gomega[ji] * o->psijG[rev * D + a];
        }
}

/* ---------- Particles::pairwiseLimiter, Particles.cpp:1735-1785 (quirk Q1) ---------- */
static double pairwise_limiter(const Orc *o, double phi0, double phi_i, double phi_j, double xijxi_abs, double xjxi_abs) {
    const int am = o->cfg.abs_mode;
    double phi_ = phi_i;
    double phi_ij = phi_i + xijxi_abs / xjxi_abs * (phi_j - phi_i);
    double phiMin, phiMax, delta1, delta2, phiMinus, phiPlus;
    if (phi_i < phi_j) {
        phiMin = phi_i;
        phiMax = phi_j;
    } else {
        phiMin = phi_j;
        phiMax = phi_i;
    }
    delta1 = o->cfg.psi1 * q1_abs(phi_i - phi_j, am);
    delta2 = o->cfg.psi2 * q1_abs(phi_i - phi_j, am);
    if ((phiMax + delta1 >= 0. && phiMax >= 0.) || (phiMax + delta1 < 0. && phiMax < 0.)) {
        phiPlus = phiMax + delta1;
    } else {
        phiPlus = phiMax / (1. + delta1 / q1_abs(phiMax, am));
    }
    if ((phiMin - delta1 >= 0. && phiMin >= 0.) || (phiMin - delta1 < 0. && phiMin < 0.)) {
        phiMinus = phiMin - delta1;
    } else {
        phiMinus = phiMin / (1. + delta1 / q1_abs(phiMin, am));
    }
    if (phi_i < phi_j) {
        double minPhiD2;
        if (phi_ij + delta2 < phi0)
            minPhiD2 = phi_ij + delta2;
        else
            minPhiD2 = phi0;
        phi_ = phiMinus > minPhiD2 ? phiMinus : minPhiD2;
    } else if (phi_i > phi_j) {
        double maxPhiD2;
        if (phi_ij - delta2 > phi0)
            maxPhiD2 = phi_ij - delta2;
        else
            maxPhiD2 = phi0;
        phi_ = phiPlus < maxPhiD2 ? phiPlus : maxPhiD2;
    }
    return phi_;
}

/* field index of W component nu: W = [rho, P, vx, vy, vz] -> grad field [0=rho,4=P,1=vx,2=vy,3=vz] */
static const int W2F[5] = {0, 4, 1, 2, 3};

/* ---------- Particles::compRiemannStatesLR, Particles.cpp:1488-1733 (quirks Q3, Q13) ---------- */
static void riemann_states(Orc *o, double dt) {
    const int D = o->D, M = o->cfg.max_ni, NW = o->D + 2;
    const double gamma = o->cfg.gamma;
    int i, jn, a, nu;
    for (i = 0; i < o->N; ++i) {
        double xijxi[3], xjxi[3], xijxj[3];
        /* Q13: xjxi[2] is never written in the first-order 3D branch (:1526-1530) */
        xjxi[2] = 0.;
        for (jn = 0; jn < o->noi[i]; ++jn) {
            int j = o->nnl[(size_t)i * M + jn];
            size_t iW = (size_t)i * M + jn;
            double *WR = &o->WijR[iW * NW], *WL = &o->WijL[iW * NW], *vF = &o->vFrame[iW * D];
            double WR_buf[5], WL_buf[5], viDiv, vjDiv;
            const double *gi[5], *gj[5];
            xjxi[0] = o->x[0][j] - o->x[0][i];
            xjxi[1] = o->x[1][j] - o->x[1][i];
            if (D == 3 && (o->cfg.q13_mode == 1 || o->cfg.quad_point_h4)) xjxi[2] = o->x[2][j] - o->x[2][i]; /* :1531 */
            if (o->cfg.quad_point_h4) { /* FIRST_ORDER_QUAD_POINT 0, :1514-1523,1531-1537 */
                for (a = 0; a < D; ++a) {
                    double xij = o->x[a][i] + o->cfg.h / 4. * xjxi[a];
                    xijxi[a] = xij - o->x[a][i];
                    xijxj[a] = xij - o->x[a][j];
                }
            } else {
                for (a = 0; a < D; ++a) {
                    xijxj[a] = .5 * (o->x[a][i] - o->x[a][j]);
                    xijxi[a] = .5 * (o->x[a][j] - o->x[a][i]);
                }
            }
            if (!o->cfg.move_particles) {
                for (a = 0; a < D; ++a) vF[a] = 0.;
            } else if (o->cfg.quad_point_h4) { /* :1556-1563 */
                double dotProd = dotp(xijxi, xjxi, D), dSqr = dotp(xjxi, xjxi, D);
                for (a = 0; a < D; ++a) vF[a] = o->v[a][i] + (o->v[a][j] - o->v[a][i]) * dotProd / dSqr;
            } else {
                for (a = 0; a < D; ++a) vF[a] = (o->v[a][i] + o->v[a][j]) / 2.;
            }
            WR[0] = o->rho[i];
            WL[0] = o->rho[j];
            WR[1] = o->P[i];
            WL[1] = o->P[j];
            for (a = 0; a < D; ++a) {
                WR[2 + a] = o->v[a][i] - vF[a];
                WL[2 + a] = o->v[a][j] - vF[a];
            }
            for (nu = 0; nu < NW; ++nu) {
                WR_buf[nu] = WR[nu];
                WL_buf[nu] = WL[nu];
                gi[nu] = &o->grad[W2F[nu]][(size_t)i * D];
                gj[nu] = &o->grad[W2F[nu]][(size_t)j * D];
            }
            for (nu = 0; nu < NW; ++nu) { /* :1609-1620 */
                WR[nu] += dotp(gi[nu], xijxi, D);
                WL[nu] += dotp(gj[nu], xijxj, D);
            }
            if (o->cfg.pairwise) { /* :1622-1637 */
                double xijxi_abs = 0., xijxj_abs = 0., xjxi_abs = 0.;
                for (a = 0; a < D; ++a) {
                    xijxi_abs += xijxi[a] * xijxi[a];
                    xijxj_abs += xijxj[a] * xijxj[a];
                    xjxi_abs += xjxi[a] * xjxi[a];
                }
                xijxi_abs = sqrt(xijxi_abs);
                xijxj_abs = sqrt(xijxj_abs);
                xjxi_abs = sqrt(xjxi_abs);
                for (nu = 0; nu < NW; ++nu) {
                    WR[nu] = pairwise_limiter(o, WR[nu], WR_buf[nu], WL_buf[nu], xijxi_abs, xjxi_abs);
                    WL[nu] = pairwise_limiter(o, WL[nu], WL_buf[nu], WR_buf[nu], xijxj_abs, xjxi_abs);
                }
            }
            /* half-step prediction, :1664-1721.  gi[2]=vxGrad, gi[3]=vyGrad, gi[4]=vzGrad, gi[0]=rhoGrad, gi[1]=PGrad */
            viDiv = gi[2][0] + gi[3][1];
            vjDiv = gj[2][0] + gj[3][1];
            if (D == 3) {
                viDiv += gi[4][2];
                vjDiv += gj[4][2];
            }
#define VI(a_) (o->v[a_][i] - vF[a_])
#define VJ(a_) (o->v[a_][j] - vF[a_])
            WR[0] -= dt / 2. * (o->rho[i] * viDiv + VI(0) * gi[0][0] + VI(1) * gi[0][1]);
            WL[0] -= dt / 2. * (o->rho[j] * vjDiv + VJ(0) * gj[0][0] + VJ(1) * gj[0][1]);
            WR[1] -= dt / 2. * (gamma * o->P[i] * viDiv + VI(0) * gi[1][0] + VI(1) * gi[1][1]);
            WL[1] -= dt / 2. * (gamma * o->P[j] * vjDiv + VJ(0) * gj[1][0] + VJ(1) * gj[1][1]);
            WR[2] -= dt / 2. * (gi[1][0] / o->rho[i] + VI(0) * gi[2][0] + VI(1) * gi[2][1]);
            WL[2] -= dt / 2. * (gj[1][0] / o->rho[j] + VJ(0) * gj[2][0] + VJ(1) * gj[2][1]);
            WR[3] -= dt / 2. * (gi[1][1] / o->rho[i] + VI(0) * gi[3][0] + VI(1) * gi[3][1]);
            WL[3] -= dt / 2. * (gj[1][1] / o->rho[j] + VJ(0) * gj[3][0] + VJ(1) * gj[3][1]);
            if (D == 3) {
                double vzL = (o->cfg.q3_mode == 1) ? VJ(2) : VI(2); /* Q3: :1717,1719 use vz[i] */
                WR[0] -= dt / 2. * VI(2) * gi[0][2];
                WL[0] -= dt / 2. * VJ(2) * gj[0][2];
                WR[1] -= dt / 2. * VI(2) * gi[1][2];
                WL[1] -= dt / 2. * VJ(2) * gj[1][2];
                WR[2] -= dt / 2. * VI(2) * gi[2][2];
                WL[2] -= dt / 2. * vzL * gj[2][2];
                WR[3] -= dt / 2. * VI(2) * gi[3][2];
                WL[3] -= dt / 2. * vzL * gj[3][2];
                WR[4] -= dt / 2. * (gi[1][2] / o->rho[i] + VI(0) * gi[4][0] + VI(1) * gi[4][1] + VI(2) * gi[4][2]);
                WL[4] -= dt / 2. * (gj[1][2] / o->rho[j] + VJ(0) * gj[4][0] + VJ(1) * gj[4][1] + VJ(2) * gj[4][2]);
            }
#undef VI
#undef VJ
        }
    }
}

/* ---------- compRiemannStatesLR ghost overload, Particles.cpp:2533-2683 (2D; no pairwise limiter) ---------- */
static void riemann_states_ghosts(Orc *o, double dt) {
    const int D = o->D, MG = o->cfg.max_gi, NW = o->D + 2;
    const double gamma = o->cfg.gamma;
    int i, jn, a, nu;
    for (i = 0; i < o->N; ++i)
        for (jn = 0; jn < o->noiG[i]; ++jn) {
            int j = o->nnlG[(size_t)i * MG + jn];
            size_t iW = (size_t)i * MG + jn;
            double *WR = &o->WijRG[iW * NW], *WL = &o->WijLG[iW * NW], *vF = &o->vFrameG[iW * D];
            double xijxi[3], xijxj[3], viDiv, vjDiv;
            const double *gi[5], *gj[5];
            if (o->cfg.quad_point_h4) { /* :2563-2570, 2602-2611 */
                double xjxi[3], dotProd, dSqr;
                for (a = 0; a < D; ++a) {
                    double xij;
                    xjxi[a] = o->gx[a][j] - o->x[a][i];
                    xij = o->x[a][i] + o->cfg.h / 4. * xjxi[a];
                    xijxi[a] = xij - o->x[a][i];
                    xijxj[a] = xij - o->gx[a][j];
                }
                dotProd = dotp(xijxi, xjxi, D);
                dSqr = dotp(xjxi, xjxi, D);
                for (a = 0; a < D; ++a)
                    vF[a] = o->cfg.move_particles ? o->v[a][i] + (o->gv[a][j] - o->v[a][i]) * dotProd / dSqr : 0.;
            } else
            for (a = 0; a < D; ++a) {
                xijxj[a] = .5 * (o->x[a][i] - o->gx[a][j]);
                xijxi[a] = .5 * (o->gx[a][j] - o->x[a][i]);
                if (o->cfg.move_particles)
                    vF[a] = (o->v[a][i] + o->gv[a][j]) / 2.;
                else
                    vF[a] = 0.;
            }
            WR[0] = o->rho[i];
            WL[0] = o->grho[j];
            WR[1] = o->P[i];
            WL[1] = o->gP[j];
            for (a = 0; a < D; ++a) {
                WR[2 + a] = o->v[a][i] - vF[a];
                WL[2 + a] = o->gv[a][j] - vF[a];
            }
            for (nu = 0; nu < NW; ++nu) {
                gi[nu] = &o->grad[W2F[nu]][(size_t)i * D];
                gj[nu] = &o->ggrad[W2F[nu]][(size_t)j * D];
            }
            for (nu = 0; nu < NW; ++nu) {
                WR[nu] += dotp(gi[nu], xijxi, D);
                WL[nu] += dotp(gj[nu], xijxj, D);
            }
            viDiv = gi[2][0] + gi[3][1];
            vjDiv = gj[2][0] + gj[3][1];
#define VI(a_) (o->v[a_][i] - vF[a_])
#define VJ(a_) (o->gv[a_][j] - vF[a_])
            WR[0] -= dt / 2. * (o->rho[i] * viDiv + VI(0) * gi[0][0] + VI(1) * gi[0][1]);
            WL[0] -= dt / 2. * (o->grho[j] * vjDiv + VJ(0) * gj[0][0] + VJ(1) * gj[0][1]);
            WR[1] -= dt / 2. * (gamma * o->P[i] * viDiv + VI(0) * gi[1][0] + VI(1) * gi[1][1]);
            WL[1] -= dt / 2. * (gamma * o->gP[j] * vjDiv + VJ(0) * gj[1][0] + VJ(1) * gj[1][1]);
            WR[2] -= dt / 2. * (gi[1][0] / o->rho[i] + VI(0) * gi[2][0] + VI(1) * gi[2][1]);
            WL[2] -= dt / 2. * (gj[1][0] / o->grho[j] + VJ(0) * gj[2][0] + VJ(1) * gj[2][1]);
            WR[3] -= dt / 2. * (gi[1][1] / o->rho[i] + VI(0) * gi[3][0] + VI(1) * gi[3][1]);
            WL[3] -= dt / 2. * (gj[1][1] / o->grho[j] + VJ(0) * gj[3][0] + VJ(1) * gj[3][1]);
#undef VI
#undef VJ
        }
}

/* ---------- Particles::solveRiemannProblems, Particles.cpp:1787-1911 (quirks Q4, Q5, Q9) ---------- */
static void solve_riemann(Orc *o) {
    const int D = o->D, M = o->cfg.max_ni, MG = o->cfg.max_gi, NW = o->D + 2;
    int i, j, d;
    for (i = 0; i < o->N; ++i) {
        for (j = 0; j < o->noi[i]; ++j) {
            size_t ii = (size_t)i * M + j, iij = 0;
            int ji = o->nnl[ii], compute = 1;
            if (ji < i) {
                int ij;
                compute = 0;
                for (ij = 0; ij < o->noi[ji]; ++ij)
                    if (o->nnl[ij + (size_t)ji * M] == i) break;
                iij = (size_t)ji * M + ij;
            }
            if (compute) { /* Riemann solver{WijL, WijR, vFrame, Aij, i} -> ctor (WR, WL, ...) */
                orc_face_flux(D, o->cfg.mfm, o->cfg.gamma, &o->WijL[ii * NW], &o->WijR[ii * NW], &o->vFrame[ii * D],
                              &o->Aij[ii * D], &o->Fij[ii * NW]);
            } else {
                for (d = 0; d < NW; ++d) o->Fij[ii * NW + d] = -o->Fij[iij * NW + d];
            }
        }
        if (!o->cfg.periodic) continue;
        for (j = 0; j < o->noiG[i]; ++j) {
            size_t ii = (size_t)i * MG + j, iij = 0;
            int ji = o->nnlG[ii], compute = 1;
            if (o->WijRG[ii * NW + 1] < 0. || o->WijLG[ii * NW + 1] < 0.) o->errFlags |= 4; /* reference: exit(6) if DEBUG_LVL */
            if (o->gparent[ji] < i) {
                int ij, par = o->gparent[ji];
                compute = 0;
                for (ij = 0; ij < o->noiG[par]; ++ij)
                    if (o->gparent[o->nnlG[ij + (size_t)par * MG]] == i) break;
                iij = (size_t)ij + (size_t)par * MG;
                if (iij >= (size_t)o->N * MG) iij = (size_t)o->N * MG - 1;
            }
            if (compute) {
                orc_face_flux(D, o->cfg.mfm, o->cfg.gamma, &o->WijLG[ii * NW], &o->WijRG[ii * NW], &o->vFrameG[ii * D],
                              &o->AijG[ii * D], &o->FijG[ii * NW]);
            } else {
                for (d = 0; d < NW; ++d) o->FijG[ii * NW + d] = -o->FijG[iij * NW + d];
            }
        }
    }
}

/* ---------- Particles::collectFluxes, Particles.cpp:1913-2011 ---------- */
static void collect_fluxes(Orc *o) {
    const int D = o->D, M = o->cfg.max_ni, MG = o->cfg.max_gi, NW = o->D + 2;
    int i, j, a;
    for (i = 0; i < o->N; ++i) {
        o->mF[i] = 0.;
        for (a = 0; a < D; ++a) o->vF[(size_t)i * D + a] = 0.;
        o->eF[i] = 0.;
        for (j = 0; j < o->noi[i]; ++j) {
            size_t ii = (size_t)j + (size_t)i * M;
            o->mF[i] += o->Fij[ii * NW + 0];
            for (a = 0; a < D; ++a) o->vF[(size_t)i * D + a] += o->Fij[ii * NW + 2 + a];
            o->eF[i] += o->Fij[ii * NW + 1];
        }
        if (o->cfg.periodic)
            for (j = 0; j < o->noiG[i]; ++j) {
                size_t ii = (size_t)j + (size_t)i * MG;
                o->mF[i] += o->FijG[ii * NW + 0];
                o->vF[(size_t)i * D + 0] += o->FijG[ii * NW + 2];
                o->vF[(size_t)i * D + 1] += o->FijG[ii * NW + 3];
                o->eF[i] += o->FijG[ii * NW + 1];
            }
    }
}

/* ---------- Particles::updateStateAndPosition, Particles.cpp:2013-2110 ---------- */
static void update_state(Orc *o, double dt) {
    const int D = o->D;
    int i, a;
    for (i = 0; i < o->N; ++i) {
        double vi[3] = {0., 0., 0.}, Q[4], v2;
        for (a = 0; a < D; ++a) vi[a] = o->v[a][i];
        if (D == 3)
            Q[0] = o->m[i] * (o->u[i] + .5 * (vi[0] * vi[0] + vi[1] * vi[1] + vi[2] * vi[2]));
        else
            Q[0] = o->m[i] * (o->u[i] + .5 * (vi[0] * vi[0] + vi[1] * vi[1]));
        for (a = 0; a < D; ++a) Q[1 + a] = o->m[i] * vi[a];
        if (!o->cfg.mfm) o->m[i] -= dt * o->mF[i];
        for (a = 0; a < D; ++a) Q[1 + a] -= dt * o->vF[(size_t)i * D + a];
        for (a = 0; a < D; ++a) o->v[a][i] = Q[1 + a] / o->m[i];
        Q[0] -= dt * o->eF[i];
        if (D == 3)
            v2 = o->v[0][i] * o->v[0][i] + o->v[1][i] * o->v[1][i] + o->v[2][i] * o->v[2][i];
        else
            v2 = o->v[0][i] * o->v[0][i] + o->v[1][i] * o->v[1][i];
        o->u[i] = Q[0] / o->m[i] - .5 * v2;
        if (o->cfg.move_particles) {
            for (a = 0; a < D; ++a) o->x[a][i] += vi[a] * dt;
            if (o->cfg.periodic)
                for (a = 0; a < D; ++a) {
                    if (o->x[a][i] < o->bmin[a]) {
                        o->x[a][i] = o->bmax[a] - (o->bmin[a] - o->x[a][i]);
                    } else if (o->bmax[a] <= o->x[a][i]) {
                        o->x[a][i] = o->bmin[a] + (o->x[a][i] - o->bmax[a]);
                    }
                }
        }
    }
}

/* ======================= driver (MeshlessScheme.cpp:39-253) ======================= */
static double now_sec(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void *orc_create(const orc_config *cfg, int N, const double *x, const double *y, const double *z,
                 const double *vx, const double *vy, const double *vz, const double *m, const double *u) {
    Orc *o = (Orc *)calloc(1, sizeof(Orc));
    const double *xs[3] = {x, y, z}, *vs[3] = {vx, vy, vz};
    int k, f, D = cfg->dim;
    size_t M = cfg->max_ni, MG = cfg->periodic ? cfg->max_gi : 0, n = N;
    o->cfg = *cfg;
    o->N = N;
    o->D = D;
    for (k = 0; k < D; ++k) {
        o->x[k] = dalloc(n);
        o->v[k] = dalloc(n);
        memcpy(o->x[k], xs[k], sizeof(double) * n);
        memcpy(o->v[k], vs[k], sizeof(double) * n);
    }
    o->m = dalloc(n); o->u = dalloc(n); o->rho = dalloc(n); o->P = dalloc(n); o->omega = dalloc(n);
    memcpy(o->m, m, sizeof(double) * n);
    memcpy(o->u, u, sizeof(double) * n);
    o->cell = ialloc(n);
    for (f = 0; f < 5; ++f) {
        o->grad[f] = dalloc(n * D);
        o->gradPre[f] = dalloc(n * D);
    }
    o->Binv = dalloc(n * D * D);
    o->nnl = ialloc(n * M); o->noi = ialloc(n);
    o->psij = dalloc(n * M * D); o->Aij = dalloc(n * M * D); o->vFrame = dalloc(n * M * D);
    o->WijL = dalloc(n * M * (D + 2)); o->WijR = dalloc(n * M * (D + 2)); o->Fij = dalloc(n * M * (D + 2));
    o->mF = dalloc(n); o->eF = dalloc(n); o->vF = dalloc(n * D);
    o->cellList = ialloc(n);
    o->one_sided = ialloc(n);
    if (cfg->periodic) {
        size_t ng = (size_t)D * n;
        for (k = 0; k < D; ++k) {
            o->gx[k] = dalloc(ng);
            o->gv[k] = dalloc(ng);
        }
        o->gparent = ialloc(ng); o->ghostMap = ialloc(n * 3);
        o->grho = dalloc(ng); o->gP = dalloc(ng); o->gomega = dalloc(ng);
        for (f = 0; f < 5; ++f) o->ggrad[f] = dalloc(ng * D);
        o->nnlG = ialloc(n * MG); o->noiG = ialloc(n);
        o->psijG = dalloc(n * MG * D); o->AijG = dalloc(n * MG * D); o->vFrameG = dalloc(n * MG * D);
        o->WijLG = dalloc(n * MG * (D + 2)); o->WijRG = dalloc(n * MG * (D + 2)); o->FijG = dalloc(n * MG * (D + 2));
        for (k = 0; k < D; ++k) {
            o->bmin[k] = cfg->box[k];
            o->bmax[k] = cfg->box[D + k];
        }
    } else {
        domain_limits(o); /* main.cpp:99-100 */
    }
    create_grid(o); /* MeshlessScheme.cpp:17 */
    return o;
}

void orc_destroy(void *ctx) {
    Orc *o = (Orc *)ctx;
    int k, f;
    for (k = 0; k < 3; ++k) {
        free(o->x[k]); free(o->v[k]); free(o->gx[k]); free(o->gv[k]);
    }
    free(o->m); free(o->u); free(o->rho); free(o->P); free(o->omega); free(o->cell);
    for (f = 0; f < 5; ++f) {
        free(o->grad[f]); free(o->gradPre[f]); free(o->ggrad[f]);
    }
    free(o->Binv); free(o->nnl); free(o->noi); free(o->psij); free(o->Aij); free(o->vFrame);
    free(o->WijL); free(o->WijR); free(o->Fij); free(o->mF); free(o->eF); free(o->vF);
    free(o->cellList); free(o->cellStart); free(o->one_sided);
    free(o->gparent); free(o->ghostMap); free(o->grho); free(o->gP); free(o->gomega);
    free(o->nnlG); free(o->noiG); free(o->psijG); free(o->AijG); free(o->vFrameG);
    free(o->WijLG); free(o->WijRG); free(o->FijG);
    free(o);
}

double orc_step(void *ctx, double dtFixed, double dtMax, int stopAfter) {
    Orc *o = (Orc *)ctx;
    const int per = o->cfg.periodic, D = o->D;
    double t0 = now_sec(), t1, timeStep;
    int f;
#define LAP(k) do { t1 = now_sec(); o->phaseSec[k] = t1 - t0; t0 = t1; } while (0)
    if (!per) { /* MeshlessScheme.cpp:41-51 */
        domain_limits(o);
        create_grid(o);
    }
    assign_cells(o); /* :53 */
    LAP(0);
    if (per) create_ghosts(o); /* :60 */
    grid_nns(o);               /* :66 */
    if (per) ghost_nns(o);     /* :69 */
    LAP(1);
    density_pressure(o); /* :73-79 */
    LAP(2);
    if (dtFixed >= 0.) { /* a zero-length step is what the reference driver does at every dump time (quirk Q7) */
        timeStep = dtFixed;
    } else {
        timeStep = global_timestep(o); /* :93 */
        o->dtCfl = timeStep;
        if (dtMax > 0. && timeStep > dtMax) timeStep = dtMax;
    }
    LAP(3);
    if (per) update_ghost_state(o); /* :109 */
    psij_tilde(o);                  /* :110 / :132 */
    {
        const double *fld[5] = {o->rho, o->v[0], o->v[1], o->v[2], o->P};
        const double *gfld[5] = {o->grho, o->gv[0], o->gv[1], o->gv[2], o->gP};
        for (f = 0; f < 5; ++f) {
            if (f == 3 && D == 2) continue;
            gradient(o, fld[f], o->grad[f], gfld[f]); /* :114-120 / :133-139 */
        }
        if (per) update_ghost_gradients(o); /* :122 */
        for (f = 0; f < 5; ++f) memcpy(o->gradPre[f], o->grad[f], sizeof(double) * (size_t)o->N * D);
        LAP(4);
        if (o->cfg.slope_limiting) {
            /* wrapper order rho, vx, vy, (vz), P: Particles.cpp:1329-1335 */
            for (f = 0; f < 5; ++f) {
                if (f == 3 && D == 2) continue;
                slope_limiter(o, fld[f], o->grad[f], gfld[f]);
            }
            if (per) update_ghost_gradients(o); /* :129 */
        }
    }
    LAP(5);
    effective_face(o); /* :148-150 */
    LAP(6);
    riemann_states(o, timeStep); /* :153 */
    if (per) riemann_states_ghosts(o, timeStep); /* :157 */
    LAP(7);
    o->dt = timeStep;
    if (stopAfter == 1) return timeStep;
    solve_riemann(o); /* :204-207 */
    LAP(8);
    collect_fluxes(o); /* :221 */
    LAP(9);
    update_state(o, timeStep); /* :224 */
    LAP(10);
#undef LAP
    return timeStep;
}

double orc_last_dt_cfl(void *ctx) { return ((Orc *)ctx)->dtCfl; }
void orc_phase_seconds(void *ctx, double *out) { memcpy(out, ((Orc *)ctx)->phaseSec, sizeof(double) * 11); }

/* Particles.cpp:2830-2886: serial ascending-i sums */
void orc_sums(void *ctx, double *out) {
    Orc *o = (Orc *)ctx;
    int i;
    double V = 0., M = 0., E = 0., px = 0., py = 0., pz = 0.;
    for (i = 0; i < o->N; ++i) V += 1. / o->omega[i];
    for (i = 0; i < o->N; ++i) M += o->m[i];
    for (i = 0; i < o->N; ++i) {
        if (o->D == 2)
            E += o->m[i] * (o->u[i] + .5 * (o->v[0][i] * o->v[0][i] + o->v[1][i] * o->v[1][i]));
        else
            E += o->m[i] * (o->u[i] + .5 * (o->v[0][i] * o->v[0][i] + o->v[1][i] * o->v[1][i] + o->v[2][i] * o->v[2][i]));
    }
    for (i = 0; i < o->N; ++i) px += o->m[i] * o->v[0][i];
    for (i = 0; i < o->N; ++i) py += o->m[i] * o->v[1][i];
    if (o->D == 3)
        for (i = 0; i < o->N; ++i) pz += o->m[i] * o->v[2][i];
    out[0] = V; out[1] = M; out[2] = E; out[3] = px; out[4] = py; out[5] = pz;
}

void orc_grid(void *ctx, int *cells, double *cellSize, double *bounds) {
    Orc *o = (Orc *)ctx;
    int k;
    for (k = 0; k < 3; ++k) {
        cells[k] = o->cells[k];
        cellSize[k] = o->cellSize[k];
    }
    for (k = 0; k < o->D; ++k) {
        bounds[k] = o->bmin[k];
        bounds[o->D + k] = o->bmax[k];
    }
}

long orc_fetch(void *ctx, const char *name, void *dst) {
    Orc *o = (Orc *)ctx;
    const long N = o->N, D = o->D, M = o->cfg.max_ni, MG = o->cfg.max_gi;
#define RET(ptr, count, type)                                           \
    do {                                                                \
        if (dst) memcpy(dst, (const void *)(ptr), sizeof(type) * (count)); \
        return (long)(count);                                           \
    } while (0)
#define IS(s) (strcmp(name, s) == 0)
    if (IS("x")) RET(o->x[0], N, double);
    if (IS("y")) RET(o->x[1], N, double);
    if (IS("vx")) RET(o->v[0], N, double);
    if (IS("vy")) RET(o->v[1], N, double);
    if (D == 3) {
        if (IS("z")) RET(o->x[2], N, double);
        if (IS("vz")) RET(o->v[2], N, double);
        if (IS("vzGrad")) RET(o->grad[3], N * D, double);
    }
    if (IS("m")) RET(o->m, N, double);
    if (IS("u")) RET(o->u, N, double);
    if (IS("rho")) RET(o->rho, N, double);
    if (IS("P")) RET(o->P, N, double);
    if (IS("omega")) RET(o->omega, N, double);
    if (IS("cell")) RET(o->cell, N, int);
    if (IS("noi")) RET(o->noi, N, int);
    if (IS("nnl")) RET(o->nnl, N * M, int);
    if (IS("rhoGrad")) RET(o->grad[0], N * D, double);
    if (IS("vxGrad")) RET(o->grad[1], N * D, double);
    if (IS("vyGrad")) RET(o->grad[2], N * D, double);
    if (IS("PGrad")) RET(o->grad[4], N * D, double);
    if (IS("gradPre")) {
        long cnt = 0;
        int f;
        for (f = 0; f < 5; ++f) {
            if (f == 3 && D == 2) continue;
            if (dst) memcpy((double *)dst + cnt, o->gradPre[f], sizeof(double) * N * D);
            cnt += N * D;
        }
        return cnt;
    }
    if (IS("Binv")) RET(o->Binv, N * D * D, double);
    if (IS("psijTilde")) RET(o->psij, N * M * D, double);
    if (IS("Aij")) RET(o->Aij, N * M * D, double);
    if (IS("WijL")) RET(o->WijL, N * M * (D + 2), double);
    if (IS("WijR")) RET(o->WijR, N * M * (D + 2), double);
    if (IS("Fij")) RET(o->Fij, N * M * (D + 2), double);
    if (IS("vFrame")) RET(o->vFrame, N * M * D, double);
    if (IS("mF")) RET(o->mF, N, double);
    if (IS("eF")) RET(o->eF, N, double);
    if (IS("vF")) RET(o->vF, N * D, double);
    if (IS("err_flags")) {
        if (dst) *(int *)dst = o->errFlags;
        return 1;
    }
    if (o->cfg.periodic) {
        if (IS("noiGhosts")) RET(o->noiG, N, int);
        if (IS("nnlGhosts")) RET(o->nnlG, N * MG, int);
        if (IS("ghostMap")) RET(o->ghostMap, N * 3, int);
        if (IS("AijGhosts")) RET(o->AijG, N * MG * D, double);
        if (IS("FijGhosts")) RET(o->FijG, N * MG * (D + 2), double);
        if (IS("WijLGhosts")) RET(o->WijLG, N * MG * (D + 2), double);
        if (IS("WijRGhosts")) RET(o->WijRG, N * MG * (D + 2), double);
        if (IS("ghost_N")) {
            if (dst) *(int *)dst = o->Ng;
            return 1;
        }
        if (IS("ghost_x")) RET(o->gx[0], o->Ng, double);
        if (IS("ghost_y")) RET(o->gx[1], o->Ng, double);
        if (IS("ghost_parent")) RET(o->gparent, o->Ng, int);
        if (IS("one_sided")) RET(o->one_sided, N, int);
    }
#undef RET
#undef IS
    return -1;
}

#ifdef RS_STATS
/* out[0..63] Newton iterations per solve, out[64..127] Brent iterations per Brent call; reset afterwards */
void orc_rs_stats(long *out) {
    int k;
    for (k = 0; k < 64; ++k) {
        out[k] = rs_stat_newton[k];
        out[64 + k] = rs_stat_brent[k];
        rs_stat_newton[k] = rs_stat_brent[k] = 0;
    }
}
#endif
