// K4 (+K5) -- effective faces, slope-limited reconstruction + half-step prediction, exact Riemann
// solver per face, gather-side flux sum, conserved-variable update and drift.
//
// Replaces, per particle and per neighbour slot (reference operation order kept):
//   Particles::compEffectiveFace        /root/reference/demonstrator/src/Particles.cpp:1290-1311 (ghosts :2504-2531)
//   Particles::compRiemannStatesLR      :1488-1733 (ghosts :2533-2683), pairwiseLimiter :1735-1785
//   Particles::solveRiemannProblems     :1787-1911, Riemann::Riemann/exact/rotateAndProjectFluxes (Riemann.cpp:7-229),
//   Helper::rotationMatrix2D/3D         Helper.cpp:39-77, RiemannSolver::solve (restated, see oracle/riemann_exact.h)
//   Particles::collectFluxes            :1913-2011, Particles::updateStateAndPosition :2013-2110
//
// Gather-side and atomic-free: thread i evaluates every face (i,j) of its own list.  The reference
// (ENFORCE_FLUX_SYM, quirk Q4) solves a face once, from the endpoint with the LOWER ORIGINAL index,
// and gives the other endpoint the exact negation; here both endpoint threads evaluate that same
// canonical orientation (operands are swapped with selects, not branches, so a warp does not
// diverge on orientation) and the non-canonical one negates.  Both threads therefore add bit-identical
// +-F and total mass, momentum and energy are conserved to round-off without any exchange.
// The per-slot buffers of the reference (psijTilde, Aij, WijL/R, Fij, vFrame: ~100 kB per particle,
// Particles.h:201-229) do not exist: psi-tilde of BOTH endpoints is recomputed from Binv and omega.
//
// Roofline: FP64 pipe (exact Riemann solver: pow/sqrt/div heavy, ~2-3 kFLOP per face, K faces per
// particle); algorithmic bytes 24 (2D) / 38 (3D) doubles per particle.
#include "mlh_internal.cuh"
#include <cfloat>

namespace {

// ---------------------------------------------------------------------------------------------
// exact Riemann solver (device restatement; iteration-for-iteration the algorithm of
// oracle/riemann_exact.h -- Newton-Raphson with Toro's adaptive guess, Brent fallback, sampling at x/t=0)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rs_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double rs_min(double a, double b) { return (b < a) ? b : a; }

__device__ __forceinline__ double rs_fb(const RsConsts &c, double rho, double P, double a, double Pstar) {
    if (Pstar > P) {
        double A = c.tdgp1 / rho;
        double B = c.gm1dgp1 * P;
        return (Pstar - P) * sqrt(A / (Pstar + B));
    }
    return c.tdgm1 * a * (pow(Pstar / P, c.gm1d2g) - 1.);
}
__device__ __forceinline__ double rs_fprimeb(const RsConsts &c, double rho, double P, double a, double Pstar) {
    if (Pstar > P) {
        double A = c.tdgp1 / rho;
        double B = c.gm1dgp1 * P;
        return (1. - 0.5 * (Pstar - P) / (B + Pstar)) * sqrt(A / (Pstar + B));
    }
    return 1. / (rho * a) * pow(Pstar / P, -c.gp1d2g);
}
__device__ __forceinline__ double rs_f(const RsConsts &c, double rhoL, double uL, double PL, double aL, double rhoR,
                                       double uR, double PR, double aR, double Pstar) {
    return rs_fb(c, rhoL, PL, aL, Pstar) + rs_fb(c, rhoR, PR, aR, Pstar) + (uR - uL);
}
__device__ __forceinline__ double rs_gb(const RsConsts &c, double rho, double P, double Pstar) {
    double A = c.tdgp1 / rho;
    double B = c.gm1dgp1 * P;
    return sqrt(A / (Pstar + B));
}
__device__ __forceinline__ double rs_guess_P(const RsConsts &c, double rhoL, double uL, double PL, double aL, double rhoR,
                                             double uR, double PR, double aR) {
    double Pguess;
    double Pmin = rs_min(PL, PR);
    double Pmax = rs_max(PL, PR);
    double qmax = Pmax / Pmin;
    double Ppv = 0.5 * (PL + PR) - 0.125 * (uR - uL) * (PL + PR) * (aL + aR);
    Ppv = rs_max(5.e-9 * (PL + PR), Ppv);
    if (qmax <= 2. && Pmin <= Ppv && Ppv <= Pmax) {
        Pguess = Ppv;
    } else if (Ppv < Pmin) {
        Pguess = pow((aL + aR - c.gm1d2 * (uR - uL)) / (aL / pow(PL, c.gm1d2g) + aR / pow(PR, c.gm1d2g)), c.tgdgm1);
    } else {
        double gL = rs_gb(c, rhoL, PL, Ppv);
        double gR = rs_gb(c, rhoR, PR, Ppv);
        Pguess = (gL * PL + gR * PR - uR + uL) / (gL + gR);
    }
    return rs_max(5.e-9 * (PL + PR), Pguess);
}

__device__ __noinline__ double rs_brent(const RsConsts cst, double rhoL, double uL, double PL, double aL, double rhoR,
                                        double uR, double PR, double aR, double lowerlimit, double upperlimit,
                                        double lowf, double upf) {
    double a = lowerlimit, b = upperlimit, c = 0., d = 1e230;
    double fa = lowf, fb = upf, fc = 0., s = 0., fs = 0.;
    bool mflag;
    if (fa * fb > 0.) return b;
    if (fabs(fa) < fabs(fb)) {
        double t = a; a = b; b = t;
        t = fa; fa = fb; fb = t;
    }
    c = a;
    fc = fa;
    mflag = true;
    while (!(fb == 0.) && (fabs(a - b) > 5.e-9 * (a + b))) {
        if ((fa != fc) && (fb != fc)) {
            s = a * fb * fc / (fa - fb) / (fa - fc) + b * fa * fc / (fb - fa) / (fb - fc) + c * fa * fb / (fc - fa) / (fc - fb);
        } else {
            s = b - fb * (b - a) / (fb - fa);
        }
        double tmp2 = 0.25 * (3. * a + b);
        if (!(((s > tmp2) && (s < b)) || ((s < tmp2) && (s > b))) || (mflag && (fabs(s - b) >= (0.5 * fabs(b - c)))) ||
            (!mflag && (fabs(s - b) >= (0.5 * fabs(c - d)))) || (mflag && (fabs(b - c) < 5.e-9 * (b + c))) ||
            (!mflag && (fabs(c - d) < 5.e-9 * (c + d)))) {
            s = 0.5 * (a + b);
            mflag = true;
        } else {
            mflag = false;
        }
        fs = rs_f(cst, rhoL, uL, PL, aL, rhoR, uR, PR, aR, s);
        d = c;
        c = b;
        fc = fb;
        if (fa * fs < 0.) {
            b = s;
            fb = fs;
        } else {
            a = s;
            fa = fs;
        }
        if (fabs(fa) < fabs(fb)) {
            double t = a; a = b; b = t;
            t = fa; fa = fb; fb = t;
        }
    }
    return b;
}

// vacuum sampling (Toro 4.6); cold path
__device__ __noinline__ int rs_solve_vacuum(const RsConsts c, double rhoL, double uL, double PL, double rhoR, double uR,
                                            double PR, double *rho, double *u, double *P) {
    const double dxdt = 0.;
    if (rhoL == 0. && rhoR == 0.) {
        *rho = 0.; *u = 0.; *P = 0.;
        return 0;
    }
    double aL = rhoL == 0. ? 0. : sqrt(c.gamma * PL / rhoL);
    double aR = rhoR == 0. ? 0. : sqrt(c.gamma * PR / rhoR);
    int side; // -1: sample left fan against vacuum, +1: right fan against vacuum
    if (rhoR == 0.) {
        side = -1;
    } else if (rhoL == 0.) {
        side = 1;
    } else {
        double SR = uR - c.tdgm1 * aR;
        double SL = uL + c.tdgm1 * aL;
        if (SR > dxdt && SL < dxdt) {
            *rho = 0.; *u = 0.; *P = 0.;
            return 0;
        }
        side = (SL < dxdt) ? 1 : -1;
    }
    if (side == -1) {
        if (uL - aL < dxdt) {
            double SL = uL + c.tdgm1 * aL;
            if (SL > dxdt) {
                double base = c.tdgp1 + c.gm1dgp1 * (uL - dxdt) / aL;
                *rho = rhoL * pow(base, c.tdgm1);
                *u = c.tdgp1 * (aL + c.gm1d2 * uL + dxdt);
                *P = PL * pow(base, c.tgdgm1);
                return -1;
            }
            *rho = 0.; *u = 0.; *P = 0.;
            return 0;
        }
        *rho = rhoL; *u = uL; *P = PL;
        return -1;
    }
    if (dxdt < uR + aR) {
        double SR = uR - c.tdgm1 * aR;
        if (SR < dxdt) {
            double base = c.tdgp1 - c.gm1dgp1 * (uR - dxdt) / aR;
            *rho = rhoR * pow(base, c.tdgm1);
            *u = c.tdgp1 * (-aR + c.gm1d2 * uR + dxdt);
            *P = PR * pow(base, c.tgdgm1);
            return 1;
        }
        *rho = 0.; *u = 0.; *P = 0.;
        return 0;
    }
    *rho = rhoR; *u = uR; *P = PR;
    return 1;
}

// returns +1 (right of the contact sampled), -1 (left), 0 (vacuum); Riemann.cpp:93-127
__device__ __forceinline__ int rs_solve(const RsConsts &c, double rhoL, double uL, double PL, double rhoR, double uR,
                                        double PR, double *rhosol, double *usol, double *Psol) {
    const double dxdt = 0.;
    if (rhoL == 0. || rhoR == 0.) return rs_solve_vacuum(c, rhoL, uL, PL, rhoR, uR, PR, rhosol, usol, Psol);
    const double aL = sqrt(c.gamma * PL / rhoL);
    const double aR = sqrt(c.gamma * PR / rhoR);
    if (c.tdgm1 * (aL + aR) <= uR - uL) return rs_solve_vacuum(c, rhoL, uL, PL, rhoR, uR, PR, rhosol, usol, Psol);
    double Pstar = 0.;
    double Pguess = rs_guess_P(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR);
    double fPstar = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pstar);
    double fPguess = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pguess);
    if (fPstar * fPguess >= 0.) {
        while (fabs(Pstar - Pguess) > 5.e-9 * (Pstar + Pguess) && fPguess < 0.) {
            Pstar = Pguess;
            fPstar = fPguess;
            Pguess = Pguess - fPguess / (rs_fprimeb(c, rhoL, PL, aL, Pguess) + rs_fprimeb(c, rhoR, PR, aR, Pguess));
            fPguess = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pguess);
        }
    }
    if (1.e6 * fabs(Pstar - Pguess) > 0.5 * (Pstar + Pguess) && fPguess > 0.) {
        Pstar = rs_brent(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pstar, Pguess, fPstar, fPguess);
    } else {
        Pstar = Pguess;
    }
    const double ustar = 0.5 * (uL + uR) + 0.5 * (rs_fb(c, rhoR, PR, aR, Pstar) - rs_fb(c, rhoL, PL, aL, Pstar));
    if (ustar < dxdt) {
        if (Pstar > PR) { // right shock
            double PdPR = Pstar / PR;
            double SR = uR + aR * sqrt(c.gp1d2g * PdPR + c.gm1d2g);
            if (SR > dxdt) {
                *rhosol = rhoR * (PdPR + c.gm1dgp1) / (c.gm1dgp1 * PdPR + 1.);
                *usol = ustar;
                *Psol = Pstar;
            } else {
                *rhosol = rhoR; *usol = uR; *Psol = PR;
            }
        } else { // right rarefaction
            double SHR = uR + aR;
            if (SHR > dxdt) {
                double PdPR = Pstar / PR;
                double STR = ustar + aR * pow(PdPR, c.gm1d2g);
                if (STR > dxdt) {
                    *rhosol = rhoR * pow(PdPR, c.ginv);
                    *usol = ustar;
                    *Psol = Pstar;
                } else {
                    double base = c.tdgp1 - c.gm1dgp1 * (uR - dxdt) / aR;
                    *rhosol = rhoR * pow(base, c.tdgm1);
                    *usol = c.tdgp1 * (-aR + c.gm1d2 * uR + dxdt);
                    *Psol = PR * pow(base, c.tgdgm1);
                }
            } else {
                *rhosol = rhoR; *usol = uR; *Psol = PR;
            }
        }
        return 1;
    }
    if (Pstar > PL) { // left shock
        double PdPL = Pstar / PL;
        double SL = uL - aL * sqrt(c.gp1d2g * PdPL + c.gm1d2g);
        if (SL < dxdt) {
            *rhosol = rhoL * (PdPL + c.gm1dgp1) / (c.gm1dgp1 * PdPL + 1.);
            *usol = ustar;
            *Psol = Pstar;
        } else {
            *rhosol = rhoL; *usol = uL; *Psol = PL;
        }
    } else { // left rarefaction
        double SHL = uL - aL;
        if (SHL < dxdt) {
            double PdPL = Pstar / PL;
            double STL = ustar - aL * pow(PdPL, c.gm1d2g);
            if (STL > dxdt) {
                double base = c.tdgp1 + c.gm1dgp1 * (uL - dxdt) / aL;
                *rhosol = rhoL * pow(base, c.tdgm1);
                *usol = c.tdgp1 * (aL + c.gm1d2 * uL + dxdt);
                *Psol = PL * pow(base, c.tgdgm1);
            } else {
                *rhosol = rhoL * pow(PdPL, c.ginv);
                *usol = ustar;
                *Psol = Pstar;
            }
        } else {
            *rhosol = rhoL; *usol = uL; *Psol = PL;
        }
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// Particles::pairwiseLimiter, Particles.cpp:1735-1785 (quirk Q1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pairwise_limiter(const Params &p, double phi0, double phi_i, double phi_j,
                                                   double xijxi_abs, double xjxi_abs) {
    const int am = p.abs_mode;
    double phi_ = phi_i;
    double phi_ij = phi_i + xijxi_abs / xjxi_abs * (phi_j - phi_i);
    double phiMin, phiMax;
    if (phi_i < phi_j) {
        phiMin = phi_i;
        phiMax = phi_j;
    } else {
        phiMin = phi_j;
        phiMax = phi_i;
    }
    double delta1 = p.psi1 * q1_abs(phi_i - phi_j, am);
    double delta2 = p.psi2 * q1_abs(phi_i - phi_j, am);
    double phiMinus, phiPlus;
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
        double minPhiD2 = (phi_ij + delta2 < phi0) ? phi_ij + delta2 : phi0;
        phi_ = phiMinus > minPhiD2 ? phiMinus : minPhiD2;
    } else if (phi_i > phi_j) {
        double maxPhiD2 = (phi_ij - delta2 > phi0) ? phi_ij - delta2 : phi0;
        phi_ = phiPlus < maxPhiD2 ? phiPlus : maxPhiD2;
    }
    return phi_;
}

template <int D>
__device__ __forceinline__ double dotD(const double *a, const double *b) { // Helper::dotProduct
    double res = 0.;
#pragma unroll
    for (int k = 0; k < D; ++k) res += a[k] * b[k];
    return res;
}

// Riemann::Riemann + exact + rotateAndProjectFluxes{2D,3D} (Riemann.cpp:7-229).  Wa = state of the
// canonical particle ("WijR" of the caller = class member WL, the LEFT state of the solver), Wb = the
// neighbour's ("WijL" = class WR, RIGHT state): quirk Q5.  W = [rho, P, vx, vy(, vz)].
template <int D>
__device__ __forceinline__ void face_flux(const Params &p, double *Wa, double *Wb, const double *vFrame, const double *A,
                                          double *F) {
    const double gamma = p.gamma;
    const double AijNorm = sqrt(dotD<D>(A, A));
    double hatA[D];
#pragma unroll
    for (int k = 0; k < D; ++k) hatA[k] = 1. / AijNorm * A[k];
    double rhoSol, PSol, vSol[D];
#pragma unroll
    for (int k = 0; k < D; ++k) vSol[k] = 0.;
    int flag;
    if (D == 2) {
        // rotationMatrix2D(hatA, unitX): Lambda = [[ax, ay], [-ay, ax]] in the reference's arithmetic
        double L0 = hatA[0] * 1. + hatA[1] * 0.;
        double L1 = -(hatA[0] * 0. - hatA[1] * 1.);
        double L2 = -L1, L3 = L0;
        double bR0 = Wb[2], bR1 = Wb[3], bL0 = Wa[2], bL1 = Wa[3];
        Wb[2] = L0 * bR0 + L1 * bR1;
        Wb[3] = L2 * bR0 + L3 * bR1;
        Wa[2] = L0 * bL0 + L1 * bL1;
        Wa[3] = L2 * bL0 + L3 * bL1;
        flag = rs_solve(p.rs, Wa[0], Wa[2], Wa[1], Wb[0], Wb[2], Wb[1], &rhoSol, &vSol[0], &PSol);
        if (flag == 1)
            vSol[1] = Wb[3];
        else if (flag == -1)
            vSol[1] = Wa[3];
        // rotationMatrix2D(unitX, hatA)
        double I0 = 1. * hatA[0] + 0. * hatA[1];
        double I1 = -(1. * hatA[1] - 0. * hatA[0]);
        double I2 = -I1, I3 = I0;
        double s0 = vSol[0], s1 = vSol[1];
        vSol[0] = I0 * s0 + I1 * s1;
        vSol[1] = I2 * s0 + I3 * s1;
        F[0] = A[0] * rhoSol * vSol[0] + A[1] * rhoSol * vSol[1];
        double vLab[2] = {vSol[0] + vFrame[0], vSol[1] + vFrame[1]};
        if (p.mfm) {
            vSol[0] = 0.;
            vSol[1] = 0.;
        }
        F[2] = A[0] * (rhoSol * vLab[0] * vSol[0] + PSol) + A[1] * rhoSol * vLab[0] * vSol[1];
        F[3] = A[0] * rhoSol * vLab[1] * vSol[0] + A[1] * (rhoSol * vLab[1] * vSol[1] + PSol);
        F[1] = A[0] * (vSol[0] * (PSol / (gamma - 1.) + rhoSol * .5 * dotD<2>(vLab, vLab)) + PSol * vLab[0]) +
               A[1] * (vSol[1] * (PSol / (gamma - 1.) + rhoSol * .5 * dotD<2>(vLab, vLab)) + PSol * vLab[1]);
    } else {
        // rotationMatrix3D(a = hatA, b = unitX): v = a x b, Rodrigues with n = 1/(1+cos) (singular for hatA = -x)
        double L[9], Li[9];
        {
            const double a0 = hatA[0], a1 = hatA[1], a2 = hatA[2];
            double v0 = a1 * 0. - a2 * 0.;
            double v1 = a2 * 1. - a0 * 0.;
            double v2 = a0 * 0. - a1 * 1.;
            double cosAB = 0. + a0 * 1. + a1 * 0. + a2 * 0.;
            double n = 1. / (1. + cosAB);
            L[0] = 1. - n * (v2 * v2 + v1 * v1);
            L[1] = -v2 + n * v0 * v1;
            L[2] = v1 + n * v0 * v2;
            L[3] = v2 + n * v0 * v1;
            L[4] = 1. - n * (v2 * v2 + v0 * v0);
            L[5] = -v0 + n * v1 * v2;
            L[6] = -v1 + n * v0 * v2;
            L[7] = v0 + n * v1 * v2;
            L[8] = 1. - n * (v1 * v1 + v0 * v0);
            // rotationMatrix3D(a = unitX, b = hatA)
            double w0 = 0. * a2 - 0. * a1;
            double w1 = 0. * a0 - 1. * a2;
            double w2 = 1. * a1 - 0. * a0;
            double cosBA = 0. + 1. * a0 + 0. * a1 + 0. * a2;
            double m = 1. / (1. + cosBA);
            Li[0] = 1. - m * (w2 * w2 + w1 * w1);
            Li[1] = -w2 + m * w0 * w1;
            Li[2] = w1 + m * w0 * w2;
            Li[3] = w2 + m * w0 * w1;
            Li[4] = 1. - m * (w2 * w2 + w0 * w0);
            Li[5] = -w0 + m * w1 * w2;
            Li[6] = -w1 + m * w0 * w2;
            Li[7] = w0 + m * w1 * w2;
            Li[8] = 1. - m * (w1 * w1 + w0 * w0);
        }
        double bR[3] = {Wb[2], Wb[3], Wb[4]}, bL[3] = {Wa[2], Wa[3], Wa[4]};
        Wb[2] = L[0] * bR[0] + L[1] * bR[1] + L[2] * bR[2];
        Wb[3] = L[3] * bR[0] + L[4] * bR[1] + L[5] * bR[2];
        Wb[4] = L[6] * bR[0] + L[7] * bR[1] + L[8] * bR[2];
        Wa[2] = L[0] * bL[0] + L[1] * bL[1] + L[2] * bL[2];
        Wa[3] = L[3] * bL[0] + L[4] * bL[1] + L[5] * bL[2];
        Wa[4] = L[6] * bL[0] + L[7] * bL[1] + L[8] * bL[2];
        flag = rs_solve(p.rs, Wa[0], Wa[2], Wa[1], Wb[0], Wb[2], Wb[1], &rhoSol, &vSol[0], &PSol);
        if (flag == 1) {
            vSol[1] = Wb[3];
            vSol[2] = Wb[4];
        } else if (flag == -1) {
            vSol[1] = Wa[3];
            vSol[2] = Wa[4];
        }
        double s[3] = {vSol[0], vSol[1], vSol[2]};
        vSol[0] = Li[0] * s[0] + Li[1] * s[1] + Li[2] * s[2];
        vSol[1] = Li[3] * s[0] + Li[4] * s[1] + Li[5] * s[2];
        vSol[2] = Li[6] * s[0] + Li[7] * s[1] + Li[8] * s[2];
        F[0] = A[0] * rhoSol * vSol[0] + A[1] * rhoSol * vSol[1] + A[2] * rhoSol * vSol[2];
        double vLab[3] = {vSol[0] + vFrame[0], vSol[1] + vFrame[1], vSol[2] + vFrame[2]};
        if (p.mfm) {
            vSol[0] = 0.;
            vSol[1] = 0.;
            vSol[2] = 0.;
        }
        F[2] = A[0] * (rhoSol * vLab[0] * vSol[0] + PSol) + A[1] * rhoSol * vLab[0] * vSol[1] + A[2] * rhoSol * vLab[0] * vSol[2];
        F[3] = A[0] * rhoSol * vLab[1] * vSol[0] + A[1] * (rhoSol * vLab[1] * vSol[1] + PSol) + A[2] * rhoSol * vLab[1] * vSol[2];
        F[4] = A[0] * rhoSol * vLab[2] * vSol[0] + A[1] * rhoSol * vLab[2] * vSol[1] + A[2] * (rhoSol * vLab[2] * vSol[2] + PSol);
        F[1] = A[0] * (vSol[0] * (PSol / (gamma - 1.) + rhoSol * .5 * dotD<3>(vLab, vLab)) + PSol * vLab[0]) +
               A[1] * (vSol[1] * (PSol / (gamma - 1.) + rhoSol * .5 * dotD<3>(vLab, vLab)) + PSol * vLab[1]) +
               A[2] * (vSol[2] * (PSol / (gamma - 1.) + rhoSol * .5 * dotD<3>(vLab, vLab)) + PSol * vLab[2]);
    }
    if (flag == 0) atomicOr(p.d.flags, MLH_F_VACUUM);
}

// device-side dt policy (MeshlessScheme.cpp:91-105): fixed dt, or CFL dt clipped to dt_max
__global__ void k_select_dt(const Params p, double dt_fixed, double dt_max) {
    double dt;
    if (dt_fixed > 0.) {
        dt = dt_fixed;
    } else {
        dt = __longlong_as_double((long long)*p.d.dt_bits);
        if (dt_max > 0. && dt > dt_max) dt = dt_max;
    }
    *p.d.dt_used = dt;
}

// W component nu -> gradient field slot: W = [rho, P, vx, vy, vz], slots rho 0, vx 1, vy 2, vz 3, P 4
__device__ __forceinline__ int w2f(int nu) { return nu == 0 ? 0 : (nu == 1 ? 4 : nu - 1); }

template <int D, bool PER>
__global__ void __launch_bounds__(128) k_flux_update(const Params p) {
    constexpr int NW = D + 2;
    const int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.own_end) return;
    const double dt = *p.d.dt_used;
    const double gamma = p.gamma;

    // ---- own bundle ----
    double xs[D], vs[D], Bs[D * D], gs[NW][D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        xs[k] = p.d.x[k][i];
        vs[k] = p.d.v[k][i];
    }
#pragma unroll
    for (int k = 0; k < D * D; ++k) Bs[k] = p.d.B[k][i];
#pragma unroll
    for (int nu = 0; nu < NW; ++nu)
#pragma unroll
        for (int k = 0; k < D; ++k) gs[nu][k] = p.d.g[w2f(nu) * 3 + k][i];
    const double rhos = p.d.rho[i], Ps = p.d.P[i], omgs = p.d.omega[i];
    const int ids = p.d.id[i];
    const int nreg = p.d.noi[i], ntot = nreg + p.d.noig[i];

    double acc[NW];
#pragma unroll
    for (int nu = 0; nu < NW; ++nu) acc[nu] = 0.;

    for (int s = 0; s < ntot; ++s) {
        const int e = p.d.nnl[(size_t)s * p.ncap + i];
        const int j = e & MLH_NNL_IDX_MASK;
        const int code = PER ? (int)((unsigned)e >> MLH_NNL_IDX_BITS) : 0;
        const int idn = p.d.id[j];
        // canonical orientation: the endpoint with the lower ORIGINAL index plays "i" (Particles.cpp:1841,1889)
        const bool canon = !(idn < ids);
        // ---- a = canonical endpoint, b = the other; operands selected, not branched ----
        double xa[D], xb[D], va[D], vb[D], Ba[D * D], Bb[D * D], ga[NW][D], gb[NW][D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double xn = p.d.x[k][j], vn = p.d.v[k][j];
            xa[k] = canon ? xs[k] : xn;
            xb[k] = canon ? xn : xs[k];
            va[k] = canon ? vs[k] : vn;
            vb[k] = canon ? vn : vs[k];
        }
#pragma unroll
        for (int k = 0; k < D * D; ++k) {
            const double bn = p.d.B[k][j];
            Ba[k] = canon ? Bs[k] : bn;
            Bb[k] = canon ? bn : Bs[k];
        }
#pragma unroll
        for (int nu = 0; nu < NW; ++nu)
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double gn = p.d.g[w2f(nu) * 3 + k][j];
                ga[nu][k] = canon ? gs[nu][k] : gn;
                gb[nu][k] = canon ? gn : gs[nu][k];
            }
        const double rhon = p.d.rho[j], Pn = p.d.P[j], omgn = p.d.omega[j];
        const double rhoa = canon ? rhos : rhon, rhob = canon ? rhon : rhos;
        const double Pa = canon ? Ps : Pn, Pb = canon ? Pn : Ps;
        const double omga = canon ? omgs : omgn, omgb = canon ? omgn : omgs;

        // ---- geometry: b's image as a sees it, a's image as b sees it (identity for regular pairs) ----
        double xbi[D], xai[D];
        if (PER && code != 0) {
            const int cab = canon ? code : reverse_code(code); // code of b's image in a's list
            const int cba = reverse_code(cab);
            bool ex = true;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xbi[k] = image_coord(xb[k], (cab >> (2 * k)) & 3, p.grid.bmin[k], p.grid.bmax[k]);
                xai[k] = image_coord(xa[k], (cba >> (2 * k)) & 3, p.grid.bmin[k], p.grid.bmax[k]);
            }
            // quirk Q9: is the pair also in the OTHER particle's list?  (the view that is not ours)
            {
                const int cview = reverse_code(code); // image of self as the neighbour sees it
                double dd[3];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int ck = (cview >> (2 * k)) & 3;
                    ex = ex && image_exists(xs[k], ck, p.grid.bmin[k], p.grid.bmax[k], p.h);
                    dd[k] = __dsub_rn(image_coord(xs[k], ck, p.grid.bmin[k], p.grid.bmax[k]), p.d.x[k][j]);
                }
                ex = ex && (dist_sqr_exact<D>(dd) < p.hSqr);
                if (!ex && !p.symmetric_seam) atomicAdd(&p.d.counters[0], 1u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xbi[k] = xb[k];
                xai[k] = xa[k];
            }
        }

        // ---- effective face A_ab = psi~_b(x_a)/omega_a - psi~_a(x_b)/omega_b (Particles.cpp:1299-1302, :2525-2528) ----
        double A[D];
        {
            double s1[3], s2[3], d1[D], d2[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                s1[k] = __dsub_rn(xa[k], xbi[k]);
                d1[k] = __dsub_rn(xbi[k], xa[k]);
                s2[k] = __dsub_rn(xb[k], xai[k]);
                d2[k] = __dsub_rn(xai[k], xb[k]);
            }
            const double r1 = sqrt(dist_sqr_exact<D>(s1));
            const double r2 = (PER && code != 0) ? sqrt(dist_sqr_exact<D>(s2)) : r1;
            const double w1 = cubic_spline(r1, p);
            const double w2 = (PER && code != 0) ? cubic_spline(r2, p) : w1;
            const double psi1 = w1 / omga, psi2 = w2 / omgb;
#pragma unroll
            for (int al = 0; al < D; ++al) {
                double t1 = 0., t2 = 0.;
#pragma unroll
                for (int be = 0; be < D; ++be) {
                    t1 += Ba[D * al + be] * d1[be] * psi1;
                    t2 += Bb[D * al + be] * d2[be] * psi2;
                }
                A[al] = 1. / omga * t1 - 1. / omgb * t2;
            }
        }

        // ---- boosted, reconstructed, predicted states (Particles.cpp:1498-1721; ghosts :2546-2672) ----
        double xjxi[3], xijxi[D], xijxj[D], vF[D], Wa[NW], Wb[NW];
        xjxi[2] = 0.; // quirk Q13 (ZERO_Z): never written in the first-order 3D branch
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (k < 2 || p.q13_mode == MLH_Q13_GEOMETRIC) xjxi[k] = xbi[k] - xa[k];
            xijxj[k] = .5 * (xa[k] - xbi[k]);
            xijxi[k] = .5 * (xbi[k] - xa[k]);
            vF[k] = p.move_particles ? (va[k] + vb[k]) / 2. : 0.;
        }
        Wa[0] = rhoa;
        Wb[0] = rhob;
        Wa[1] = Pa;
        Wb[1] = Pb;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            Wa[2 + k] = va[k] - vF[k];
            Wb[2 + k] = vb[k] - vF[k];
        }
        double Wa0[NW], Wb0[NW];
#pragma unroll
        for (int nu = 0; nu < NW; ++nu) {
            Wa0[nu] = Wa[nu];
            Wb0[nu] = Wb[nu];
        }
#pragma unroll
        for (int nu = 0; nu < NW; ++nu) {
            Wa[nu] += dotD<D>(ga[nu], xijxi);
            Wb[nu] += dotD<D>(gb[nu], xijxj);
        }
        if (p.pairwise && code == 0) { // the ghost overload has no pairwise limiter (:2632-2644)
            double na = 0., nb = 0., nab = 0.;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                na += xijxi[k] * xijxi[k];
                nb += xijxj[k] * xijxj[k];
                nab += xjxi[k] * xjxi[k];
            }
            na = sqrt(na);
            nb = sqrt(nb);
            nab = sqrt(nab);
#pragma unroll
            for (int nu = 0; nu < NW; ++nu) {
                const double wa = pairwise_limiter(p, Wa[nu], Wa0[nu], Wb0[nu], na, nab);
                const double wb = pairwise_limiter(p, Wb[nu], Wb0[nu], Wa0[nu], nb, nab);
                Wa[nu] = wa;
                Wb[nu] = wb;
            }
        }
        {
            // gradient rows: [0] rho, [1] P, [2] vx, [3] vy, [4] vz
            double aDiv = ga[2][0] + ga[3][1];
            double bDiv = gb[2][0] + gb[3][1];
            if (D == 3) {
                aDiv += ga[NW - 1][D - 1];
                bDiv += gb[NW - 1][D - 1];
            }
            const double wa0 = va[0] - vF[0], wa1 = va[1] - vF[1];
            const double wb0 = vb[0] - vF[0], wb1 = vb[1] - vF[1];
            Wa[0] -= dt / 2. * (rhoa * aDiv + wa0 * ga[0][0] + wa1 * ga[0][1]);
            Wb[0] -= dt / 2. * (rhob * bDiv + wb0 * gb[0][0] + wb1 * gb[0][1]);
            Wa[1] -= dt / 2. * (gamma * Pa * aDiv + wa0 * ga[1][0] + wa1 * ga[1][1]);
            Wb[1] -= dt / 2. * (gamma * Pb * bDiv + wb0 * gb[1][0] + wb1 * gb[1][1]);
            Wa[2] -= dt / 2. * (ga[1][0] / rhoa + wa0 * ga[2][0] + wa1 * ga[2][1]);
            Wb[2] -= dt / 2. * (gb[1][0] / rhob + wb0 * gb[2][0] + wb1 * gb[2][1]);
            Wa[3] -= dt / 2. * (ga[1][1] / rhoa + wa0 * ga[3][0] + wa1 * ga[3][1]);
            Wb[3] -= dt / 2. * (gb[1][1] / rhob + wb0 * gb[3][0] + wb1 * gb[3][1]);
            if (D == 3) {
                const double wa2 = va[D - 1] - vF[D - 1], wb2 = vb[D - 1] - vF[D - 1];
                const double wq3 = (p.q3_mode == MLH_Q3_FIXED) ? wb2 : wa2; // quirk Q3 (:1717,:1719)
                Wa[0] -= dt / 2. * wa2 * ga[0][D - 1];
                Wb[0] -= dt / 2. * wb2 * gb[0][D - 1];
                Wa[1] -= dt / 2. * wa2 * ga[1][D - 1];
                Wb[1] -= dt / 2. * wb2 * gb[1][D - 1];
                Wa[2] -= dt / 2. * wa2 * ga[2][D - 1];
                Wb[2] -= dt / 2. * wq3 * gb[2][D - 1];
                Wa[3] -= dt / 2. * wa2 * ga[3][D - 1];
                Wb[3] -= dt / 2. * wq3 * gb[3][D - 1];
                Wa[NW - 1] -= dt / 2. * (ga[1][D - 1] / rhoa + wa0 * ga[NW - 1][0] + wa1 * ga[NW - 1][1] + wa2 * ga[NW - 1][D - 1]);
                Wb[NW - 1] -= dt / 2. * (gb[1][D - 1] / rhob + wb0 * gb[NW - 1][0] + wb1 * gb[NW - 1][1] + wb2 * gb[NW - 1][D - 1]);
            }
        }
        if (PER && code != 0 && (Wa[1] < 0. || Wb[1] < 0.)) atomicOr(p.d.flags, MLH_F_NEG_GHOST_PRESSURE);

        // ---- one exact Riemann problem along A, fluxes projected on A ----
        double F[NW];
        face_flux<D>(p, Wa, Wb, vF, A, F);
        const double sgn = canon ? 1. : -1.;
#pragma unroll
        for (int nu = 0; nu < NW; ++nu) acc[nu] += sgn * F[nu]; // collectFluxes, :1926-2008
    }

    if (p.debug_capture) {
        p.d.flux[0][i] = acc[0];
        p.d.flux[1][i] = acc[1];
#pragma unroll
        for (int k = 0; k < D; ++k) p.d.flux[2 + k][i] = acc[2 + k];
    }

    // ---- K5: updateStateAndPosition, Particles.cpp:2013-2110 ----
    {
        double m = p.d.m[i], u = p.d.u[i];
        double Q[D + 1];
        double v2 = 0.;
        if (D == 3)
            v2 = vs[0] * vs[0] + vs[1] * vs[1] + vs[D - 1] * vs[D - 1];
        else
            v2 = vs[0] * vs[0] + vs[1] * vs[1];
        Q[0] = m * (u + .5 * v2);
#pragma unroll
        for (int k = 0; k < D; ++k) Q[1 + k] = m * vs[k];
        if (!p.mfm) m -= dt * acc[0];
        double vn[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            Q[1 + k] -= dt * acc[2 + k];
            vn[k] = Q[1 + k] / m;
        }
        Q[0] -= dt * acc[1];
        if (D == 3)
            v2 = vn[0] * vn[0] + vn[1] * vn[1] + vn[D - 1] * vn[D - 1];
        else
            v2 = vn[0] * vn[0] + vn[1] * vn[1];
        u = Q[0] / m - .5 * v2;
        const int o = i - p.own_begin;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double x = xs[k];
            if (p.move_particles) {
                x += vs[k] * dt;
                if (PER) {
                    if (x < p.grid.bmin[k]) {
                        x = p.grid.bmax[k] - (p.grid.bmin[k] - x);
                    } else if (p.grid.bmax[k] <= x) {
                        x = p.grid.bmin[k] + (x - p.grid.bmax[k]);
                    }
                }
            }
            p.d.cx[k][o] = x;
            p.d.cv[k][o] = vn[k];
        }
        p.d.cm[o] = m;
        p.d.cu[o] = u;
        p.d.cid[o] = ids;
    }
}

} // namespace

int mlh_launch_flux(mlh_ctx *c, double dt_fixed, double dt_max) {
    Params &p = c->p;
    int n = p.own_end - p.own_begin;
    mlh_prof_begin(c, KID_SELECT_DT);
    k_select_dt<<<1, 1, 0, c->stream>>>(p, dt_fixed, dt_max);
    mlh_prof_end(c, KID_SELECT_DT);
    mlh_prof_begin(c, KID_FLUX);
    if (p.D == 2 && p.periodic)
        k_flux_update<2, true><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else if (p.D == 2)
        k_flux_update<2, false><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else if (p.periodic)
        k_flux_update<3, true><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else
        k_flux_update<3, false><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_FLUX);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    p.ncur = n;
    return MLH_OK;
}
