/*
 * oracle/riemann_exact.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Restatement of the exact ideal-gas Riemann solver the reference calls at
 *   /root/reference/demonstrator/src/Riemann.cpp:84      (RiemannSolver solver{gamma})
 *   /root/reference/demonstrator/src/Riemann.cpp:93-94   (solver.solve(rhoL,uL,PL,rhoR,uR,PR,rho,u,P))
 *   /root/reference/demonstrator/src/Riemann.cpp:104-127 (meaning of the returned flag: +1 right, -1 left, 0 vacuum)
 *
 * PARITY UNPINNED: the reference does not vendor this solver.  It is the
 * third-party header `RiemannSolver.hpp` of github.com/bwvdnbro/python_finite_volume_solver,
 * fetched from the `master` branch (no pinned version) by
 * /root/reference/demonstrator/Makefile:93-94, and absent from /root/reference and
 * from this image (no network).  What follows restates its published algorithm
 * (Toro, "Riemann Solvers and Numerical Methods for Fluid Dynamics", ch. 4:
 * Newton-Raphson on the pressure function with the adaptive PVRS / two-rarefaction /
 * two-shock initial guess, Brent fallback, sampling at x/t = dxdt) and is THE
 * definition the oracle, the reference-source build (oracle/_ref) and the CUDA
 * kernels are all held to.  The reference has no test or golden vector for it.
 * What IS pinned (tests/test_riemann_kat.py): the SOLUTION -- Toro's tables 4.1-4.3 to five digits, and an
 * independent 40-digit restatement of the exact solution (mpmath, from the book) on 1500 random states to <= 1e-10
 * (measured 1.5e-12).  What stays unpinned: the third-party code's iteration path inside its 5e-9 stopping rule.
 *
 * Plain C99 (also valid C++), header-only.
 */
#ifndef MLH_ORACLE_RIEMANN_EXACT_H
#define MLH_ORACLE_RIEMANN_EXACT_H

/* In C++ use <cmath>, NOT the C++ <math.h> wrapper: the wrapper does `using std::abs`,
 * which would silently turn the reference's unqualified abs(double) calls
 * (Particles.cpp:1416-1417,1748-1759) from `int abs(int)` into fabs. */
#ifdef __cplusplus
#include <cmath>
#else
#include <math.h>
#endif

typedef struct {
    double gamma;
    double gp1d2g;  /* (gamma+1)/(2 gamma) */
    double gm1d2g;  /* (gamma-1)/(2 gamma) */
    double gm1dgp1; /* (gamma-1)/(gamma+1) */
    double tdgp1;   /* 2/(gamma+1) */
    double tdgm1;   /* 2/(gamma-1) */
    double gm1d2;   /* (gamma-1)/2 */
    double tgdgm1;  /* 2 gamma/(gamma-1) */
    double ginv;    /* 1/gamma */
} rs_consts;

static inline void rs_init(rs_consts *c, double gamma) {
    c->gamma = gamma;
    c->gp1d2g = 0.5 * (gamma + 1.) / gamma;
    c->gm1d2g = 0.5 * (gamma - 1.) / gamma;
    c->gm1dgp1 = (gamma - 1.) / (gamma + 1.);
    c->tdgp1 = 2. / (gamma + 1.);
    c->tdgm1 = 2. / (gamma - 1.);
    c->gm1d2 = 0.5 * (gamma - 1.);
    c->tgdgm1 = 2. * gamma / (gamma - 1.);
    c->ginv = 1. / gamma;
}

/* std::max / std::min semantics (returns first argument unless strictly ordered) */
static inline double rs_max(double a, double b) { return (a < b) ? b : a; }
static inline double rs_min(double a, double b) { return (b < a) ? b : a; }

static inline double rs_soundspeed(const rs_consts *c, double rho, double P) {
    return sqrt(c->gamma * P / rho);
}

/* Toro eq. (4.6)/(4.7): pressure function of one side */
static inline double rs_fb(const rs_consts *c, double rho, double P, double a, double Pstar) {
    double fval;
    if (Pstar > P) {
        double A = c->tdgp1 / rho;
        double B = c->gm1dgp1 * P;
        fval = (Pstar - P) * sqrt(A / (Pstar + B));
    } else {
        fval = c->tdgm1 * a * (pow(Pstar / P, c->gm1d2g) - 1.);
    }
    return fval;
}

static inline double rs_f(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                          double rhoR, double uR, double PR, double aR, double Pstar) {
    return rs_fb(c, rhoL, PL, aL, Pstar) + rs_fb(c, rhoR, PR, aR, Pstar) + (uR - uL);
}

/* derivative of the pressure function of one side (Toro eq. 4.37) */
static inline double rs_fprimeb(const rs_consts *c, double rho, double P, double a, double Pstar) {
    double fval;
    if (Pstar > P) {
        double A = c->tdgp1 / rho;
        double B = c->gm1dgp1 * P;
        fval = (1. - 0.5 * (Pstar - P) / (B + Pstar)) * sqrt(A / (Pstar + B));
    } else {
        fval = 1. / (rho * a) * pow(Pstar / P, -c->gp1d2g);
    }
    return fval;
}

static inline double rs_fprime(const rs_consts *c, double rhoL, double PL, double aL,
                               double rhoR, double PR, double aR, double Pstar) {
    return rs_fprimeb(c, rhoL, PL, aL, Pstar) + rs_fprimeb(c, rhoR, PR, aR, Pstar);
}

static inline double rs_gb(const rs_consts *c, double rho, double P, double Pstar) {
    double A = c->tdgp1 / rho;
    double B = c->gm1dgp1 * P;
    return sqrt(A / (Pstar + B));
}

/* Toro section 4.3.2: adaptive initial guess, floored at 5e-9 (PL+PR) */
static inline double rs_guess_P(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                                double rhoR, double uR, double PR, double aR) {
    double Pguess;
    double Pmin = rs_min(PL, PR);
    double Pmax = rs_max(PL, PR);
    double qmax = Pmax / Pmin;
    double Ppv = 0.5 * (PL + PR) - 0.125 * (uR - uL) * (PL + PR) * (aL + aR);
    Ppv = rs_max(5.e-9 * (PL + PR), Ppv);
    if (qmax <= 2. && Pmin <= Ppv && Ppv <= Pmax) {
        Pguess = Ppv;
    } else {
        if (Ppv < Pmin) {
            /* two rarefactions */
            Pguess = pow((aL + aR - c->gm1d2 * (uR - uL)) /
                             (aL / pow(PL, c->gm1d2g) + aR / pow(PR, c->gm1d2g)),
                         c->tgdgm1);
        } else {
            /* two shocks */
            double gL = rs_gb(c, rhoL, PL, Ppv);
            double gR = rs_gb(c, rhoR, PR, Ppv);
            Pguess = (gL * PL + gR * PR - uR + uL) / (gL + gR);
        }
    }
    Pguess = rs_max(5.e-9 * (PL + PR), Pguess);
    return Pguess;
}

#ifdef RS_STATS
/* optional instrumentation (make RS_STATS=1): histogram of root-finder iterations per solve */
static long rs_stat_newton[64], rs_stat_brent[64];
#define RS_STAT_ADD(h, n) ((h)[(n) < 63 ? (n) : 63]++)
#else
#define RS_STAT_ADD(h, n) ((void)0)
#endif

/* Brent's method on [lower, upper] with f(lower)*f(upper) < 0, relative tolerance 5e-9*(a+b) */
static inline double rs_brent(const rs_consts *cst, double rhoL, double uL, double PL, double aL,
                              double rhoR, double uR, double PR, double aR,
                              double lowerlimit, double upperlimit, double lowf, double upf) {
    double a = lowerlimit, b = upperlimit, c = 0., d = 1e230;
    double fa = lowf, fb = upf, fc = 0., s = 0., fs = 0.;
    int mflag;
    int nit = 0;
    if (fa * fb > 0.) {
        return b; /* not bracketed: caller's precondition violated, keep upper */
    }
    if (fabs(fa) < fabs(fb)) {
        double t = a; a = b; b = t;
        t = fa; fa = fb; fb = t;
    }
    c = a;
    fc = fa;
    mflag = 1;
    while (!(fb == 0.) && (fabs(a - b) > 5.e-9 * (a + b))) {
        if ((fa != fc) && (fb != fc)) {
            /* inverse quadratic interpolation */
            s = a * fb * fc / (fa - fb) / (fa - fc) + b * fa * fc / (fb - fa) / (fb - fc) +
                c * fa * fb / (fc - fa) / (fc - fb);
        } else {
            /* secant rule */
            s = b - fb * (b - a) / (fb - fa);
        }
        {
            double tmp2 = 0.25 * (3. * a + b);
            if (!(((s > tmp2) && (s < b)) || ((s < tmp2) && (s > b))) ||
                (mflag && (fabs(s - b) >= (0.5 * fabs(b - c)))) ||
                (!mflag && (fabs(s - b) >= (0.5 * fabs(c - d)))) ||
                (mflag && (fabs(b - c) < 5.e-9 * (b + c))) ||
                (!mflag && (fabs(c - d) < 5.e-9 * (c + d)))) {
                s = 0.5 * (a + b);
                mflag = 1;
            } else {
                mflag = 0;
            }
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
        ++nit;
    }
    RS_STAT_ADD(rs_stat_brent, nit);
    (void)nit;
    return b;
}

/* ---- sampling (Toro section 4.5) ---- */
static inline void rs_sample_right_shock(const rs_consts *c, double rhoR, double uR, double PR, double aR,
                                         double ustar, double Pstar, double *rho, double *u, double *P,
                                         double dxdt) {
    double PdPR = Pstar / PR;
    double SR = uR + aR * sqrt(c->gp1d2g * PdPR + c->gm1d2g);
    if (SR > dxdt) {
        *rho = rhoR * (PdPR + c->gm1dgp1) / (c->gm1dgp1 * PdPR + 1.);
        *u = ustar;
        *P = Pstar;
    } else {
        *rho = rhoR;
        *u = uR;
        *P = PR;
    }
}

static inline void rs_sample_right_rarefaction(const rs_consts *c, double rhoR, double uR, double PR, double aR,
                                               double ustar, double Pstar, double *rho, double *u, double *P,
                                               double dxdt) {
    double SHR = uR + aR;
    if (SHR > dxdt) {
        double PdPR = Pstar / PR;
        double STR = ustar + aR * pow(PdPR, c->gm1d2g);
        if (STR > dxdt) {
            *rho = rhoR * pow(PdPR, c->ginv);
            *u = ustar;
            *P = Pstar;
        } else {
            double base = c->tdgp1 - c->gm1dgp1 * (uR - dxdt) / aR;
            *rho = rhoR * pow(base, c->tdgm1);
            *u = c->tdgp1 * (-aR + c->gm1d2 * uR + dxdt);
            *P = PR * pow(base, c->tgdgm1);
        }
    } else {
        *rho = rhoR;
        *u = uR;
        *P = PR;
    }
}

static inline void rs_sample_left_shock(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                                        double ustar, double Pstar, double *rho, double *u, double *P,
                                        double dxdt) {
    double PdPL = Pstar / PL;
    double SL = uL - aL * sqrt(c->gp1d2g * PdPL + c->gm1d2g);
    if (SL < dxdt) {
        *rho = rhoL * (PdPL + c->gm1dgp1) / (c->gm1dgp1 * PdPL + 1.);
        *u = ustar;
        *P = Pstar;
    } else {
        *rho = rhoL;
        *u = uL;
        *P = PL;
    }
}

static inline void rs_sample_left_rarefaction(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                                              double ustar, double Pstar, double *rho, double *u, double *P,
                                              double dxdt) {
    double SHL = uL - aL;
    if (SHL < dxdt) {
        double PdPL = Pstar / PL;
        double STL = ustar - aL * pow(PdPL, c->gm1d2g);
        if (STL > dxdt) {
            double base = c->tdgp1 + c->gm1dgp1 * (uL - dxdt) / aL;
            *rho = rhoL * pow(base, c->tdgm1);
            *u = c->tdgp1 * (aL + c->gm1d2 * uL + dxdt);
            *P = PL * pow(base, c->tgdgm1);
        } else {
            *rho = rhoL * pow(PdPL, c->ginv);
            *u = ustar;
            *P = Pstar;
        }
    } else {
        *rho = rhoL;
        *u = uL;
        *P = PL;
    }
}

/* ---- vacuum (Toro section 4.6) ---- */
static inline int rs_sample_right_vacuum(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                                         double *rho, double *u, double *P, double dxdt) {
    if (uL - aL < dxdt) {
        double SL = uL + c->tdgm1 * aL; /* vacuum front */
        if (SL > dxdt) {
            double base = c->tdgp1 + c->gm1dgp1 * (uL - dxdt) / aL;
            *rho = rhoL * pow(base, c->tdgm1);
            *u = c->tdgp1 * (aL + c->gm1d2 * uL + dxdt);
            *P = PL * pow(base, c->tgdgm1);
            return -1;
        } else {
            *rho = 0.; *u = 0.; *P = 0.;
            return 0;
        }
    } else {
        *rho = rhoL; *u = uL; *P = PL;
        return -1;
    }
}

static inline int rs_sample_left_vacuum(const rs_consts *c, double rhoR, double uR, double PR, double aR,
                                        double *rho, double *u, double *P, double dxdt) {
    if (dxdt < uR + aR) {
        double SR = uR - c->tdgm1 * aR; /* vacuum front */
        if (SR < dxdt) {
            double base = c->tdgp1 - c->gm1dgp1 * (uR - dxdt) / aR;
            *rho = rhoR * pow(base, c->tdgm1);
            *u = c->tdgp1 * (-aR + c->gm1d2 * uR + dxdt);
            *P = PR * pow(base, c->tgdgm1);
            return 1;
        } else {
            *rho = 0.; *u = 0.; *P = 0.;
            return 0;
        }
    } else {
        *rho = rhoR; *u = uR; *P = PR;
        return 1;
    }
}

static inline int rs_sample_vacuum_generation(const rs_consts *c, double rhoL, double uL, double PL, double aL,
                                              double rhoR, double uR, double PR, double aR,
                                              double *rho, double *u, double *P, double dxdt) {
    double SR = uR - c->tdgm1 * aR;
    double SL = uL + c->tdgm1 * aL;
    if (SR > dxdt && SL < dxdt) {
        *rho = 0.; *u = 0.; *P = 0.;
        return 0;
    } else {
        if (SL < dxdt) {
            return rs_sample_left_vacuum(c, rhoR, uR, PR, aR, rho, u, P, dxdt);
        } else {
            return rs_sample_right_vacuum(c, rhoL, uL, PL, aL, rho, u, P, dxdt);
        }
    }
}

static inline int rs_solve_vacuum(const rs_consts *c, double rhoL, double uL, double PL,
                                  double rhoR, double uR, double PR,
                                  double *rho, double *u, double *P, double dxdt) {
    double aL, aR;
    if (rhoL == 0. && rhoR == 0.) {
        *rho = 0.; *u = 0.; *P = 0.;
        return 0;
    }
    if (rhoR == 0.) {
        aL = rs_soundspeed(c, rhoL, PL);
        return rs_sample_right_vacuum(c, rhoL, uL, PL, aL, rho, u, P, dxdt);
    }
    if (rhoL == 0.) {
        aR = rs_soundspeed(c, rhoR, PR);
        return rs_sample_left_vacuum(c, rhoR, uR, PR, aR, rho, u, P, dxdt);
    }
    aL = rs_soundspeed(c, rhoL, PL);
    aR = rs_soundspeed(c, rhoR, PR);
    return rs_sample_vacuum_generation(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, rho, u, P, dxdt);
}

/*
 * Solve and sample at x/t = dxdt.  Returns +1 if the sampled point lies right of
 * the contact, -1 if left, 0 if it lies in vacuum (Riemann.cpp:104-127).
 * `iters` (may be NULL) receives the Newton iteration count (test hook for
 * iteration-for-iteration comparison with the CUDA solver).
 */
static inline int rs_solve(const rs_consts *c, double rhoL, double uL, double PL,
                           double rhoR, double uR, double PR,
                           double *rhosol, double *usol, double *Psol, double dxdt, int *iters) {
    double aL, aR, Pstar, Pguess, fPstar, fPguess, ustar;
    int it = 0;
    if (iters) *iters = 0;
    if (rhoL == 0. || rhoR == 0.) {
        return rs_solve_vacuum(c, rhoL, uL, PL, rhoR, uR, PR, rhosol, usol, Psol, dxdt);
    }
    aL = rs_soundspeed(c, rhoL, PL);
    aR = rs_soundspeed(c, rhoR, PR);
    if (c->tdgm1 * (aL + aR) <= uR - uL) {
        return rs_solve_vacuum(c, rhoL, uL, PL, rhoR, uR, PR, rhosol, usol, Psol, dxdt);
    }
    Pstar = 0.;
    Pguess = rs_guess_P(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR);
    fPstar = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pstar);
    fPguess = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pguess);
    if (fPstar * fPguess >= 0.) {
        /* Newton-Raphson until convergence or until a bracket for Brent appears */
        while (fabs(Pstar - Pguess) > 5.e-9 * (Pstar + Pguess) && fPguess < 0.) {
            Pstar = Pguess;
            fPstar = fPguess;
            Pguess = Pguess - fPguess / rs_fprime(c, rhoL, PL, aL, rhoR, PR, aR, Pguess);
            fPguess = rs_f(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pguess);
            ++it;
        }
    }
    RS_STAT_ADD(rs_stat_newton, it);
    if (1.e6 * fabs(Pstar - Pguess) > 0.5 * (Pstar + Pguess) && fPguess > 0.) {
        Pstar = rs_brent(c, rhoL, uL, PL, aL, rhoR, uR, PR, aR, Pstar, Pguess, fPstar, fPguess);
    } else {
        Pstar = Pguess;
    }
    if (iters) *iters = it;
    ustar = 0.5 * (uL + uR) + 0.5 * (rs_fb(c, rhoR, PR, aR, Pstar) - rs_fb(c, rhoL, PL, aL, Pstar));
    if (ustar < dxdt) {
        if (Pstar > PR) {
            rs_sample_right_shock(c, rhoR, uR, PR, aR, ustar, Pstar, rhosol, usol, Psol, dxdt);
        } else {
            rs_sample_right_rarefaction(c, rhoR, uR, PR, aR, ustar, Pstar, rhosol, usol, Psol, dxdt);
        }
        return 1;
    } else {
        if (Pstar > PL) {
            rs_sample_left_shock(c, rhoL, uL, PL, aL, ustar, Pstar, rhosol, usol, Psol, dxdt);
        } else {
            rs_sample_left_rarefaction(c, rhoL, uL, PL, aL, ustar, Pstar, rhosol, usol, Psol, dxdt);
        }
        return -1;
    }
}

#endif /* MLH_ORACLE_RIEMANN_EXACT_H */
