/*
 * oracle/ref_build/shim/RiemannSolver.hpp -- TEST INFRASTRUCTURE.
 *
 * Stand-in for the third-party header the reference includes at
 * /root/reference/demonstrator/include/Riemann.h:9 and downloads (unpinned, `master`)
 * at /root/reference/demonstrator/Makefile:93-94.  Same class name, constructor and
 * `solve` signature as used at /root/reference/demonstrator/src/Riemann.cpp:84,93-94;
 * the arithmetic is the restatement in oracle/riemann_exact.h (PARITY UNPINNED, see there).
 */
#ifndef RIEMANNSOLVER_HPP
#define RIEMANNSOLVER_HPP

#include "../../riemann_exact.h"

class RiemannSolver {
public:
    RiemannSolver(double gamma) { rs_init(&_c, gamma); }

    inline int solve(double rhoL, double uL, double PL, double rhoR, double uR, double PR,
                     double &rhosol, double &usol, double &Psol, double dxdt = 0.) const {
        return rs_solve(&_c, rhoL, uL, PL, rhoR, uR, PR, &rhosol, &usol, &Psol, dxdt, (int *)0);
    }

private:
    rs_consts _c;
};

#endif
