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
// Every pair (i,j) of the neighbour lists is ONE face.  The reference (ENFORCE_FLUX_SYM, quirk Q4) solves a face
// once, from the endpoint with the LOWER ORIGINAL index, and hands the other endpoint the exact negation
// (Particles.cpp:1838-1858, ghosts :1886-1907).  Same here: K2 marks the owner of every list slot, k_face_index
// numbers the owned slots (exclusive scan of the per-particle counts) and resolves, for every non-owned slot, the
// face index its partner assigned -- so each face is evaluated once and both endpoints GATHER +-F in their own list
// order.  No atomics, bit-reproducible, and total mass / momentum / energy are conserved to round-off (both ends
// add the same bits with opposite sign).  Faces whose partner has no list on this rank (halo particle of another
// slab) or does not list the pair (one-sided periodic pair, quirk Q9) are owned by the listing side; two ranks then
// evaluate a cut face redundantly with identical operands, so conservation holds across slabs without a flux exchange.
//
// Kernels:
//   k_face_index      thread per particle, MLH_FI_TRIP list slots per trip, OWNERS only: face list fa/fe (owner +
//                     canonical bit, list entry; staged in shared memory, written as full lines) and the slot -> face
//                     map (fmap); the owner scatters the face index into the partner's slot, which it finds from the
//                     group start (grp) + the offset r K2 recorded (no search).
//   k_face_states (K4a)  thread per face (persistent grid, two waves): A_ij, boosted + reconstructed + limited +
//                        predicted states of both endpoints, rotated into the face frame -> face record (4D+4 doubles)
//                        in a field-major staging buffer.  Gather/latency-bound; consecutive faces share their owner
//                        (warp broadcast), the face list is read one trip ahead, the records move with 256-bit loads
//                        (3D also prefetches the next trip's records into L1).
//   k_face_setup / k_face_iterate / k_face_finish (K4b)  the Riemann class of the reference: start of the exact
//                        solver on the six normal-direction fields (faces that need iterations go to a queue; faces
//                        whose guess already is the root are finished) / persistent lanes iterate queued faces to
//                        convergence (Newton-Raphson, Brent with the early exit of DESIGN.md 3.2), state in registers,
//                        warp-private cp.async rings / star state at x/t = 0, rotation back, projection -> F (D+2
//                        doubles, canonical orientation).  setup and finish stream at 55-80 % of the HBM peak, the
//                        iteration is a dependent FP64 chain per lane.
//                        [The first version fused everything into one 255-register, 145 KB kernel: ncu showed 56 %
//                        of the warp stalls were instruction fetches and 14 % FP64-pipe use -- profiles/r01a_*.]
//   k_flux_sum_update (K4c/K5) thread per particle, four slots per trip: signed sum of the faces of its slots in list
//                        order, conserved update, drift, dt published, bounding box of the new positions reduced.
// The per-slot buffers of the reference (psijTilde, Aij, WijL/R, Fij, vFrame for ALL particles, ~100 kB per
// particle, Particles.h:201-229) are replaced by the staging buffer and 32/48 B of flux per face.
#include "mlh_internal.cuh"
#include <cfloat>
#include <cstdlib>
#include <cuda_pipeline.h>

namespace {

// ---------------------------------------------------------------------------------------------
// exact Riemann solver (device restatement of oracle/riemann_exact.h: Newton-Raphson with Toro's
// adaptive guess, Brent fallback, sampling at x/t=0 -- same control flow, same stopping rules).
// Two deliberate arithmetic differences, both ~1e-15 relative (the bar is 1e-10):
//   * pow(x,y) = exp(y log x) (one noinline copy) instead of a correctly rounded pow;
//   * f_K and f_K' are evaluated together at each Newton point and cached, and the identities
//     x^-(g+1)/2g = x^(g-1)/2g / x,  x^(1/g) = x / (x^(g-1)/2g)^2,  b^(2g/(g-1)) = b^(2/(g-1)) b^2
//     replace the extra pow calls of the textbook formulas.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rs_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double rs_min(double a, double b) { return (b < a) ? b : a; }

__device__ __noinline__ double mlh_pow(double x, double y) { return x == 0. ? 0. : exp(y * log(x)); }

// 1/x for normal positive x: MUFU.RCP64H estimate + two Newton steps
__device__ __forceinline__ double rs_rcp(double x) {
    if (!(x > 1e-300 && x < 1e300)) return 1. / x;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.);
    y = fma(y, e, y);
    e = fma(-x, y, 1.);
    return fma(y, e, y);
}

// x^((gamma-1)/(2 gamma)), the one power the root finder evaluates every iteration.  For gamma = 5/3 and 7/5 the
// exponent is 1/5 resp. 1/7: z ~ x^(-1/n) from a single-precision seed (relative error e < 1e-6), residual
// r = 1 - x z^n, and ONE third-order correction z <- z (1 + r/n + (n+1)/(2 n^2) r^2) -- the first terms of
// (1 - r)^(-1/n); the remainder 0.09 r^3 (r ~ n e) is < 1e-17.  Then x^(1/n) = x z^(n-1).
// ~10 FP64 instructions on a short dependency chain (two Newton steps took 15) instead of ~130 for exp(y log x);
// x = 1 gives exactly 1 (as pow does), which keeps f(P) = 0 exact on faces between identical states.  Other gamma,
// and x outside [1e-12, 1e12] (where the float seed is no longer good enough), take the generic path.
// ONE_STEP = false (setup / finish kernels): the earlier form with two division-free Newton steps
// z <- z + z (1 - x z^n)/n, same accuracy; kept there because the shorter form changes their register allocation
// (setup 54 -> 84) and with it the occupancy their grids are tuned for (A/B r01x: finish 0.125 -> 0.188 ms).
template <bool ONE_STEP>
__device__ __forceinline__ double rs_root_pow(const RsConsts &c, double x) {
    const int n = c.root_n;
    const double rn = c.gm1d2g; // 1/n
    if (ONE_STEP) {
        if (n == 0 || !(x > 1e-12 && x < 1e12)) return mlh_pow(x, c.gm1d2g);
        const double c2 = 0.5 * rn * (rn + 1.);
        double z = (double)exp2f(-__log2f((float)x) * (float)rn);
        const double z2 = z * z, z4 = z2 * z2;
        const double zn = (n == 5) ? z4 * z : z4 * z2 * z;
        const double r = fma(-x, zn, 1.);
        z = fma(z, r * fma(c2, r, rn), z);
        const double y2 = z * z, y4 = y2 * y2;
        return x * ((n == 5) ? y4 : y4 * y2);
    }
    if (n == 0 || !(x > 1e-30 && x < 1e30)) return mlh_pow(x, c.gm1d2g);
    double z = (double)exp2f(-__log2f((float)x) * (float)rn);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double z2 = z * z, z4 = z2 * z2;
        const double zn = (n == 5) ? z4 * z : z4 * z2 * z;
        z = fma(z, fma(-x, zn, 1.) * rn, z);
    }
    const double z2 = z * z, z4 = z2 * z2;
    return x * ((n == 5) ? z4 : z4 * z2);
}

// f_K(Ps) and (optionally) f_K'(Ps) of BOTH sides, Toro eqs. 4.6/4.7/4.37.  In smooth flow P* lies between
// PL and PR, i.e. every face has one (weak) shock side and one rarefaction side -- but WHICH side differs from
// lane to lane.  Evaluating "left, then right" would run the pow and the sqrt twice with half-empty warps;
// instead each lane first serves its first rarefaction side and its first shock side, whichever they are,
// and only lanes with two sides of the same kind take a second round.
struct RsEval {
    double fL, fR, fpL, fpR, wL, wR; // w_K = (Ps/P_K)^((g-1)/2g), rarefaction sides only
};
// RECIP: Ps/P_K is formed as Ps * (1/P_K) with the reciprocals passed in -- k_face_iterate only, where the SAME face may
// take this path or rs_f_nobranch depending on what the other lanes of its warp are doing: both must round alike, or
// the result would depend on the (atomic) queue order.
template <bool NEED_FP, bool RECIP = false>
__device__ __forceinline__ void rs_eval2(const RsConsts &c, double rhoL, double PL, double aL, double rhoR, double PR,
                                         double aR, double Ps, RsEval &e, double iPL = 0., double iPR = 0.) {
    const bool shL = Ps > PL, shR = Ps > PR;
    double wL = 0., wR = 0., rL = 0., rR = 0., qL = 0., qR = 0.;
    if (!shL || !shR) {
        const bool firstL = !shL;
        const double r1 = RECIP ? Ps * (firstL ? iPL : iPR) : Ps / (firstL ? PL : PR);
        const double w1 = rs_root_pow<RECIP>(c, r1);
        if (firstL) { wL = w1; rL = r1; } else { wR = w1; rR = r1; }
        if (!shL && !shR) {
            rR = RECIP ? Ps * iPR : Ps / PR;
            wR = rs_root_pow<RECIP>(c, rR);
        }
    }
    if (shL || shR) {
        const bool firstL = shL;
        // sqrt(A_K / (Ps + B_K)) = sqrt(2/(g+1)) / sqrt(rho_K (Ps + B_K)): one rsqrt instead of two divisions and a sqrt
        const double q1 = c.sqrt_tdgp1 * rsqrt((firstL ? rhoL : rhoR) * (Ps + c.gm1dgp1 * (firstL ? PL : PR)));
        if (firstL) qL = q1; else qR = q1;
        if (shL && shR) qR = c.sqrt_tdgp1 * rsqrt(rhoR * (Ps + c.gm1dgp1 * PR));
    }
    e.wL = wL;
    e.wR = wR;
    e.fL = shL ? (Ps - PL) * qL : c.tdgm1 * aL * (wL - 1.);
    e.fR = shR ? (Ps - PR) * qR : c.tdgm1 * aR * (wR - 1.);
    if (NEED_FP) {
        // shock: (1 - (Ps-P)/(2(B+Ps))) q      rarefaction: (Ps/P)^-(g+1)/2g / (rho a) = w / (r rho a)
        const double tL = (shL ? 0.5 * (Ps - PL) : wL) / (shL ? (c.gm1dgp1 * PL + Ps) : (rL * rhoL * aL));
        const double tR = (shR ? 0.5 * (Ps - PR) : wR) / (shR ? (c.gm1dgp1 * PR + Ps) : (rR * rhoR * aR));
        e.fpL = shL ? (1. - tL) * qL : tL;
        e.fpR = shR ? (1. - tR) * qR : tR;
    }
}

__device__ __forceinline__ double rs_gb(const RsConsts &c, double rho, double P, double Pstar) {
    // sqrt(A / (Pstar + B)), A = 2/((g+1) rho), B = (g-1)/(g+1) P, as one rsqrt (the form rs_eval2 uses)
    return c.sqrt_tdgp1 * rsqrt(rho * (Pstar + c.gm1dgp1 * P));
}

// x^(2 gamma/(gamma-1)) = x^n for the gammas whose (gamma-1)/(2 gamma) is 1/n (n = 5: 5/3, n = 7: 7/5)
__device__ __forceinline__ double rs_int_pow(int n, double x) {
    const double x2 = x * x, x4 = x2 * x2;
    return n == 5 ? x4 * x : x4 * x2 * x;
}

// Toro's two-rarefaction / two-shock guesses (4.46, 4.48).  In smooth flow nearly EVERY face comes here (the PVRS value
// falls outside [Pmin, Pmax] as soon as |du| exceeds the pressure jump), half of them into each branch: with the generic
// pow = exp(y log x) this function was 45 % of k_face_setup's instructions on KH (3 pow calls of ~130 instructions at
// 14 of 32 lanes, profiles/r2i_setup_lines_kh1000j.txt).  For gamma = 5/3 and 7/5 the exponents are 1/n and n:
// two n-th roots (rs_root_pow) and an integer power instead.
__device__ __noinline__ double rs_guess_nonlinear(const RsConsts c, double rhoL, double uL, double PL, double aL, double rhoR,
                                                  double uR, double PR, double aR, double Ppv, double Pmin) {
    if (Ppv < Pmin) { // two rarefactions
        const double num = aL + aR - c.gm1d2 * (uR - uL);
        if (c.root_n != 0 && PL > 1e-30 && PL < 1e30 && PR > 1e-30 && PR < 1e30)
            return rs_int_pow(c.root_n, num / (aL / rs_root_pow<false>(c, PL) + aR / rs_root_pow<false>(c, PR)));
        return mlh_pow(num / (aL / mlh_pow(PL, c.gm1d2g) + aR / mlh_pow(PR, c.gm1d2g)), c.tgdgm1);
    }
    double gL = rs_gb(c, rhoL, PL, Ppv); // two shocks
    double gR = rs_gb(c, rhoR, PR, Ppv);
    return (gL * PL + gR * PR - uR + uL) / (gL + gR);
}

struct RsProblem {
    double rhoL, PL, aL, rhoR, PR, aR, du; // du = uR - uL
    double Pguess, fPguess, f0, fpsum;     // f(Pguess), f(0), fL'(Pguess) + fR'(Pguess)
    double iPL, iPR;                       // 1/PL, 1/PR (k_face_iterate only: one division per face instead of one per iteration)
};

// Root finder as a resumable state machine: one call of rs_iter_step() = one new trial pressure + one
// evaluation of f, for either method.  Control flow and iterates are those of oracle/riemann_exact.h
// (rs_solve's Newton loop, then rs_brent; the inverse quadratic step is written over one common
// denominator).  Brent is a MAIN path: Newton only runs when the initial guess lies below the root,
// which is the minority of faces, and Brent's iteration count has a long tail (3..20) -- hence the
// per-iteration regrouping in k_face_riemann.
//   Newton fields: a = Pstar, fa = f(Pstar), b = Pguess, fb = f(Pguess), c = fL'(Pguess) + fR'(Pguess)
//   Brent fields:  a, b, c, d, fa, fb, fc, mflag as in the textbook
enum { RS_DONE = 0, RS_NEWTON = 1, RS_BRENT = 2 };
struct RsIter {
    double a, b, c, d, fa, fb, fc;
    double fpb; // Brent: f' at the upper end of the bracket it started from = lower bound of f' on the bracket (f is concave)
    int method, mflag;
};

// Early exit of Brent's method, exact to 2e-13 relative.  f(P) = f_L + f_R + du is increasing and CONCAVE (Toro 4.3.1), so on
// a bracket [a, b] that started at an upper end U: f'(x) >= f'(U) = fpb for every x <= U.  If the best point b is the upper
// end and 0 < f(b) <= 1e-13 b fpb, the root lies in [b (1 - 1e-13), b], and everything the reference's iteration can still
// return -- a later b has |f| <= f(b), hence lies within f(b)/fpb of the root -- lies in [b (1 - 2e-13), b].  The reference
// (RiemannSolver::solve, restated in oracle/riemann_exact.h) spends 27-29 further bisections on exactly these faces: its
// interpolation steps collapse onto b and the lower end creeps up from 0 by halving until |a - b| < 5e-9 (a + b).
// CPU study on the oracle's faces (3 steps in): the rule fires on 56 % / 57 % / 30 % of the faces of KH / Sedov / fluid block
// and removes 75 % / 35 % / 65 % of ALL root-finder iterations; largest deviation of P* from the reference's value 1.0e-13.
#define MLH_BRENT_CERTAIN 1e-13
__device__ __forceinline__ bool rs_brent_certain(const RsIter &it) {
    return (it.b > it.a) & (it.fb > 0.) & (it.fb <= MLH_BRENT_CERTAIN * it.b * it.fpb);
}

// enter Brent on [lower, upper]; returns false if it terminates immediately (result in it.b)
__device__ __forceinline__ bool rs_brent_begin(RsIter &it, double lower, double upper, double lowf, double upf, double fp_upper) {
    double a = lower, b = upper, fa = lowf, fb = upf;
    it.method = RS_DONE;
    it.b = b;
    it.fpb = fp_upper;
    if (fa * fb > 0.) return false; // not bracketed: keep upper (as the oracle)
    if (fabs(fa) < fabs(fb)) {
        double t = a; a = b; b = t;
        t = fa; fa = fb; fb = t;
    }
    it.a = a; it.b = b; it.fa = fa; it.fb = fb;
    it.c = a; it.fc = fa; it.d = 1e230; it.mflag = 1;
    if (!(fb == 0.) && (fabs(a - b) > 5.e-9 * (a + b)) && !rs_brent_certain(it)) {
        it.method = RS_BRENT;
        return true;
    }
    return false;
}

// after rs_setup: which method runs first (the oracle's `if (fPstar*fPguess >= 0)` / Brent test)
__device__ __forceinline__ void rs_iter_begin(const RsProblem &q, RsIter &it) {
    const double Pstar = 0., fPstar = q.f0, Pguess = q.Pguess, fPguess = q.fPguess;
    it.method = RS_DONE;
    it.mflag = 0;
    it.b = Pguess;
    it.a = Pstar; it.fa = fPstar; it.fb = fPguess; it.c = q.fpsum; it.d = 0.; it.fc = 0.;
    it.fpb = q.fpsum;
    if (fPstar * fPguess >= 0.) {
        if (fabs(Pstar - Pguess) > 5.e-9 * (Pstar + Pguess) && fPguess < 0.) {
            it.method = RS_NEWTON;
            return;
        }
    }
    if (1.e6 * fabs(Pstar - Pguess) > 0.5 * (Pstar + Pguess) && fPguess > 0.) rs_brent_begin(it, Pstar, Pguess, fPstar, fPguess, q.fpsum);
}

// trial pressure of this iteration
__device__ __forceinline__ double rs_iter_trial(RsIter &it) {
    if (it.method == RS_NEWTON) {
        it.a = it.b; // Pstar = Pguess
        it.fa = it.fb;
        it.b = it.b - it.fb / it.c;
        return it.b;
    }
    const double a = it.a, b = it.b, c = it.c, d = it.d, fa = it.fa, fb = it.fb, fc = it.fc;
    const bool mflag = it.mflag != 0;
    double s;
    if ((fa != fc) && (fb != fc)) {
        const double dab = fa - fb, dac = fa - fc, dbc = fb - fc;
        s = (a * fb * fc * dbc - b * fa * fc * dac + c * fa * fb * dab) / (dab * dac * dbc);
    } else {
        s = b - fb * (b - a) / (fb - fa);
    }
    const double tmp2 = 0.25 * (3. * a + b);
    if (!(((s > tmp2) && (s < b)) || ((s < tmp2) && (s > b))) || (mflag && (fabs(s - b) >= (0.5 * fabs(b - c)))) ||
        (!mflag && (fabs(s - b) >= (0.5 * fabs(c - d)))) || (mflag && (fabs(b - c) < 5.e-9 * (b + c))) ||
        (!mflag && (fabs(c - d) < 5.e-9 * (c + d)))) {
        s = 0.5 * (a + b);
        it.mflag = 1;
    } else {
        it.mflag = 0;
    }
    return s;
}

// digest f(s) (and f'(s) for Newton); afterwards it.method == RS_DONE means P* = it.b
__device__ __forceinline__ void rs_iter_update(RsIter &it, double s, double fs, double fps) {
    if (it.method == RS_NEWTON) {
        it.fb = fs;
        it.c = fps;
        const double Pstar = it.a, Pguess = it.b;
        if (fabs(Pstar - Pguess) > 5.e-9 * (Pstar + Pguess) && fs < 0.) return; // next Newton iteration
        it.method = RS_DONE;
        if (1.e6 * fabs(Pstar - Pguess) > 0.5 * (Pstar + Pguess) && fs > 0.) rs_brent_begin(it, Pstar, Pguess, it.fa, fs, fps);
        return;
    }
    double a = it.a, b = it.b, fa = it.fa, fb = it.fb;
    it.d = it.c;
    it.c = b;
    it.fc = fb;
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
    it.a = a; it.b = b; it.fa = fa; it.fb = fb;
    if (!(!(fb == 0.) && (fabs(a - b) > 5.e-9 * (a + b))) || rs_brent_certain(it)) it.method = RS_DONE;
}

// f_L(Ps) + f_R(Ps) + du without branches: both sides evaluate the rarefaction AND the shock expression (same
// formulas as rs_eval2, so the values are identical) and select.  Inside a warp the sides of the 32 faces are a mix of
// both kinds anyway, so the divergent version executed all four paths with half-empty warps (profiles/r01k); here the
// two sides are independent instruction streams that overlap in the FP64 pipe.
__device__ __forceinline__ double rs_f_nobranch(const RsConsts &c, const RsProblem &q, double Ps) {
    const double wL = rs_root_pow<true>(c, Ps * q.iPL), wR = rs_root_pow<true>(c, Ps * q.iPR);
    const double qL = c.sqrt_tdgp1 * rsqrt(q.rhoL * (Ps + c.gm1dgp1 * q.PL));
    const double qR = c.sqrt_tdgp1 * rsqrt(q.rhoR * (Ps + c.gm1dgp1 * q.PR));
    const double fL = (Ps > q.PL) ? (Ps - q.PL) * qL : c.tdgm1 * q.aL * (wL - 1.);
    const double fR = (Ps > q.PR) ? (Ps - q.PR) * qR : c.tdgm1 * q.aR * (wR - 1.);
    return fL + fR + q.du;
}

// one Brent iteration (rs_iter_trial + f + rs_iter_update for method == RS_BRENT) with selects instead of branches;
// same expressions, same decisions
__device__ __forceinline__ void rs_brent_step(const RsConsts &cst, const RsProblem &q, RsIter &it) {
    const double a = it.a, b = it.b, c = it.c, d = it.d, fa = it.fa, fb = it.fb, fc = it.fc;
    const bool mflag = it.mflag != 0;
    const bool iqi = (fa != fc) & (fb != fc);
    const double dab = fa - fb, dac = fa - fc, dbc = fb - fc;
    const double numI = a * fb * fc * dbc - b * fa * fc * dac + c * fa * fb * dab;
    const double denI = dab * dac * dbc;
    const double numS = fb * (b - a), denS = fb - fa;
    const double qv = (iqi ? numI : numS) / (iqi ? denI : denS);
    double s = iqi ? qv : b - qv;
    const double tmp2 = 0.25 * (3. * a + b);
    const bool between = ((s > tmp2) & (s < b)) | ((s < tmp2) & (s > b));
    const double sb = fabs(s - b), bc = fabs(b - c), cd = fabs(c - d);
    const bool bis = (!between) | (mflag & (sb >= 0.5 * bc)) | ((!mflag) & (sb >= 0.5 * cd)) | (mflag & (bc < 5.e-9 * (b + c))) |
                     ((!mflag) & (cd < 5.e-9 * (c + d)));
    s = bis ? 0.5 * (a + b) : s;
    it.mflag = bis ? 1 : 0;
    const double fs = rs_f_nobranch(cst, q, s);
    it.d = c;
    it.c = b;
    it.fc = fb;
    const bool left = fa * fs < 0.;
    const double na = left ? a : s, nfa = left ? fa : fs, nb = left ? s : b, nfb = left ? fs : fb;
    const bool sw = fabs(nfa) < fabs(nfb);
    it.a = sw ? nb : na;
    it.fa = sw ? nfb : nfa;
    it.b = sw ? na : nb;
    it.fb = sw ? nfa : nfb;
    if (!(!(it.fb == 0.) && (fabs(it.a - it.b) > 5.e-9 * (it.a + it.b))) || rs_brent_certain(it)) it.method = RS_DONE;
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
                *rho = rhoL * mlh_pow(base, c.tdgm1);
                *u = c.tdgp1 * (aL + c.gm1d2 * uL + dxdt);
                *P = PL * mlh_pow(base, c.tgdgm1);
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
            *rho = rhoR * mlh_pow(base, c.tdgm1);
            *u = c.tdgp1 * (-aR + c.gm1d2 * uR + dxdt);
            *P = PR * mlh_pow(base, c.tgdgm1);
            return 1;
        }
        *rho = 0.; *u = 0.; *P = 0.;
        return 0;
    }
    *rho = rhoR; *u = uR; *P = PR;
    return 1;
}

// The solver is split in three stages so that a thread block can regroup its faces between them
// (k_face_riemann): setup (sound speeds, vacuum test, initial guess, f at 0 and at the guess), root (Newton /
// Brent for P*), sample (star state at x/t = 0).  Together they are RiemannSolver::solve (Riemann.cpp:93-94).

// (k_face_setup / k_face_finish pin their register budgets with __launch_bounds__: ptxas' natural allocation moved
// between 54 and 84 registers with unrelated edits, and the persistent grids are sized for a given occupancy)
// returns false if the (generated) vacuum path must be taken
__device__ __forceinline__ bool rs_setup(const RsConsts &c, double rhoL, double uL, double PL, double rhoR, double uR, double PR,
                                         RsProblem &q) {
    if (rhoL == 0. || rhoR == 0.) return false;
    q.rhoL = rhoL; q.PL = PL; q.rhoR = rhoR; q.PR = PR;
    q.aL = sqrt(c.gamma * PL / rhoL);
    q.aR = sqrt(c.gamma * PR / rhoR);
    q.du = uR - uL;
    if (c.tdgm1 * (q.aL + q.aR) <= q.du) return false;
    // initial guess (Toro 4.3.2), floored at 5e-9 (PL+PR)
    const double Pmin = rs_min(PL, PR), Pmax = rs_max(PL, PR);
    const double qmax = Pmax / Pmin;
    double Ppv = 0.5 * (PL + PR) - 0.125 * q.du * (PL + PR) * (q.aL + q.aR);
    Ppv = rs_max(5.e-9 * (PL + PR), Ppv);
    double Pguess;
    if (qmax <= 2. && Pmin <= Ppv && Ppv <= Pmax)
        Pguess = Ppv;
    else
        Pguess = rs_guess_nonlinear(c, rhoL, uL, PL, q.aL, rhoR, uR, PR, q.aR, Ppv, Pmin);
    q.Pguess = rs_max(5.e-9 * (PL + PR), Pguess);
    // f(0): both sides are rarefactions with (0/P)^.. = 0 exactly  ->  f_K(0) = -2 a_K/(g-1)
    if (PL > 0. && PR > 0.) {
        q.f0 = c.tdgm1 * q.aL * (0. - 1.) + c.tdgm1 * q.aR * (0. - 1.) + q.du;
    } else {
        RsEval e0;
        rs_eval2<false>(c, rhoL, PL, q.aL, rhoR, PR, q.aR, 0., e0);
        q.f0 = e0.fL + e0.fR + q.du;
    }
    RsEval e;
    rs_eval2<true>(c, rhoL, PL, q.aL, rhoR, PR, q.aR, q.Pguess, e);
    q.fPguess = e.fL + e.fR + q.du;
    q.fpsum = e.fpL + e.fpR;
    return true;
}

// star state at x/t = 0; returns +1 (right of the contact sampled) or -1 (left); Riemann.cpp:104-127 reads the flag
__device__ __forceinline__ int rs_sample(const RsConsts &c, const RsProblem &q, double uL, double uR, double Pstar,
                                         double *rhosol, double *usol, double *Psol) {
    RsEval e;
    rs_eval2<false>(c, q.rhoL, q.PL, q.aL, q.rhoR, q.PR, q.aR, Pstar, e);
    const double ustar = 0.5 * (uL + uR) + 0.5 * (e.fR - e.fL);
    // one code path serves both sides: right family waves move with u + a.., left with u - a..
    const bool right = ustar < 0.;
    const double sg = right ? 1. : -1.;
    const double uS = right ? uR : uL, aS = right ? q.aR : q.aL, PS = right ? q.PR : q.PL, rhoS = right ? q.rhoR : q.rhoL;
    const double wS = right ? e.wR : e.wL;
    double rho_o = rhoS, u_o = uS, P_o = PS;
    if (Pstar > PS) { // shock
        const double PdP = Pstar / PS;
        const double Sspeed = uS + sg * aS * sqrt(c.gp1d2g * PdP + c.gm1d2g);
        if (right ? (Sspeed > 0.) : (Sspeed < 0.)) {
            rho_o = rhoS * (PdP + c.gm1dgp1) / (c.gm1dgp1 * PdP + 1.);
            u_o = ustar;
            P_o = Pstar;
        }
    } else { // rarefaction
        const double head = uS + sg * aS;
        if (right ? (head > 0.) : (head < 0.)) {
            const double PdP = Pstar / PS;
            const double tail = ustar + sg * aS * wS; // wS = PdP^((g-1)/2g)
            // right: star state if tail > 0 else fan; left: fan if tail > 0 else star state
            if (tail > 0. ? right : !right) {
                rho_o = rhoS * (PdP / (wS * wS)); // PdP^(1/g)
                u_o = ustar;
                P_o = Pstar;
            } else {
                const double base = c.tdgp1 - sg * c.gm1dgp1 * uS / aS;
                const double bp = mlh_pow(base, c.tdgm1);
                rho_o = rhoS * bp;
                u_o = c.tdgp1 * (-sg * aS + c.gm1d2 * uS);
                P_o = PS * (bp * base * base); // base^(2g/(g-1))
            }
        }
    }
    *rhosol = rho_o;
    *usol = u_o;
    *Psol = P_o;
    return right ? 1 : -1;
}

// ---------------------------------------------------------------------------------------------
// Particles::pairwiseLimiter, Particles.cpp:1735-1785 (quirk Q1)
// ---------------------------------------------------------------------------------------------
// Both endpoints of a face call the limiter with (phi_i, phi_j) swapped: phiMin/phiMax, delta1/delta2 and phiPlus/phiMinus
// are the same values for the two calls (|phi_i - phi_j| is symmetric in either abs mode), so they are computed once
// per component (PairLimits) and each side only does its own selection -- same expressions, same results.
struct PairLimits {
    double delta2, phiPlus, phiMinus;
};
__device__ __forceinline__ PairLimits pairwise_limits(const Params &p, double phi_i, double phi_j) {
    const int am = p.abs_mode;
    const bool lt = phi_i < phi_j;
    const double phiMin = lt ? phi_i : phi_j, phiMax = lt ? phi_j : phi_i;
    const double ad = q1_abs(phi_i - phi_j, am);
    const double delta1 = p.psi1 * ad;
    PairLimits r;
    r.delta2 = p.psi2 * ad;
    if ((phiMax + delta1 >= 0. && phiMax >= 0.) || (phiMax + delta1 < 0. && phiMax < 0.)) {
        r.phiPlus = phiMax + delta1;
    } else {
        r.phiPlus = phiMax / (1. + delta1 / q1_abs(phiMax, am));
    }
    if ((phiMin - delta1 >= 0. && phiMin >= 0.) || (phiMin - delta1 < 0. && phiMin < 0.)) {
        r.phiMinus = phiMin - delta1;
    } else {
        r.phiMinus = phiMin / (1. + delta1 / q1_abs(phiMin, am));
    }
    return r;
}
// ratio = |x_ij - x_i| / |x_j - x_i| (evaluated once per side: `xijxi_abs / xjxi_abs * (phi_j - phi_i)` is left-associative)
__device__ __forceinline__ double pairwise_limiter(const PairLimits &l, double phi0, double phi_i, double phi_j, double ratio) {
    // the two branches of the reference (phi_i < phi_j / phi_i > phi_j) as selects: in a warp both occur, lane by lane
    const bool lt = phi_i < phi_j, gt = phi_i > phi_j;
    const double phi_ij = phi_i + ratio * (phi_j - phi_i);
    // (one comparison direction per lane, predicate logic instead of nested selects: the nested form compiled to
    // divergent branches, 12 % of K4a's instructions at 14 active lanes -- profiles/r01u)
    const double t = phi_ij + (lt ? l.delta2 : -l.delta2);
    const bool take_t = lt ? (t < phi0) : (t > phi0);
    const double m = take_t ? t : phi0;                    // minPhiD2 / maxPhiD2
    const double lim = lt ? l.phiMinus : l.phiPlus;
    const bool take_lim = lt ? (lim > m) : (lim < m);
    const double r = take_lim ? lim : m;
    return (lt || gt) ? r : phi_i;
}

template <int D>
__device__ __forceinline__ double dotD(const double *a, const double *b) { // Helper::dotProduct
    double res = 0.;
#pragma unroll
    for (int k = 0; k < D; ++k) res += a[k] * b[k];
    return res;
}

// Riemann::Riemann (Riemann.cpp:7-81): unit normal of the face and the rotation Lambda taking it to the
// x axis (Helper::rotationMatrix2D/3D, Helper.cpp:39-77); the L and R velocities are rotated in place.
// Wa = state of the canonical particle ("WijR" of the caller = class member WL, the LEFT state of the solver),
// Wb = the neighbour's ("WijL" = class WR, RIGHT state): quirk Q5.  W = [rho, P, vx, vy(, vz)].
template <int D> struct FaceFrame {
    double L[D == 2 ? 2 : 9]; // 2D: (ax, ay) of [[ax, ay], [-ay, ax]]; 3D: row-major Rodrigues matrix
};
template <int D>
__device__ __forceinline__ void face_frame(const double *A, FaceFrame<D> &fr) {
    double n2 = 0.;
#pragma unroll
    for (int k = 0; k < D; ++k) n2 += A[k] * A[k];
#ifdef MLH_FRAME_RSQRT
    const double inv = rsqrt(n2);
#else
    const double inv = 1. / sqrt(n2);
#endif
    if (D == 2) {
        fr.L[0] = inv * A[0];
        fr.L[1] = inv * A[1];
    } else {
        // rotationMatrix3D(a = hatA, b = unitX): v = a x b = (0, a2, -a1), Rodrigues with n = 1/(1+cos)
        // (singular for hatA = -x, as the reference); rotationMatrix3D(unitX, hatA) is its transpose
        const double a0 = inv * A[0], a1 = inv * A[1], a2 = inv * A[D - 1];
        const double v1 = a2, v2 = -a1;
        const double n = 1. / (1. + a0);
        double *L = fr.L;
        L[0] = 1. - n * (v2 * v2 + v1 * v1);
        L[1] = -v2;
        L[2] = v1;
        L[3] = v2;
        L[4] = 1. - n * (v2 * v2);
        L[5] = n * v1 * v2;
        L[6] = -v1;
        L[7] = n * v1 * v2;
        L[8] = 1. - n * (v1 * v1);
    }
}
template <int D>
__device__ __forceinline__ void face_rotate(const double *A, double *Wa, double *Wb, FaceFrame<D> &fr) {
    face_frame<D>(A, fr);
    if (D == 2) {
        const double L0 = fr.L[0], L1 = fr.L[1];
        const double bR0 = Wb[2], bR1 = Wb[3], bL0 = Wa[2], bL1 = Wa[3];
        Wb[2] = L0 * bR0 + L1 * bR1;
        Wb[3] = -L1 * bR0 + L0 * bR1;
        Wa[2] = L0 * bL0 + L1 * bL1;
        Wa[3] = -L1 * bL0 + L0 * bL1;
    } else {
        const double *L = fr.L;
        const double bR[3] = {Wb[2], Wb[3], Wb[D + 1]}, bL[3] = {Wa[2], Wa[3], Wa[D + 1]};
        Wb[2] = L[0] * bR[0] + L[1] * bR[1] + L[2] * bR[2];
        Wb[3] = L[3] * bR[0] + L[4] * bR[1] + L[5] * bR[2];
        Wb[D + 1] = L[6] * bR[0] + L[7] * bR[1] + L[8] * bR[2];
        Wa[2] = L[0] * bL[0] + L[1] * bL[1] + L[2] * bL[2];
        Wa[3] = L[3] * bL[0] + L[4] * bL[1] + L[5] * bL[2];
        Wa[D + 1] = L[6] * bL[0] + L[7] * bL[1] + L[8] * bL[2];
    }
}

// Riemann::exact after the solve (Riemann.cpp:104-141) + rotateAndProjectFluxes{2D,3D} (:144-229): transverse
// velocity of the sampled side, rotation back, fluxes projected on the un-normalised A.  F = [m, E, px, py(, pz)].
template <int D>
__device__ __forceinline__ void face_project(const Params &p, int flag, double rhoSol, double uSol, double PSol, const double *Wa,
                                             const double *Wb, const FaceFrame<D> &fr, const double *vFrame, const double *A,
                                             double *F) {
    const double gamma = p.gamma;
    double vSol[D];
#pragma unroll
    for (int k = 0; k < D; ++k) vSol[k] = 0.;
    vSol[0] = uSol;
    if (flag == 1) {
#pragma unroll
        for (int k = 1; k < D; ++k) vSol[k] = Wb[2 + k];
    } else if (flag == -1) {
#pragma unroll
        for (int k = 1; k < D; ++k) vSol[k] = Wa[2 + k];
    }
    if (D == 2) {
        const double L0 = fr.L[0], L1 = fr.L[1];
        const double s0 = vSol[0], s1 = vSol[1];
        vSol[0] = L0 * s0 - L1 * s1; // rotationMatrix2D(unitX, hatA) = [[ax, -ay], [ay, ax]]
        vSol[1] = L1 * s0 + L0 * s1;
    } else {
        const double *L = fr.L;
        const double s[3] = {vSol[0], vSol[1], vSol[D - 1]};
        vSol[0] = L[0] * s[0] + L[3] * s[1] + L[6] * s[2];
        vSol[1] = L[1] * s[0] + L[4] * s[1] + L[7] * s[2];
        vSol[D - 1] = L[2] * s[0] + L[5] * s[1] + L[8] * s[2];
    }
    double vLab[D], Av = 0., v2 = 0.;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        Av += A[k] * rhoSol * vSol[k];
        vLab[k] = vSol[k] + vFrame[k];
        v2 += vLab[k] * vLab[k];
    }
    F[0] = Av;
    if (p.mfm) {
#pragma unroll
        for (int k = 0; k < D; ++k) vSol[k] = 0.;
    }
    const double ekin = PSol / (gamma - 1.) + rhoSol * .5 * v2;
    double FE = 0.;
#pragma unroll
    for (int al = 0; al < D; ++al) {
        double Fp = A[al] * PSol;
#pragma unroll
        for (int be = 0; be < D; ++be) Fp += A[be] * rhoSol * vLab[al] * vSol[be];
        F[2 + al] = Fp;
        FE += A[al] * (vSol[al] * ekin + PSol * vLab[al]);
    }
    F[1] = FE;
}

// device-side dt policy (MeshlessScheme.cpp:91-105): fixed dt, or CFL dt (min-reduced into dt_bits by K3b, all-reduced
// over the ranks) clipped to dt_max.  Evaluated by every thread that needs dt (no extra one-thread kernel); the update
// kernel publishes it in dt_used for the host.
__device__ __forceinline__ double select_dt(const Params &p, double dt_fixed, double dt_max) {
    if (dt_fixed >= 0.) return dt_fixed; // 0 is a legal step: the reference driver takes a zero-length step at every dump time (quirk Q7)
    double dt = __longlong_as_double((long long)*p.d.dt_bits);
    if (dt_max > 0. && dt > dt_max) dt = dt_max;
    return dt;
}

// W component nu -> gradient field slot: W = [rho, P, vx, vy, vz], slots rho 0, vx 1, vy 2, vz 3, P 4
__device__ __forceinline__ int w2f(int nu) { return nu == 0 ? 0 : (nu == 1 ? 4 : nu - 1); }

// staging record of one face, written by K4a with the velocities ALREADY ROTATED into the face frame (Riemann::Riemann,
// Riemann.cpp:19-81): the six fields of the one-dimensional Riemann problem come first, so the solver setup reads only
// those; the finish kernel reads everything.  4D+4 doubles, field-major (field k of face f of the chunk at
// stage[k * cstride + f]) so that K4a's stores and K4b's loads are coalesced.
//   [0..5] rhoL PL uL rhoR PR uR   (L = canonical endpoint a, u = velocity along the face normal)
//   [6..]  transverse velocities of a (D-1), of b (D-1), vFrame (D), A_ij (D)
// Staging traffic is written once and read once or twice by later kernels, never re-used from L2 before it is evicted
// (a chunk's record alone is several times the 126 MB L2).  Streaming (evict-first) accesses for it, so that it does
// not displace the pk1/pk2 records and F that ARE gathered repeatedly, changed nothing or LOST (A/B r2v, Sedov 61^3 /
// KH 1M / KH 4M: setup 0.142 -> 0.144, 0.995 -> 1.026, 3.97 -> 4.13 ms, every other kernel within 1 %): off
#ifndef MLH_STREAM_HINTS
#define MLH_STREAM_HINTS 0
#endif
#if MLH_STREAM_HINTS
#define MLH_ST_STAGE(ptr, val) __stcs((ptr), (val))
#define MLH_LD_STAGE(ptr) __ldcs(ptr)
#else
#define MLH_ST_STAGE(ptr, val) (*(ptr) = (val))
#define MLH_LD_STAGE(ptr) (*(ptr))
#endif
template <int D> struct FaceRec {
    static constexpr int NW = D + 2;
    static constexpr int RHOL = 0, PL = 1, UL = 2, RHOR = 3, PR = 4, UR = 5, VTA = 6, VTB = 6 + (D - 1), VF = 6 + 2 * (D - 1),
                         AA = 6 + 2 * (D - 1) + D, NREC = 2 * NW + 2 * D;
    // layout of the debug record (mlh_debug_fetch "face_rec": un-rotated, the reference's per-slot arrays)
    static constexpr int DBG_WA = 0, DBG_WB = NW, DBG_VF = 2 * NW, DBG_AA = 2 * NW + D;
};
template <int D>
__device__ __forceinline__ void face_store(double *rec, size_t fs, const double *Wa, const double *Wb, const double *vF, const double *A) {
    using R = FaceRec<D>;
    MLH_ST_STAGE(&rec[R::RHOL * fs], Wa[0]);
    MLH_ST_STAGE(&rec[R::PL * fs], Wa[1]);
    MLH_ST_STAGE(&rec[R::UL * fs], Wa[2]);
    MLH_ST_STAGE(&rec[R::RHOR * fs], Wb[0]);
    MLH_ST_STAGE(&rec[R::PR * fs], Wb[1]);
    MLH_ST_STAGE(&rec[R::UR * fs], Wb[2]);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        MLH_ST_STAGE(&rec[(R::VTA + k - 1) * fs], Wa[2 + k]);
        MLH_ST_STAGE(&rec[(R::VTB + k - 1) * fs], Wb[2 + k]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        MLH_ST_STAGE(&rec[(R::VF + k) * fs], vF[k]);
        MLH_ST_STAGE(&rec[(R::AA + k) * fs], A[k]);
    }
}

#define MLH_FACE_TILE 128

// ---------------------------------------------------------------------------------------------
// face list: the owner of every pair numbers the face and tells its partner
// ---------------------------------------------------------------------------------------------
// Thread per particle, four list slots per trip.  K2 left, in every OWNED slot, the rank of the slot among the owner's
// owned slots and where the partner j keeps the pair: j's list is ordered stencil cell by stencil cell and ascending
// inside a cell, so this particle sits there at slot grp[cell of i as j sees it][j] + r (r = particles of i's cell
// below i that list j, counted by K2's ballots).  The owner writes fa/fe, turns its own map entry into the global face
// index and stores the same index (with the "adds -F" bit) into the partner's slot: ONE gather (grp) and one scattered
// store per face, nothing for the slots a particle does not own.  [Before: every non-owned slot chased group start,
// member mask, list length and face base of its partner -- five dependent gathers, 0.10 ms at 61^3,
// profiles/r02_k_face_index_ncu_full.txt; a search of the partner's list cost 0.27-0.42 ms, profiles/r01h.]
// Periodic-image slots (few) and cells of more than 32 particles find the partner's slot by scanning its list.
#ifndef MLH_FI_TRIP
#define MLH_FI_TRIP 8 // list slots per trip (their loads and gathers are independent; 8 vs 4: 1.62 -> 1.58 ms at KH 4M, r3k)
#endif
#ifndef MLH_FI_STAGE
#define MLH_FI_STAGE 4096 // faces staged per block (32 KB): 128 particles x ~16 (3D) .. ~24 (2D) owned slots
#endif
template <bool PER>
__global__ void __launch_bounds__(128) k_face_index(const Params p) {
    // The faces of a block's 128 consecutive particles are one contiguous range of the face list: (owner, entry) pairs
    // are collected in shared memory and leave as full lines, instead of two 4-byte stores per face scattered over as
    // many sectors (k_face_index was LSU-queue bound on them: lg_throttle 19.6, profiles/r2c_k_face_index_scatter_*).
    __shared__ int s_fa[MLH_FI_STAGE], s_fe[MLH_FI_STAGE];
    const int i0 = p.own_begin + blockIdx.x * blockDim.x;
    const int i = i0 + threadIdx.x;
    const int ilast = min(i0 + (int)blockDim.x, p.own_end); // one past the block's last particle
    const int fbase = p.d.face_start[i0], fend = min(p.d.face_start[ilast], p.fcap);
    const int max_ni = p.max_ni;
    bool over = false;
    if (i < p.own_end) {
    const int nreg = p.d.noi[i], ntot = nreg + p.d.noig[i];
    const int fs = p.d.face_start[i];
    for (int s0 = 0; s0 < ntot; s0 += MLH_FI_TRIP) {
        unsigned v[MLH_FI_TRIP], g0[MLH_FI_TRIP];
        int e[MLH_FI_TRIP];
#pragma unroll
        for (int q = 0; q < MLH_FI_TRIP; ++q) {
            const int s = s0 + q < ntot ? s0 + q : ntot - 1;
            const size_t at = (size_t)s * p.ncap + i;
            v[q] = p.d.fmap[at];
            e[q] = p.d.nnl[at];
        }
#pragma unroll
        for (int q = 0; q < MLH_FI_TRIP; ++q) { // the four gathers of a trip are independent
            g0[q] = 0u;
            if ((v[q] & MLH_K2_OWNED) && !(v[q] & (MLH_K2_NOPARTNER | MLH_K2_GHOST)))
                g0[q] = p.d.grp[(size_t)((v[q] >> MLH_K2_SC_SHIFT) & 31u) * p.ncap + (e[q] & MLH_NNL_IDX_MASK)];
        }
#pragma unroll
        for (int q = 0; q < MLH_FI_TRIP; ++q) {
            const int s = s0 + q;
            if (s >= ntot) break;
            if (!(v[q] & MLH_K2_OWNED)) continue; // the partner owns the pair: it fills this slot
            const size_t at = (size_t)s * p.ncap + i;
            const unsigned sign = v[q] & 1u;
            const int f = fs + (int)((v[q] >> MLH_K2_RANK_SHIFT) & MLH_K2_RANK_MASK);
            if (f >= p.fcap) {
                over = true;
                p.d.fmap[at] = MLH_FMAP_SKIP;
                continue;
            }
            const int fav = i | (int)(sign << 31); // bit 31: the partner is the canonical endpoint (lower original index)
            if (f - fbase < MLH_FI_STAGE) {
                s_fa[f - fbase] = fav;
                s_fe[f - fbase] = e[q];
            } else { // more faces in this block than the stage holds
                p.d.fa[f] = fav;
                p.d.fe[f] = e[q];
            }
            p.d.fmap[at] = ((unsigned)f << 2) | MLH_K2_OWNED | sign;
            if (v[q] & MLH_K2_NOPARTNER) continue;
            const int j = e[q] & MLH_NNL_IDX_MASK;
            int t = -1;
            if (PER && (v[q] & MLH_K2_GHOST)) { // image of j in this list <-> image of i (opposite shift) in j's list
                const int want = i | (reverse_code((int)((unsigned)e[q] >> MLH_NNL_IDX_BITS)) << MLH_NNL_IDX_BITS);
                const int nr = p.d.noi[j], nt = nr + p.d.noig[j];
                for (int k = nr; k < nt; ++k)
                    if (p.d.nnl[(size_t)k * p.ncap + j] == want) {
                        t = k;
                        break;
                    }
            } else if (!(v[q] & MLH_K2_R_OVER)) {
                t = (int)g0[q] + (int)((v[q] >> MLH_K2_R_SHIFT) & 31u);
                if (t >= max_ni) t = -1; // j's list was cut at max_interactions before this pair
            } else {
                const int nr = p.d.noi[j];
                for (int k = (int)g0[q]; k < nr; ++k)
                    if (p.d.nnl[(size_t)k * p.ncap + j] == i) {
                        t = k;
                        break;
                    }
            }
            if (t >= 0) p.d.fmap[(size_t)t * p.ncap + j] = ((unsigned)f << 2) | 1u; // the partner adds -F
        }
    }
    }
    __syncthreads();
    const int nst = min(fend - fbase, MLH_FI_STAGE);
    for (int k = threadIdx.x; k < nst; k += blockDim.x) {
        p.d.fa[fbase + k] = s_fa[k];
        p.d.fe[fbase + k] = s_fe[k];
    }
    if (over) atomicOr(p.d.flags, MLH_F_MAX_INTERACTIONS);
}

#define MLH_PSTAR_VACUUM (-1.) // marker in the P* array: vacuum present or generated, solved in k_face_finish
template <int D>
__device__ __forceinline__ void face_load(const double *rec, size_t fs, double *Wa, double *Wb, double *vF, double *A) {
    using R = FaceRec<D>;
    Wa[0] = MLH_LD_STAGE(&rec[R::RHOL * fs]);
    Wa[1] = MLH_LD_STAGE(&rec[R::PL * fs]);
    Wa[2] = MLH_LD_STAGE(&rec[R::UL * fs]);
    Wb[0] = MLH_LD_STAGE(&rec[R::RHOR * fs]);
    Wb[1] = MLH_LD_STAGE(&rec[R::PR * fs]);
    Wb[2] = MLH_LD_STAGE(&rec[R::UR * fs]);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        Wa[2 + k] = MLH_LD_STAGE(&rec[(R::VTA + k - 1) * fs]);
        Wb[2 + k] = MLH_LD_STAGE(&rec[(R::VTB + k - 1) * fs]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        vF[k] = MLH_LD_STAGE(&rec[(R::VF + k) * fs]);
        A[k] = MLH_LD_STAGE(&rec[(R::AA + k) * fs]);
    }
}

// queue layout (chunk-sized, SoA): qd[k * cstride + q], k = 0..6 problem (rhoL, PL, aL, rhoR, PR, aR, du), 7..10 start of
// the root finder (Pguess, f(Pguess), f(0), f'(Pguess)), 11 = f' bound of the bracket (rs_brent_certain); qi[q] = face index relative to the chunk.  Faces that start
// with Newton-Raphson are appended from the front (q = 0, 1, ..), faces that start with Brent from the back
// (q = cstride-1, cstride-2, ..), so that the warps of k_face_iterate work on one kind at a time.  To keep the
// append counters from serialising (one L2 atomic per warp and kind), the queue is split into MLH_Q_REGIONS regions
// of equal capacity with their own pair of counters; the warp-tile of 32 faces number t appends to region t % REGIONS.
#define MLH_Q_FIELDS 12
#define MLH_Q_REGIONS 128
__host__ __device__ __forceinline__ int q_region_cap(int cstride) { // faces per region (multiple of 32)
    const int nwt = (cstride + 31) / 32;
    return (nwt + MLH_Q_REGIONS - 1) / MLH_Q_REGIONS * 32;
}
// The start of RiemannSolver::solve for one face (body of k_face_setup, also the tail of the fused k_face_states):
// sound speeds, vacuum test, initial guess, f(0), f(guess) of the one-dimensional problem in the face frame.  Faces
// that need no iteration get P* at once, the others are appended to the solver queue.  The whole warp must call it
// (ballots).
__device__ __forceinline__ void face_setup_and_queue(const Params &p, bool valid, int fl, double rhoL, double PL, double uL,
                                                     double rhoR, double PR, double uR,
                                                     double *__restrict__ pstar, double *__restrict__ qd, int *__restrict__ qi,
                                                     int *__restrict__ qcount, int rcap) {
    const int lane = threadIdx.x & 31;
    int method = RS_DONE;
    RsProblem q;
    RsIter it;
    if (valid) {
        // left = canonical particle a, right = b, along +A (Riemann.cpp:93-94)
        if (rs_setup(p.rs, rhoL, uL, PL, rhoR, uR, PR, q)) {
            rs_iter_begin(q, it);
            method = it.method;
            if (method == RS_DONE) MLH_ST_STAGE(pstar + fl, it.b);
        } else {
            MLH_ST_STAGE(pstar + fl, MLH_PSTAR_VACUUM);
        }
    }
    // one atomic instruction reserves the space of both kinds (lane 0: Newton starts, lane 1: Brent starts): the warp
    // waits for ONE L2 round trip before its queue stores, not two (the wait was 19 % of the kernel's stall samples,
    // profiles/r2t_k_face_setup_kh1000j_lines.txt)
    const unsigned mN = __ballot_sync(0xffffffffu, method == RS_NEWTON), mB = __ballot_sync(0xffffffffu, method == RS_BRENT);
    if (mN | mB) {
        const int region = (fl >> 5) % MLH_Q_REGIONS;
        const int cnt = lane == 0 ? __popc(mN) : __popc(mB);
        int base = 0;
        if (lane < 2 && cnt) base = atomicAdd(qcount + 2 * region + lane, cnt);
        const int baseN = __shfl_sync(0xffffffffu, base, 0), baseB = __shfl_sync(0xffffffffu, base, 1);
        if (method == RS_NEWTON || method == RS_BRENT) {
            const bool brent = method == RS_BRENT;
            int at = (brent ? baseB : baseN) + __popc((brent ? mB : mN) & ((1u << lane) - 1u));
            at = region * rcap + (brent ? rcap - 1 - at : at);
            double *d = qd + at;
            const size_t qs = (size_t)MLH_Q_REGIONS * rcap; // field stride of the queue
            MLH_ST_STAGE(d + 0 * qs, q.rhoL); MLH_ST_STAGE(d + 1 * qs, q.PL); MLH_ST_STAGE(d + 2 * qs, q.aL);
            MLH_ST_STAGE(d + 3 * qs, q.rhoR); MLH_ST_STAGE(d + 4 * qs, q.PR); MLH_ST_STAGE(d + 5 * qs, q.aR);
            MLH_ST_STAGE(d + 6 * qs, q.du);
            MLH_ST_STAGE(d + 7 * qs, brent ? it.a : it.fa);
            MLH_ST_STAGE(d + 8 * qs, it.b);
            MLH_ST_STAGE(d + 9 * qs, brent ? it.fa : it.fb);
            MLH_ST_STAGE(d + 10 * qs, brent ? it.fb : it.c);
            MLH_ST_STAGE(d + 11 * qs, it.fpb);
            MLH_ST_STAGE(qi + at, fl);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4a: one thread per face
// ---------------------------------------------------------------------------------------------
// resident blocks per SM (register cap): 3D needs 168 registers to run spill-free (3 blocks: 0.360 -> 0.339 ms at 61^3),
// 2D fits 128 and gains from the fourth block (KH 1M 0.86 ms vs 1.04 ms at 3) -- A/B r01v, profiles/README.md
#ifndef MLH_K4A_BLOCKS_2D
#define MLH_K4A_BLOCKS_2D 4
#endif
#ifndef MLH_K4A_BLOCKS_3D
#define MLH_K4A_BLOCKS_3D 3
#endif
#define MLH_K4A_BLOCKS_PER_SM(D) ((D) == 3 ? MLH_K4A_BLOCKS_3D : MLH_K4A_BLOCKS_2D)
#ifndef MLH_FUSE_SETUP
#define MLH_FUSE_SETUP 0
#endif
// L1 prefetch of the next trip's gather records: still a gain in 3D (0.273 -> 0.261 ms at 61^3), but since the records
// move with 256-bit loads it LOSES in 2D, where the kernel is bound by L1 look-ups and every prefetch is one more
// (A/B r3d, KH 1M / 4M: 0.764 -> 0.659, 3.05 -> 2.68 ms without it)
#ifndef MLH_K4A_PREFETCH
#define MLH_K4A_PREFETCH(D) ((D) == 3)
#endif
// FUSE (off): the solver setup (k_face_setup's body) runs on the states while they are still in registers, so the staged
// record is read once instead of twice.  Measured (A/B, profiles/README.md r01q): no gain at Sedov 61^3 (0.561 ms vs
// 0.370 + 0.188 ms) and a loss at KH 1M (2.52 vs 0.98 + 1.36 ms) -- 3x the spills and a 75 KB instruction footprint.
template <int D, bool PER, bool FUSE>
__global__ void __launch_bounds__(MLH_FACE_TILE, MLH_K4A_BLOCKS_PER_SM(D)) k_face_states(const Params p, double *__restrict__ stage, int f0, int cstride, double dt_fixed, double dt_max,
                                                                                      double *__restrict__ pstar, double *__restrict__ qd, int *__restrict__ qi, int *__restrict__ qcount) {
    constexpr int NW = D + 2;
    constexpr int PK1 = MLH_PK1(D), PK2 = MLH_PK2(D);
    using R = FaceRec<D>;
    const int nfaces = min(p.d.face_start[p.own_end], p.fcap);
    const int f1 = min(nfaces, f0 + cstride);
    const double dt = select_dt(p, dt_fixed, dt_max);
    const double gamma = p.gamma;
    const int nround = FUSE ? (f1 - f0 + 31) / 32 * 32 : f1 - f0; // FUSE: whole warps stay in the loop (ballots of the queue append)
    const int rcap = q_region_cap(cstride);
    // The face list is read one trip ahead, and the gather records of the NEXT trip's endpoints are prefetched into L1
    // while this trip computes: a block's 128 consecutive faces belong to ~8 neighbouring owners that share most of
    // their partners, ~50 distinct records (11 KB) per trip -- the kernel was waiting on exactly those first-touch
    // misses (long-scoreboard 5.9 of 10 stalled warps at 25 % occupancy, profiles/r01s).
    // [Copying the four records of the next trip to shared memory with cp.async (16-byte chunks, own slots, no barrier)
    // instead LOST by 2.4-2.7x (A/B r3c: 0.26 -> 0.62 ms at 61^3, 3.07 -> 8.39 ms at KH 4M): 20-34 scattered 16-byte
    // copies per thread cost more LSU issue than the 256-bit gathers they replace and take 40-70 KB of L1 per block.]
    const int stride = gridDim.x * MLH_FACE_TILE;
    int fl = blockIdx.x * MLH_FACE_TILE + threadIdx.x;
    int fav_next = 0, e_next = 0;
    if (fl < nround && f0 + fl < f1) {
        fav_next = p.d.fa[f0 + fl];
        e_next = p.d.fe[f0 + fl];
    }
    for (; fl < nround; fl += stride) {
        const int f = f0 + fl;
        const bool valid = f < f1;
        const int fav = fav_next, e = e_next;
        const bool valid_next = fl + stride < nround && f + stride < f1;
        if (valid_next) {
            fav_next = p.d.fa[f + stride];
            e_next = p.d.fe[f + stride];
        }
        double A[D], Wa[NW], Wb[NW];
        if (valid) {
        const int i = fav & 0x7FFFFFFF;
        const int j = e & MLH_NNL_IDX_MASK;
        const int code = PER ? (int)((unsigned)e >> MLH_NNL_IDX_BITS) : 0;
        // canonical orientation: the endpoint with the lower ORIGINAL index plays "i" (Particles.cpp:1841,1889);
        // K2 compared the ids when it marked the owner, k_face_index passed the result on in bit 31 of fa
        const bool canon = fav >= 0;
        const int ia = canon ? i : j, ib = canon ? j : i; // a = canonical endpoint, b = the other
        double A1[PK1], B1[PK1]; // packed records: x, v, rho, P, cs, omega
        load_packed<PK1>(p.d.pk1 + (size_t)ia * PK1, A1);
        load_packed<PK1>(p.d.pk1 + (size_t)ib * PK1, B1);
        double A2[PK2], B2[PK2]; // packed records: Binv, limited gradients (W order) -- requested together with A1/B1
        load_packed<PK2>(p.d.pk2 + (size_t)ia * PK2, A2);
        load_packed<PK2>(p.d.pk2 + (size_t)ib * PK2, B2);
        const double *xa = A1, *xb = B1, *va = A1 + D, *vb = B1 + D;
        const double rhoa = A1[2 * D], rhob = B1[2 * D], Pa = A1[2 * D + 1], Pb = B1[2 * D + 1];
        const double omga = A1[2 * D + 3], omgb = B1[2 * D + 3];

        // ---- geometry: b's image as a sees it, a's image as b sees it (identity for regular pairs) ----
        double xbi[D], xai[D];
        if (PER && code != 0) {
            const int cab = canon ? code : reverse_code(code); // code of b's image in a's list
            const int cba = reverse_code(cab);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xbi[k] = image_coord(xb[k], (cab >> (2 * k)) & 3, p.grid.bmin[k], p.grid.bmax[k]);
                xai[k] = image_coord(xa[k], (cba >> (2 * k)) & 3, p.grid.bmin[k], p.grid.bmax[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xbi[k] = xb[k];
                xai[k] = xa[k];
            }
        }

        // ---- effective face A_ab = psi~_b(x_a)/omega_a - psi~_a(x_b)/omega_b (Particles.cpp:1299-1302, :2525-2528) ----
        double ga[NW][D], gb[NW][D];
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
            const double ioa = 1. / omga, iob = 1. / omgb;
            const double psi1 = w1 * ioa, psi2 = w2 * iob;
#pragma unroll
            for (int al = 0; al < D; ++al) {
                double t1 = 0., t2 = 0.;
#pragma unroll
                for (int be = 0; be < D; ++be) {
                    t1 += A2[D * al + be] * d1[be] * psi1;
                    t2 += B2[D * al + be] * d2[be] * psi2;
                }
                A[al] = ioa * t1 - iob * t2;
            }
#pragma unroll
            for (int nu = 0; nu < NW; ++nu)
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    ga[nu][k] = A2[D * D + nu * D + k];
                    gb[nu][k] = B2[D * D + nu * D + k];
                }
        }

        if (MLH_K4A_PREFETCH(D) && valid_next) { // the list entries of the next trip have arrived by now; start its record fetches
            const int in = fav_next & 0x7FFFFFFF, jn = e_next & MLH_NNL_IDX_MASK;
            const char *r1i = (const char *)(p.d.pk1 + (size_t)in * PK1), *r1j = (const char *)(p.d.pk1 + (size_t)jn * PK1);
            const char *r2i = (const char *)(p.d.pk2 + (size_t)in * PK2), *r2j = (const char *)(p.d.pk2 + (size_t)jn * PK2);
#pragma unroll
            for (int o = 0; o < PK1 * 8; o += 128) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r1i + o));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r1j + o));
            }
            asm volatile("prefetch.global.L1 [%0];" ::"l"(r1i + PK1 * 8 - 8));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(r1j + PK1 * 8 - 8));
#pragma unroll
            for (int o = 0; o < PK2 * 8; o += 128) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r2i + o));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r2j + o));
            }
            asm volatile("prefetch.global.L1 [%0];" ::"l"(r2i + PK2 * 8 - 8));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(r2j + PK2 * 8 - 8));
        }

        // ---- boosted, reconstructed, predicted states (Particles.cpp:1498-1721; ghosts :2546-2672) ----
        double xjxi[3], xijxi[D], xijxj[D], vF[D];
        xjxi[2] = 0.; // quirk Q13 (ZERO_Z): never written in the first-order 3D branch
        if (!p.quad_h4) { // FIRST_ORDER_QUAD_POINT 1: face at the midpoint (:1504-1512,1526-1529,1550-1553)
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (k < 2 || p.q13_mode == MLH_Q13_GEOMETRIC) xjxi[k] = xbi[k] - xa[k];
                xijxj[k] = .5 * (xa[k] - xbi[k]);
                xijxi[k] = .5 * (xbi[k] - xa[k]);
                vF[k] = p.move_particles ? (va[k] + vb[k]) / 2. : 0.;
            }
        } else { // FIRST_ORDER_QUAD_POINT 0: face at x_a + h/4 (x_b - x_a), frame velocity interpolated to it (:1514-1523,1531-1537,1556-1563)
            double dotProd = 0., dSqr = 0.;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xjxi[k] = xbi[k] - xa[k];
                const double xij = xa[k] + p.h4 * xjxi[k];
                xijxi[k] = xij - xa[k];
                xijxj[k] = xij - xbi[k];
                dotProd += xijxi[k] * xjxi[k];
                dSqr += xjxi[k] * xjxi[k];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) vF[k] = p.move_particles ? va[k] + (vb[k] - va[k]) * dotProd / dSqr : 0.;
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
            const double ra = na / nab, rb = nb / nab;
#pragma unroll
            for (int nu = 0; nu < NW; ++nu) {
                const PairLimits lim = pairwise_limits(p, Wa0[nu], Wb0[nu]);
                const double wa = pairwise_limiter(lim, Wa[nu], Wa0[nu], Wb0[nu], ra);
                const double wb = pairwise_limiter(lim, Wb[nu], Wb0[nu], Wa0[nu], rb);
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
            // grad P / rho as a product with 1/rho: limited gradients are very often exactly 0, and 0/rho takes the
            // ~60-instruction special-operand path of the FP64 division (15 % of this kernel's instructions, profiles/r01o)
            const double ira = 1. / rhoa, irb = 1. / rhob;
            const double hdt = dt / 2.;
            Wa[0] -= hdt * (rhoa * aDiv + wa0 * ga[0][0] + wa1 * ga[0][1]);
            Wb[0] -= hdt * (rhob * bDiv + wb0 * gb[0][0] + wb1 * gb[0][1]);
            Wa[1] -= hdt * (gamma * Pa * aDiv + wa0 * ga[1][0] + wa1 * ga[1][1]);
            Wb[1] -= hdt * (gamma * Pb * bDiv + wb0 * gb[1][0] + wb1 * gb[1][1]);
            Wa[2] -= hdt * (ga[1][0] * ira + wa0 * ga[2][0] + wa1 * ga[2][1]);
            Wb[2] -= hdt * (gb[1][0] * irb + wb0 * gb[2][0] + wb1 * gb[2][1]);
            Wa[3] -= hdt * (ga[1][1] * ira + wa0 * ga[3][0] + wa1 * ga[3][1]);
            Wb[3] -= hdt * (gb[1][1] * irb + wb0 * gb[3][0] + wb1 * gb[3][1]);
            if (D == 3) {
                const double wa2 = va[D - 1] - vF[D - 1], wb2 = vb[D - 1] - vF[D - 1];
                const double wq3 = (p.q3_mode == MLH_Q3_FIXED) ? wb2 : wa2; // quirk Q3 (:1717,:1719)
                Wa[0] -= hdt * wa2 * ga[0][D - 1];
                Wb[0] -= hdt * wb2 * gb[0][D - 1];
                Wa[1] -= hdt * wa2 * ga[1][D - 1];
                Wb[1] -= hdt * wb2 * gb[1][D - 1];
                Wa[2] -= hdt * wa2 * ga[2][D - 1];
                Wb[2] -= hdt * wq3 * gb[2][D - 1];
                Wa[3] -= hdt * wa2 * ga[3][D - 1];
                Wb[3] -= hdt * wq3 * gb[3][D - 1];
                Wa[NW - 1] -= hdt * (ga[1][D - 1] * ira + wa0 * ga[NW - 1][0] + wa1 * ga[NW - 1][1] + wa2 * ga[NW - 1][D - 1]);
                Wb[NW - 1] -= hdt * (gb[1][D - 1] * irb + wb0 * gb[NW - 1][0] + wb1 * gb[NW - 1][1] + wb2 * gb[NW - 1][D - 1]);
            }
        }
        if (PER && code != 0 && (Wa[1] < 0. || Wb[1] < 0.)) atomicOr(p.d.flags, MLH_F_NEG_GHOST_PRESSURE);

        if (p.debug_capture) { // parity harness: the reference's per-slot WijR / WijL / vFrame / Aij of this face (mlh_debug_fetch "face_rec")
            double *dbg = p.d.dbg_face + (size_t)f * R::NREC;
#pragma unroll
            for (int nu = 0; nu < NW; ++nu) {
                dbg[R::DBG_WA + nu] = Wa[nu];
                dbg[R::DBG_WB + nu] = Wb[nu];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                dbg[R::DBG_VF + k] = vF[k];
                dbg[R::DBG_AA + k] = A[k];
            }
        }
        // ---- Riemann::Riemann (Riemann.cpp:7-81): velocities into the face frame, once, here -- the solver setup then
        // reads 6 of the record's 4D+4 fields and the finish kernel does not rotate again ----
        {
            FaceFrame<D> fr;
            face_rotate<D>(A, Wa, Wb, fr);
        }
        // ---- stage the record (field-major inside the chunk: coalesced here and in K4b) ----
        face_store<D>(stage + (f - f0), (size_t)cstride, Wa, Wb, vF, A);
        } // valid
        if (FUSE) face_setup_and_queue(p, valid, fl, Wa[0], Wa[1], Wa[2], Wb[0], Wb[1], Wb[2], pstar, qd, qi, qcount, rcap);
    }
}

// ---------------------------------------------------------------------------------------------
// K4b: the Riemann class of the reference, in three kernels so that every warp instruction has (nearly) all of its
// lanes doing the same thing:
//   k_face_setup    thread per face: rotate, sound speeds, vacuum test, initial guess, f(0), f(guess).  Faces that need
//                   no iteration (f(guess) == 0: identical states) get P* at once; the others append their problem
//                   (7 doubles) and root-finder state (7 doubles + flags) to a queue in HBM.
//   k_face_iterate  persistent lanes, one queued face per lane, state in registers, NO barriers and no shared
//                   memory: a lane iterates until its face converges, stores P*, and takes the next queue entry.
//                   The iteration counts are strongly bimodal (oracle statistics, KH 2D: ~70 % of the faces enter
//                   Brent, of those half take 3-6 iterations and a third 27-29 because Brent degenerates to bisection
//                   on [0, Pguess]) -- lane-per-face to completion left warps 28 % full and
//                   block-level regrouping (previous version, profiles/r01e..r01i) spent half of its warp time at
//                   barriers.  Here a lane is idle only during the ~50-instruction refill of its neighbours.
//   k_face_finish   thread per face: star state at x/t = 0, rotation back, projection -> F (canonical orientation).
// ---------------------------------------------------------------------------------------------
#define MLH_RS_FIELDS 14
#ifndef MLH_K4B_BLOCKS_PER_SM
#define MLH_K4B_BLOCKS_PER_SM 6
#endif


#ifndef MLH_SETUP_BLOCKS
#define MLH_SETUP_BLOCKS 8
#endif
#ifndef MLH_SETUP_PREFETCH
// 1: the six fields of the NEXT trip travel to shared memory (cp.async, no registers held) while this trip computes.
// [Requesting them into registers LOST (A/B r2g, Sedov 61^3 / KH 1M: 0.131 -> 0.143 ms, 1.259 -> 1.339 ms at 8
// blocks/SM with spills; 0.129 / 1.327 at 6 blocks without).]
#define MLH_SETUP_PREFETCH 1
#endif
template <int D>
__global__ void __launch_bounds__(MLH_FACE_TILE, MLH_SETUP_BLOCKS) k_face_setup(const Params p, const double *__restrict__ stage, double *__restrict__ pstar,
                                                              double *__restrict__ qd, int *__restrict__ qi, int *__restrict__ qcount,
                                                              int f0, int cstride) {
    const int nfaces = min(p.d.face_start[p.own_end], p.fcap);
    const int f1 = min(nfaces, f0 + cstride);
    const size_t fs = (size_t)cstride;
    const int nround = (f1 - f0 + 31) / 32 * 32; // whole warps stay in the loop (ballots in face_setup_and_queue)
    const int rcap = q_region_cap(cstride);
    using R = FaceRec<D>;
    const int stride = gridDim.x * MLH_FACE_TILE;
    int fl = blockIdx.x * MLH_FACE_TILE + threadIdx.x;
#if !MLH_SETUP_PREFETCH
    for (; fl < nround; fl += stride) { // (A/B variant: fields loaded when the trip starts)
        const bool valid = f0 + fl < f1;
        double w[6] = {1., 1., 0., 1., 1., 0.};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 6; ++k) w[k] = MLH_LD_STAGE(stage + k * fs + fl);
        }
        face_setup_and_queue(p, valid, fl, w[R::RHOL], w[R::PL], w[R::UL], w[R::RHOR], w[R::PR], w[R::UR], pstar, qd, qi, qcount, rcap);
    }
#else
    // every thread copies and reads only its own slots: no block barrier, the thread's own wait_prior is enough
    __shared__ double s_w[2][6][MLH_FACE_TILE];
    int buf = 0;
    if (fl < nround && f0 + fl < f1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(&s_w[0][k][threadIdx.x], stage + k * fs + fl, 8);
    }
    __pipeline_commit();
    for (; fl < nround; fl += stride) {
        const bool valid = f0 + fl < f1;
        const int fnext = fl + stride;
        if (fnext < nround && f0 + fnext < f1) {
#pragma unroll
            for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(&s_w[buf ^ 1][k][threadIdx.x], stage + k * fs + fnext, 8);
        }
        __pipeline_commit();
        __pipeline_wait_prior(1); // this trip's group has landed; the next trip's stays in flight
        double w[6] = {1., 1., 0., 1., 1., 0.};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 6; ++k) w[k] = s_w[buf][k][threadIdx.x];
        }
        face_setup_and_queue(p, valid, fl, w[R::RHOL], w[R::PL], w[R::UL], w[R::RHOR], w[R::PR], w[R::UR], pstar, qd, qi, qcount, rcap);
        buf ^= 1;
    }
#endif
}

// One queued face per lane, iteration state in registers, no block-level synchronisation.  Each warp owns a ring of
// 64 queue entries in shared memory that it fills 32 entries at a time with cp.async (one coalesced request per
// field, in flight while the lanes iterate); a lane whose face has converged stores P* and takes the next ring entry.
// [A per-lane refill straight from HBM fetched a 128-byte line for every 8-byte field: 37 GB of DRAM reads for a
// 3.7 GB queue and the whole warp waiting on them every pass -- profiles/r01j.]
#define MLH_RING 64
__global__ void __launch_bounds__(MLH_FACE_TILE, MLH_K4B_BLOCKS_PER_SM) k_face_iterate(const Params p, double *__restrict__ pstar, const double *__restrict__ qd,
                                                                                       const int *__restrict__ qi, const int *__restrict__ qcount, int cstride) {
    __shared__ double ring_d[MLH_FACE_TILE / 32][MLH_Q_FIELDS][MLH_RING];
    __shared__ int ring_i[MLH_FACE_TILE / 32][MLH_RING];
    __shared__ char ring_k[MLH_FACE_TILE / 32][MLH_RING]; // how the entry starts: RS_NEWTON | RS_BRENT
    // batches of <= 32 entries: first the Newton starts of all regions, then the Brent starts.  pre[k] = number of
    // batches before (kind, region) pair k = kind * REGIONS + region.
    __shared__ int pre[2 * MLH_Q_REGIONS + 1];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < 2 * MLH_Q_REGIONS; ++k) {
            pre[k] = acc;
            acc += (qcount[2 * (k % MLH_Q_REGIONS) + k / MLH_Q_REGIONS] + 31) / 32;
        }
        pre[2 * MLH_Q_REGIONS] = acc;
    }
    __syncthreads();
    const int NB = pre[2 * MLH_Q_REGIONS];
    const int rcap = q_region_cap(cstride);
    const size_t fs = (size_t)MLH_Q_REGIONS * rcap; // field stride of the queue
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    const int W = gridDim.x * (MLH_FACE_TILE / 32);
    int bq = blockIdx.x * (MLH_FACE_TILE / 32) + wib; // this warp's next batch (bq, bq + W, ..)
    int head = 0, avail = 0, pend = 0;                // ring: first ready slot, ready entries, entries in flight
    double (*rd)[MLH_RING] = ring_d[wib];
    int *ri = ring_i[wib];
    char *rk = ring_k[wib];
    int face = -1;
    RsProblem q;
    q.rhoL = q.PL = q.aL = q.rhoR = q.PR = q.aR = 1.;
    q.du = q.Pguess = q.fPguess = q.f0 = q.fpsum = 0.;
    q.iPL = q.iPR = 1.;
    RsIter it;
    it.method = RS_DONE;
    it.mflag = 0;
    it.a = it.b = it.c = it.d = it.fa = it.fb = it.fc = it.fpb = 0.;
    for (;;) {
        // ---- prefetch the next batch while at most half of the ring is occupied ----
        if (pend == 0 && avail <= MLH_RING - 32 && bq < NB) {
            int k = 0; // largest k with pre[k] <= bq
#pragma unroll
            for (int step = MLH_Q_REGIONS; step > 0; step >>= 1)
                if (k + step <= 2 * MLH_Q_REGIONS - 1 && pre[k + step] <= bq) k += step;
            const bool newton = k < MLH_Q_REGIONS;
            const int region = newton ? k : k - MLH_Q_REGIONS;
            const int e = 32 * (bq - pre[k]) + lane;
            const int cnt = min(32, qcount[2 * region + (newton ? 0 : 1)] - (e - lane));
            if (lane < cnt) {
                const int src = region * rcap + (newton ? e : rcap - 1 - e);
                const int slot = (head + avail + lane) & (MLH_RING - 1);
#pragma unroll
                for (int k = 0; k < MLH_Q_FIELDS; ++k) __pipeline_memcpy_async(&rd[k][slot], qd + k * fs + src, 8);
                __pipeline_memcpy_async(&ri[slot], qi + src, 4);
                rk[slot] = newton ? RS_NEWTON : RS_BRENT;
            }
            __pipeline_commit();
            pend = cnt;
            bq += W;
        }
        // ---- lanes whose face has converged store P* and take the next ring entry ----
        const bool done = it.method == RS_DONE;
        if (done && face >= 0) {
            MLH_ST_STAGE(pstar + face, it.b);
            face = -1;
        }
        const unsigned need = __ballot_sync(0xffffffffu, done);
        if (need) {
            const int n = __popc(need);
            if (avail < n && pend) {
                __pipeline_wait_prior(0);
                __syncwarp();
                avail += pend;
                pend = 0;
            }
            const int rank = __popc(need & below);
            if (done && rank < avail) {
                const int slot = (head + rank) & (MLH_RING - 1);
                q.rhoL = rd[0][slot]; q.PL = rd[1][slot]; q.aL = rd[2][slot]; q.rhoR = rd[3][slot]; q.PR = rd[4][slot];
                q.aR = rd[5][slot]; q.du = rd[6][slot];
                // reciprocals by two Newton steps from the hardware estimate (<= 1 ulp; 6 instructions instead of the ~25 of
                // an IEEE division -- this refill runs with ~5 of 32 lanes and the two divisions were 5 % of the kernel's
                // warp instructions on KH, profiles/r2k_iterate_lines_kh1000j.txt).  Both code paths of the iteration use
                // the same iPL / iPR, so results stay independent of the queue order.
                q.iPL = rs_rcp(q.PL);
                q.iPR = rs_rcp(q.PR);
                const double v0 = rd[7][slot], v1 = rd[8][slot], v2 = rd[9][slot], v3 = rd[10][slot];
                it.fpb = rd[11][slot];
                face = ri[slot];
                const bool nw = rk[slot] == RS_NEWTON;
                it.method = nw ? RS_NEWTON : RS_BRENT;
                it.mflag = nw ? 0 : 1;
                it.a = nw ? 0. : v0;  // Newton: Pstar = 0, f(0); Pguess, f(Pguess), f'(Pguess)
                it.fa = nw ? v0 : v2; // Brent: a, b, fa, fb as rs_brent_begin left them, c = a, fc = fa
                it.b = v1;
                it.fb = nw ? v2 : v3;
                it.c = nw ? v3 : v0;
                it.fc = nw ? 0. : v2;
                it.d = nw ? 0. : 1e230;
            }
            const int taken = min(n, avail);
            head = (head + taken) & (MLH_RING - 1);
            avail -= taken;
            __syncwarp(); // the slots just read may be overwritten by the next prefetch
        }
        const bool active = it.method != RS_DONE;
        if (!__any_sync(0xffffffffu, active)) {
            if (pend == 0 && avail == 0 && bq >= NB) break;
            continue;
        }
        if (active) {
            if (__any_sync(__activemask(), it.method == RS_NEWTON)) { // Newton starts come first, batch-wise
                const double trial = rs_iter_trial(it);
                RsEval e;
                rs_eval2<true, true>(p.rs, q.rhoL, q.PL, q.aL, q.rhoR, q.PR, q.aR, trial, e, q.iPL, q.iPR);
                rs_iter_update(it, trial, e.fL + e.fR + q.du, e.fpL + e.fpR);
            } else {
                rs_brent_step(p.rs, q, it);
            }
        }
    }
}

// 8 resident blocks (<= 64 registers, a few spilled values): HBM-bound streaming, more loads in flight win
// (r01z: KH 1M 0.78 -> 0.71 ms, 61^3 0.126 -> 0.123 ms against 6 blocks at 80 registers)
#ifndef MLH_FINISH_BLOCKS
#define MLH_FINISH_BLOCKS 8
#endif
template <int D>
__global__ void __launch_bounds__(MLH_FACE_TILE, MLH_FINISH_BLOCKS) k_face_finish(const Params p, const double *__restrict__ stage, const double *__restrict__ pstar,
                                                               int f0, int cstride) {
    constexpr int NW = D + 2;
    constexpr int FREC = MLH_FREC(D);
    const int nfaces = min(p.d.face_start[p.own_end], p.fcap);
    const int f1 = min(nfaces, f0 + cstride);
    const size_t fs = (size_t)cstride;
    bool vacuum = false;
    for (int f = f0 + blockIdx.x * MLH_FACE_TILE + threadIdx.x; f < f1; f += gridDim.x * MLH_FACE_TILE) {
        double Wa[NW], Wb[NW], vF[D], A[D], F[NW];
        FaceFrame<D> fr;
        face_load<D>(stage + (f - f0), fs, Wa, Wb, vF, A); // velocities already in the face frame (K4a)
        face_frame<D>(A, fr);
        const double Ps = MLH_LD_STAGE(pstar + (f - f0));
        double rhoSol, uSol, PSol;
        int flag;
        if (Ps == MLH_PSTAR_VACUUM) {
            flag = rs_solve_vacuum(p.rs, Wa[0], Wa[2], Wa[1], Wb[0], Wb[2], Wb[1], &rhoSol, &uSol, &PSol);
            vacuum = vacuum || flag == 0;
        } else {
            RsProblem q;
            q.rhoL = Wa[0]; q.PL = Wa[1]; q.rhoR = Wb[0]; q.PR = Wb[1];
            q.aL = sqrt(p.rs.gamma * q.PL / q.rhoL); // as rs_setup
            q.aR = sqrt(p.rs.gamma * q.PR / q.rhoR);
            flag = rs_sample(p.rs, q, Wa[2], Wb[2], Ps, &rhoSol, &uSol, &PSol);
        }
        face_project<D>(p, flag, rhoSol, uSol, PSol, Wa, Wb, fr, vF, A, F);
        double Fr[FREC];
#pragma unroll
        for (int nu = 0; nu < FREC; ++nu) Fr[nu] = nu < NW ? F[nu] : 0.;
        store_packed<FREC>(p.d.F + (size_t)f * FREC, Fr);
    }
    if (vacuum) atomicOr(p.d.flags, MLH_F_VACUUM);
}

// ---------------------------------------------------------------------------------------------
// K4c / K5: collectFluxes (:1926-2008) in list order + updateStateAndPosition (:2013-2110)
// ---------------------------------------------------------------------------------------------
template <int D, bool PER>
__global__ void __launch_bounds__(128) k_flux_sum_update(const Params p, double dt_fixed, double dt_max) {
    constexpr int NW = D + 2;
    constexpr int FREC = MLH_FREC(D);
    const int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    // non-periodic runs rebuild the search grid from the particle bounding box every step (MeshlessScheme.cpp:41-51):
    // the box of the NEW positions is reduced here (getDomainLimits, quirks Q2/Q8; initialised by k_density_matrix), so
    // the next step starts without a pass over the particles
    double bmn[D], bmx[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        bmn[k] = DBL_MAX;
        bmx[k] = DBL_MIN;
    }
    const double dt = select_dt(p, dt_fixed, dt_max);
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.d.dt_used = dt;
    if (i < p.own_end) {
    const int ntot = p.d.noi[i] + p.d.noig[i];
    double acc[NW];
#pragma unroll
    for (int nu = 0; nu < NW; ++nu) acc[nu] = 0.;
    // four slots per trip: the map entries and then the four flux records are fetched together (independent gathers),
    // the sums still run in list order
    for (int s0 = 0; s0 < ntot; s0 += 4) {
        unsigned v[4];
        double Fr[4][FREC];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = s0 + q < ntot ? p.d.fmap[(size_t)(s0 + q) * p.ncap + i] : MLH_FMAP_SKIP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = (int)(v[q] >> 2);
            if (v[q] == MLH_FMAP_SKIP || f >= p.fcap) { // no face (the other side's list was cut at capacity)
                v[q] = MLH_FMAP_SKIP;
                continue;
            }
            load_packed<FREC>(p.d.F + (size_t)f * FREC, Fr[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (v[q] == MLH_FMAP_SKIP) continue;
            const double sgn = (v[q] & 1u) ? -1. : 1.;
#pragma unroll
            for (int nu = 0; nu < NW; ++nu) acc[nu] += sgn * Fr[q][nu];
        }
    }
    if (p.debug_capture) {
        p.d.flux[0][i] = acc[0];
        p.d.flux[1][i] = acc[1];
#pragma unroll
        for (int k = 0; k < D; ++k) p.d.flux[2 + k][i] = acc[2 + k];
    }
    double xs[D], vs[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        xs[k] = p.d.x[k][i];
        vs[k] = p.d.v[k][i];
    }
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
        xs[k] = x;
    }
    p.d.cm[o] = m;
    p.d.cu[o] = u;
    const int id = p.d.id[i];
    p.d.cid[o] = id;
    if (!PER) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (xs[k] < bmn[k]) bmn[k] = xs[k];
            if (id != 0) {
                if (xs[k] > bmx[k]) bmx[k] = xs[k];
            } else {
                p.d.bbox[6 + k] = xs[k];
            }
            if (id == 1) p.d.bbox[9 + k] = xs[k];
        }
    }
    }
    if (!PER) mlh_bbox_block_reduce<D>(p, bmn, bmx);
}

template <int D, bool PER>
int launch_chunks(mlh_ctx *c, double dt_fixed, double dt_max) {
    Params &p = c->p;
    const int n = p.own_end - p.own_begin;
    const int chunk = c->stage_chunk;
    const int grid_persistent = c->num_sms * 8;
    // blocks per SM of the streaming face kernels (setup / finish): tunable for A/B runs
    // persistent grids = two clean waves of the blocks each kernel keeps resident (register budgets pinned by
    // __launch_bounds__); sweeps: tools/grid_sweep.sh, profiles/README.md r01t / r01w / r01z
    static const int g_setup = getenv("MLH_GRID_SETUP") ? atoi(getenv("MLH_GRID_SETUP")) : 24; // 3 waves of 8: 0.155 -> 0.152 / 1.335 -> 1.302 ms (r02h)
    static const int g_finish = getenv("MLH_GRID_FINISH") ? atoi(getenv("MLH_GRID_FINISH")) : 2 * MLH_FINISH_BLOCKS;
    // K4a: two clean waves of its resident blocks (3 per SM in 3D, 4 in 2D): 0.339 -> 0.309 ms at 61^3 (r01w)
    static const int g_states = getenv("MLH_GRID_STATES") ? atoi(getenv("MLH_GRID_STATES")) : 2 * MLH_K4A_BLOCKS_PER_SM(D);
    // Riemann iteration: resident blocks per SM it is launched with (a smaller grid leaves room for kernels of another
    // stream, tools/overlap_probe.py)
    static const int g_iter = getenv("MLH_GRID_ITERATE") ? max(1, min(MLH_K4B_BLOCKS_PER_SM, atoi(getenv("MLH_GRID_ITERATE")))) : MLH_K4B_BLOCKS_PER_SM;
    cudaStream_t st = c->stream;
    // The face count lives on the device (face_start[own_end]); without a host round trip the chunk loop covers the
    // face CAPACITY and the kernels of chunks beyond the last face return at once.  One chunk in the usual case.
    for (long f0 = 0; f0 < p.fcap; f0 += chunk) {
        double *pstar = c->stage + (size_t)FaceRec<D>::NREC * chunk;
        double *qd = pstar + chunk;
        const size_t qs = (size_t)MLH_Q_REGIONS * q_region_cap(chunk);
        int *qi = (int *)(qd + (size_t)MLH_Q_FIELDS * qs);
        int *qcount = qi + qs;
        cudaMemsetAsync(qcount, 0, 2 * MLH_Q_REGIONS * sizeof(int), st);
        mlh_prof_begin(c, KID_FACES);
        k_face_states<D, PER, MLH_FUSE_SETUP != 0><<<c->num_sms * g_states, MLH_FACE_TILE, 0, st>>>(p, c->stage, (int)f0, chunk, dt_fixed, dt_max, pstar, qd, qi, qcount);
        mlh_prof_end(c, KID_FACES);
        if (!MLH_FUSE_SETUP) {
            mlh_prof_begin(c, KID_FLUX_SETUP);
            k_face_setup<D><<<c->num_sms * g_setup, MLH_FACE_TILE, 0, st>>>(p, c->stage, pstar, qd, qi, qcount, (int)f0, chunk);
            mlh_prof_end(c, KID_FLUX_SETUP);
        }
        mlh_prof_begin(c, KID_FLUX);
        k_face_iterate<<<c->num_sms * g_iter, MLH_FACE_TILE, 0, st>>>(p, pstar, qd, qi, qcount, chunk);
        mlh_prof_end(c, KID_FLUX);
        mlh_prof_begin(c, KID_FLUX_FINISH);
        k_face_finish<D><<<c->num_sms * g_finish, MLH_FACE_TILE, 0, st>>>(p, c->stage, pstar, (int)f0, chunk);
        mlh_prof_end(c, KID_FLUX_FINISH);
    }
    mlh_prof_begin(c, KID_UPDATE);
    k_flux_sum_update<D, PER><<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, st>>>(p, dt_fixed, dt_max);
    mlh_prof_end(c, KID_UPDATE);
    return MLH_OK;
}

} // namespace

// face list of this step (after K2): exclusive scan of the owned-slot counts + k_face_index
int mlh_launch_face_index(mlh_ctx *c) {
    Params &p = c->p;
    const int n = p.own_end - p.own_begin;
    if (n <= 0) return MLH_OK;
    mlh_prof_begin(c, KID_FACE_INDEX);
    // face_start is indexed by the SRT index; entries below own_begin are never read
    mlh_exclusive_scan(c, p.d.nown + p.own_begin, p.d.face_start + p.own_begin, p.d.face_scan_tmp, n);
    if (p.periodic)
        k_face_index<true><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else
        k_face_index<false><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_FACE_INDEX);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}

// staging buffer per face of one chunk: record (4D+4 doubles), P*, solver queue (MLH_Q_FIELDS = 12 doubles + 1 int)
int mlh_stage_alloc(mlh_ctx *c) {
    Params &p = c->p;
    const size_t per_face = (size_t)(4 * p.D + 4 + 1 + MLH_Q_FIELDS) * sizeof(double) + sizeof(int);
    // budget: mlh_config.stage_bytes, else an eighth of the device memory (22 GB of B200's 180 GB: KH 4 M = 95 M faces in
    // one chunk), at most a third of what was free when first asked, never below 1 GiB.  Asked ONCE per context:
    // cudaMemGetInfo costs milliseconds of host time (it stalled every step of the small workloads, visit r2z).
    if (c->stage_budget == 0) {
        size_t budget0 = (size_t)6 << 30;
        if (c->cfg.stage_bytes > 0) {
            budget0 = (size_t)c->cfg.stage_bytes;
        } else {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                budget0 = total_b / 8;
                if (budget0 > free_b / 3) budget0 = free_b / 3;
                if (budget0 < ((size_t)1 << 30)) budget0 = (size_t)1 << 30;
            } else {
                cudaGetLastError();
            }
        }
        c->stage_budget = budget0;
    }
    const size_t budget = c->stage_budget;
    long chunk = (long)(budget / per_face);
    chunk = chunk / MLH_FACE_TILE * MLH_FACE_TILE;
    if (chunk < 4 * MLH_FACE_TILE) chunk = 4 * MLH_FACE_TILE;
    long need = ((long)p.fcap + MLH_FACE_TILE - 1) / MLH_FACE_TILE * MLH_FACE_TILE;
    if (chunk > need) chunk = need;
    if (chunk > (1L << 28) - MLH_FACE_TILE) chunk = (1L << 28) - MLH_FACE_TILE; // queue entries pack the face index into 28 bits
    if (c->stage && chunk == c->stage_chunk) return MLH_OK;
    if (c->stage) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->stage);
        c->stage = nullptr;
    }
    // the queue regions round the chunk up to a multiple of 32 * MLH_Q_REGIONS faces
    cudaError_t e = cudaMalloc(&c->stage, per_face * ((size_t)chunk + 32 * MLH_Q_REGIONS) + 2 * MLH_Q_REGIONS * sizeof(int) + 256);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "cudaMalloc of the %.2f GB face staging buffer failed: %s (lower mlh_config.stage_bytes)",
                 per_face * (double)chunk / 1e9, cudaGetErrorString(e));
        cudaGetLastError();
        return MLH_E_CUDA;
    }
    c->stage_chunk = (int)chunk;
    return MLH_OK;
}

int mlh_launch_flux(mlh_ctx *c, double dt_fixed, double dt_max) {
    Params &p = c->p;
    int n = p.own_end - p.own_begin;
    int rc = mlh_stage_alloc(c);
    if (rc != MLH_OK) return rc;
    if (p.D == 2 && p.periodic)
        launch_chunks<2, true>(c, dt_fixed, dt_max);
    else if (p.D == 2)
        launch_chunks<2, false>(c, dt_fixed, dt_max);
    else if (p.periodic)
        launch_chunks<3, true>(c, dt_fixed, dt_max);
    else
        launch_chunks<3, false>(c, dt_fixed, dt_max);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    p.ncur = n;
    c->bbox_valid = !p.periodic; // reduced by k_flux_sum_update
    return MLH_OK;
}

// ---------------------------------------------------------------------------------------------
// stand-alone face solver (mlh_riemann_faces): the Riemann class of the reference for a batch of faces
// ---------------------------------------------------------------------------------------------
namespace {
template <int D>
__global__ void k_pack_faces(const double *__restrict__ WL, const double *__restrict__ WR, const double *__restrict__ vF,
                             const double *__restrict__ A, double *__restrict__ stage, int n, int cstride) {
    using R = FaceRec<D>;
    constexpr int NW = D + 2;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double Wa[NW], Wb[NW], v[D], a[D];
    for (int nu = 0; nu < NW; ++nu) {
        Wa[nu] = WL[(size_t)f * NW + nu];
        Wb[nu] = WR[(size_t)f * NW + nu];
    }
    for (int k = 0; k < D; ++k) {
        v[k] = vF[(size_t)f * D + k];
        a[k] = A[(size_t)f * D + k];
    }
    FaceFrame<D> fr;
    face_rotate<D>(a, Wa, Wb, fr); // Riemann::Riemann, as K4a does for the faces of a step
    face_store<D>(stage + f, (size_t)cstride, Wa, Wb, v, a);
}
template <int D>
__global__ void k_unpack_fluxes(const double *__restrict__ F, double *__restrict__ out, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    for (int nu = 0; nu < D + 2; ++nu) out[(size_t)f * (D + 2) + nu] = F[(size_t)f * MLH_FREC(D) + nu];
}

template <int D>
int riemann_faces(mlh_ctx *c, int n, const double *hWL, const double *hWR, const double *hvF, const double *hA, double *hF) {
    constexpr int NW = D + 2;
    cudaStream_t st = c->stream;
    const int chunk = (n + MLH_FACE_TILE - 1) / MLH_FACE_TILE * MLH_FACE_TILE;
    const size_t qs = (size_t)MLH_Q_REGIONS * q_region_cap(chunk);
    const size_t nd_in = (size_t)n * (2 * NW + 2 * D);
    const size_t doubles = nd_in + (size_t)FaceRec<D>::NREC * chunk + chunk + MLH_Q_FIELDS * qs + (size_t)n * MLH_FREC(D) + (size_t)n * NW;
    const size_t bytes = doubles * sizeof(double) + (qs + 2 * MLH_Q_REGIONS + 8) * sizeof(int);
    // the host mirror of the reference's Riemann class calls this once per face (Riemann.h:19,26): the device scratch is
    // kept between calls and only grows
    if (bytes + 16 > c->rf_bytes) {
        if (c->rf_buf) {
            cudaStreamSynchronize(st);
            cudaFree(c->rf_buf);
            c->rf_buf = nullptr;
            c->rf_bytes = 0;
        }
        const size_t want = bytes + 16 < ((size_t)1 << 20) ? ((size_t)1 << 20) : bytes + 16;
        MLH_CUDA_CHECK(c, cudaMalloc(&c->rf_buf, want));
        c->rf_bytes = want;
    }
    char *buf = c->rf_buf;
    double *dWL = (double *)buf, *dWR = dWL + (size_t)n * NW, *dvF = dWR + (size_t)n * NW, *dA = dvF + (size_t)n * D;
    double *stage = dA + (size_t)n * D;
    double *pstar = stage + (size_t)FaceRec<D>::NREC * chunk;
    double *qd = pstar + chunk;
    double *F = qd + MLH_Q_FIELDS * qs;
    double *out = F + (size_t)n * MLH_FREC(D);
    int *qi = (int *)(out + (size_t)n * NW);
    int *qcount = qi + qs;
    int *nfaces = qcount + 2 * MLH_Q_REGIONS;
    cudaMemcpyAsync(dWL, hWL, sizeof(double) * n * NW, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dWR, hWR, sizeof(double) * n * NW, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dvF, hvF, sizeof(double) * n * D, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dA, hA, sizeof(double) * n * D, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(qcount, 0, 2 * MLH_Q_REGIONS * sizeof(int), st);
    cudaMemcpyAsync(nfaces, &n, sizeof(int), cudaMemcpyHostToDevice, st);
    // a parameter block whose face list is "n faces, flux array F"
    Params p = c->p;
    p.own_end = 0;
    p.d.face_start = nfaces;
    p.fcap = n;
    p.d.F = F;
    if (!p.d.flags) { // context without particles: private flag word at the end of the scratch
        unsigned *flags = (unsigned *)(buf + ((bytes + 7) / 8 * 8));
        cudaMemsetAsync(flags, 0, sizeof(unsigned), st);
        p.d.flags = flags;
    }
    const int grid = mlh_blocks(n, MLH_FACE_TILE);
    k_pack_faces<D><<<grid, MLH_FACE_TILE, 0, st>>>(dWL, dWR, dvF, dA, stage, n, chunk);
    k_face_setup<D><<<grid, MLH_FACE_TILE, 0, st>>>(p, stage, pstar, qd, qi, qcount, 0, chunk);
    k_face_iterate<<<c->num_sms * MLH_K4B_BLOCKS_PER_SM, MLH_FACE_TILE, 0, st>>>(p, pstar, qd, qi, qcount, chunk);
    k_face_finish<D><<<grid, MLH_FACE_TILE, 0, st>>>(p, stage, pstar, 0, chunk);
    k_unpack_fluxes<D><<<grid, MLH_FACE_TILE, 0, st>>>(F, out, n);
    c->launches += 5;
    cudaMemcpyAsync(hF, out, sizeof(double) * n * NW, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "mlh_riemann_faces: %s", cudaGetErrorString(e));
        return MLH_E_CUDA;
    }
    return MLH_OK;
}
} // namespace

extern "C" int mlh_riemann_faces(mlh_ctx *c, long n, const double *WR, const double *WL, const double *vFrame, const double *Aij,
                                 double *Fij) {
    if (!c || n < 0 || (n > 0 && (!WR || !WL || !vFrame || !Aij || !Fij))) return MLH_E_INVALID;
    if (n == 0) return MLH_OK;
    if (n > (1L << 27)) {
        snprintf(c->err, sizeof(c->err), "mlh_riemann_faces: at most 2^27 faces per call");
        return MLH_E_INVALID;
    }
    cudaSetDevice(c->cfg.device);
    return c->p.D == 2 ? riemann_faces<2>(c, (int)n, WL, WR, vFrame, Aij, Fij) : riemann_faces<3>(c, (int)n, WL, WR, vFrame, Aij, Fij);
}
