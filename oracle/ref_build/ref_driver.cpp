/*
 * oracle/ref_build/ref_driver.cpp -- TEST INFRASTRUCTURE (reference-source oracle, oracle/_ref).
 *
 * C-ABI driver around the UNMODIFIED reference sources
 *   /root/reference/demonstrator/src/{Particles,Domain,Riemann,Helper,Logger}.cpp
 * compiled where they lie (see oracle/ref_build/Makefile).  It replaces main.cpp /
 * ConfigParser / InitialDistribution (which need cxxopts, Boost and HDF5, all absent
 * here) and replays the body of MeshlessScheme::run() (MeshlessScheme.cpp:39-253)
 * phase by phase through the reference's own public Particles/Domain methods, copying
 * intermediates out of the (force-opened) private members between phases.
 * Nothing here computes physics.
 */
#include "Particles.h"
#include "Domain.h"
#include "Helper.h"
#include "Logger.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <vector>

structlog LOGCFG = {}; // main.cpp:14 (default level WARN)

#ifndef MESHLESS_FINITE_MASS /* absent from the testcase headers; `#if` then reads it as 0 */
#define MESHLESS_FINITE_MASS 0
#endif

namespace {

struct RefCtx {
    int N;
    double h, gamma;
    Particles *p;
    Particles *ghosts; // periodic: DIM*N ghost holder (MeshlessScheme.cpp:12); else dummy
    Domain *domain;
    Helper helper;
    double dt;     // last dt used
    double dtCfl;  // last compGlobalTimestep result
    int capture;   // copy pre-limiter gradients
    std::vector<double> gradPre; // [field][N][DIM], fields rho,vx,vy,(vz),P
    double phaseSec[16];
};

void copyGrad(std::vector<double> &dst, size_t off, double (*g)[DIM], int N) {
    std::memcpy(dst.data() + off, &g[0][0], sizeof(double) * DIM * N);
}

} // namespace

extern "C" {

/* compile-time configuration of this build (parameter.h values) */
void ref_info(int *iv, double *dv) {
    iv[0] = DIM;
    iv[1] = PERIODIC_BOUNDARIES;
    iv[2] = MAX_NUM_INTERACTIONS;
    iv[3] = MAX_NUM_GHOST_INTERACTIONS;
    iv[4] = PAIRWISE_LIMITER;
    iv[5] = SLOPE_LIMITING;
    iv[6] = MESHLESS_FINITE_MASS;
    iv[7] = ENFORCE_FLUX_SYM;
    iv[8] = MOVE_PARTICLES;
    iv[9] = FIRST_ORDER_QUAD_POINT;
    iv[10] = ADAPTIVE_TIMESTEP;
#ifdef MLH_REF_FABS
    iv[11] = 1;
#else
    iv[11] = 0;
#endif
    dv[0] = CFL;
    dv[1] = BETA;
    dv[2] = PSI_1;
    dv[3] = PSI_2;
}

void *ref_create(int N, const double *x, const double *y, const double *z, const double *vx,
                 const double *vy, const double *vz, const double *m, const double *u,
                 const double *box /* [minX,minY(,minZ),maxX,maxY(,maxZ)] periodic only */, double h,
                 double gamma) {
    RefCtx *c = new RefCtx();
    c->N = N;
    c->h = h;
    c->gamma = gamma;
    c->dt = 0.;
    c->dtCfl = 0.;
    c->capture = 1;
    c->p = new Particles(N); // main.cpp:90
    for (int i = 0; i < N; ++i) { // InitialDistribution.cpp:41-59
        c->p->m[i] = m[i];
        c->p->u[i] = u[i];
        c->p->matId[i] = 0;
        c->p->x[i] = x[i];
        c->p->vx[i] = vx[i];
        c->p->y[i] = y[i];
        c->p->vy[i] = vy[i];
#if DIM == 3
        c->p->z[i] = z[i];
        c->p->vz[i] = vz[i];
#endif
    }
    double limits[2 * DIM];
#if PERIODIC_BOUNDARIES
    for (int k = 0; k < 2 * DIM; ++k) limits[k] = box[k]; // main.cpp:96-97
    c->ghosts = new Particles(DIM * N, true);             // MeshlessScheme.cpp:12
#else
    c->p->getDomainLimits(limits); // main.cpp:99-100
    c->ghosts = nullptr;
#endif
    Domain::Cell bb{limits};
    c->domain = new Domain(bb);
    c->domain->createGrid(h); // MeshlessScheme.cpp:17
    c->gradPre.assign((size_t)(DIM + 2) * DIM * N, 0.);
    return c;
}

void ref_destroy(void *ctx) {
    RefCtx *c = (RefCtx *)ctx;
    delete c->p;
    delete c->ghosts;
    delete c->domain;
    delete c;
}

void ref_set_capture(void *ctx, int on) { ((RefCtx *)ctx)->capture = on; }

/*
 * One pass of the loop body of MeshlessScheme::run() (MeshlessScheme.cpp:39-253), minus
 * logging, the snapshot dump (:165-195) and the dump-time clipping of dt (:94-101), which
 * are driver policy.  dtFixed >= 0 -> use it (ADAPTIVE_TIMESTEP 0 semantics, :104); else
 * dt = compGlobalTimestep (:93), optionally clipped to dtMax if dtMax > 0.
 * stopAfter: 0 = full step; 1 = stop before solveRiemannProblems (state as at the dump point).
 * Returns dt used.
 */
double ref_step(void *ctx, double dtFixed, double dtMax, int stopAfter) {
    RefCtx *c = (RefCtx *)ctx;
    Particles *particles = c->p;
    Domain &domain = *c->domain;
    const int N = c->N;
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    auto lap = [&](int k) {
        auto t1 = clk::now();
        c->phaseSec[k] = std::chrono::duration<double>(t1 - t0).count();
        t0 = t1;
    };
#if !PERIODIC_BOUNDARIES
    double domainLimits[DIM * 2];
    particles->getDomainLimits(domainLimits); // :43
    Domain::Cell boundingBox{domainLimits};
    domain.bounds = boundingBox;
    domain.createGrid(c->h); // :49
#endif
    particles->assignParticlesAndCells(domain); // :53
    lap(0);
#if PERIODIC_BOUNDARIES
    Particles &ghostParticles = *c->ghosts;
    particles->createGhostParticles(domain, ghostParticles, c->h); // :60
#endif
    particles->gridNNS(domain, c->h); // :66
#if PERIODIC_BOUNDARIES
    particles->ghostNNS(domain, ghostParticles, c->h); // :69
#endif
    lap(1);
    particles->compDensity(c->h); // :73
#if PERIODIC_BOUNDARIES
    particles->compDensity(ghostParticles, c->h); // :75
#endif
    particles->compPressure(c->gamma); // :79
    lap(2);
    double timeStep;
    if (dtFixed >= 0.) { /* a zero-length step is what the reference driver does at every dump time (quirk Q7) */
        timeStep = dtFixed;
    } else {
        timeStep = particles->compGlobalTimestep(c->gamma, c->h); // :93
        c->dtCfl = timeStep;
        if (dtMax > 0. && timeStep > dtMax) timeStep = dtMax;
    }
    lap(3);
#if PERIODIC_BOUNDARIES
    particles->updateGhostState(ghostParticles);                 // :109
    particles->compPsijTilde(c->helper, ghostParticles, c->h);   // :110
    particles->gradient(particles->rho, particles->rhoGrad, ghostParticles.rho, ghostParticles);
    particles->gradient(particles->vx, particles->vxGrad, ghostParticles.vx, ghostParticles);
    particles->gradient(particles->vy, particles->vyGrad, ghostParticles.vy, ghostParticles);
#if DIM == 3
    particles->gradient(particles->vz, particles->vzGrad, ghostParticles.vz, ghostParticles);
#endif
    particles->gradient(particles->P, particles->PGrad, ghostParticles.P, ghostParticles);
    particles->updateGhostGradients(ghostParticles); // :122
#else
    particles->compPsijTilde(c->helper, c->h); // :132
    particles->gradient(particles->rho, particles->rhoGrad);
    particles->gradient(particles->vx, particles->vxGrad);
    particles->gradient(particles->vy, particles->vyGrad);
#if DIM == 3
    particles->gradient(particles->vz, particles->vzGrad);
#endif
    particles->gradient(particles->P, particles->PGrad);
#endif
    if (c->capture) {
        size_t s = (size_t)DIM * N, k = 0;
        copyGrad(c->gradPre, s * k++, particles->rhoGrad, N);
        copyGrad(c->gradPre, s * k++, particles->vxGrad, N);
        copyGrad(c->gradPre, s * k++, particles->vyGrad, N);
#if DIM == 3
        copyGrad(c->gradPre, s * k++, particles->vzGrad, N);
#endif
        copyGrad(c->gradPre, s * k++, particles->PGrad, N);
    }
    lap(4);
#if SLOPE_LIMITING
#if PERIODIC_BOUNDARIES
    particles->slopeLimiter(c->h, &ghostParticles);  // :127
    particles->updateGhostGradients(ghostParticles); // :129
#else
    particles->slopeLimiter(c->h); // :142
#endif
#endif
    lap(5);
    particles->compEffectiveFace(); // :148
#if PERIODIC_BOUNDARIES
    particles->compEffectiveFace(ghostParticles); // :150
#endif
    lap(6);
    particles->compRiemannStatesLR(timeStep, c->h, c->gamma); // :153
#if PERIODIC_BOUNDARIES
    particles->compRiemannStatesLR(timeStep, c->h, c->gamma, ghostParticles); // :157
#endif
    lap(7);
    c->dt = timeStep;
    if (stopAfter == 1) return timeStep;
#if PERIODIC_BOUNDARIES
    particles->solveRiemannProblems(c->gamma, ghostParticles); // :204
    lap(8);
    particles->collectFluxes(c->helper, ghostParticles); // :221
#else
    {
        Particles dummyGhosts{0, true}; // :206
        particles->solveRiemannProblems(c->gamma, dummyGhosts);
        lap(8);
        particles->collectFluxes(c->helper, dummyGhosts);
    }
#endif
    lap(9);
    particles->updateStateAndPosition(timeStep, domain); // :224
    lap(10);
    return timeStep;
}

double ref_last_dt_cfl(void *ctx) { return ((RefCtx *)ctx)->dtCfl; }
void ref_phase_seconds(void *ctx, double *out) { std::memcpy(out, ((RefCtx *)ctx)->phaseSec, sizeof(double) * 11); }

/* conservation sums (Particles.cpp:2830-2886): V, M, E, px, py, pz */
void ref_sums(void *ctx, double *out) {
    RefCtx *c = (RefCtx *)ctx;
    out[0] = c->p->sumVolume();
    out[1] = c->p->sumMass();
    out[2] = c->p->sumEnergy();
    out[3] = c->p->sumMomentumX();
    out[4] = c->p->sumMomentumY();
#if DIM == 3
    out[5] = c->p->sumMomentumZ();
#else
    out[5] = 0.;
#endif
}

void ref_grid(void *ctx, int *cells, double *cellSize, double *bounds) {
    RefCtx *c = (RefCtx *)ctx;
    cells[0] = c->domain->cellsX;
    cells[1] = c->domain->cellsY;
    cellSize[0] = c->domain->cellSizeX;
    cellSize[1] = c->domain->cellSizeY;
    bounds[0] = c->domain->bounds.minX;
    bounds[1] = c->domain->bounds.minY;
    bounds[DIM] = c->domain->bounds.maxX;
    bounds[DIM + 1] = c->domain->bounds.maxY;
#if DIM == 3
    cells[2] = c->domain->cellsZ;
    cellSize[2] = c->domain->cellSizeZ;
    bounds[2] = c->domain->bounds.minZ;
    bounds[DIM + 2] = c->domain->bounds.maxZ;
#else
    cells[2] = 1;
    cellSize[2] = 0.;
#endif
}

/*
 * Copy a named array into dst (dst == NULL: just return its length in elements).
 * Doubles unless noted.  Per-slot arrays have MAX_NUM_INTERACTIONS (or ..GHOST..) slots per particle.
 */
long ref_fetch(void *ctx, const char *name, void *dst) {
    RefCtx *c = (RefCtx *)ctx;
    Particles *p = c->p;
    const long N = c->N;
    const std::string n(name);
#define RET(ptr, count, type)                                                \
    do {                                                                     \
        if (dst) std::memcpy(dst, (const void *)(ptr), sizeof(type) * (count)); \
        return (long)(count);                                                \
    } while (0)
    if (n == "x") RET(p->x, N, double);
    if (n == "y") RET(p->y, N, double);
    if (n == "vx") RET(p->vx, N, double);
    if (n == "vy") RET(p->vy, N, double);
#if DIM == 3
    if (n == "z") RET(p->z, N, double);
    if (n == "vz") RET(p->vz, N, double);
    if (n == "vzGrad") RET(p->vzGrad, N * DIM, double);
#endif
    if (n == "m") RET(p->m, N, double);
    if (n == "u") RET(p->u, N, double);
    if (n == "rho") RET(p->rho, N, double);
    if (n == "P") RET(p->P, N, double);
    if (n == "omega") RET(p->omega, N, double);
    if (n == "cell") RET(p->cell, N, int);
    if (n == "noi") RET(p->noi, N, int);
    if (n == "nnl") RET(p->nnl, N * MAX_NUM_INTERACTIONS, int);
    if (n == "rhoGrad") RET(p->rhoGrad, N * DIM, double);
    if (n == "vxGrad") RET(p->vxGrad, N * DIM, double);
    if (n == "vyGrad") RET(p->vyGrad, N * DIM, double);
    if (n == "PGrad") RET(p->PGrad, N * DIM, double);
    if (n == "gradPre") RET(c->gradPre.data(), (long)c->gradPre.size(), double);
    if (n == "psijTilde") RET(p->psijTilde_xi, N * MAX_NUM_INTERACTIONS * DIM, double);
    if (n == "Aij") RET(p->Aij, N * MAX_NUM_INTERACTIONS * DIM, double);
    if (n == "WijL") RET(p->WijL, N * MAX_NUM_INTERACTIONS * (DIM + 2), double);
    if (n == "WijR") RET(p->WijR, N * MAX_NUM_INTERACTIONS * (DIM + 2), double);
    if (n == "Fij") RET(p->Fij, N * MAX_NUM_INTERACTIONS * (DIM + 2), double);
    if (n == "vFrame") RET(p->vFrame, N * MAX_NUM_INTERACTIONS * DIM, double);
    if (n == "mF") RET(p->mF, N, double);
    if (n == "eF") RET(p->eF, N, double);
    if (n == "vF") RET(p->vF, N * DIM, double);
#if PERIODIC_BOUNDARIES
    if (n == "noiGhosts") RET(p->noiGhosts, N, int);
    if (n == "nnlGhosts") RET(p->nnlGhosts, N * MAX_NUM_GHOST_INTERACTIONS, int);
    if (n == "ghostMap") RET(p->ghostMap, N * (DIM + 1), int);
    if (n == "AijGhosts") RET(p->AijGhosts, N * MAX_NUM_GHOST_INTERACTIONS * DIM, double);
    if (n == "FijGhosts") RET(p->FijGhosts, N * MAX_NUM_GHOST_INTERACTIONS * (DIM + 2), double);
    if (n == "WijLGhosts") RET(p->WijLGhosts, N * MAX_NUM_GHOST_INTERACTIONS * (DIM + 2), double);
    if (n == "WijRGhosts") RET(p->WijRGhosts, N * MAX_NUM_GHOST_INTERACTIONS * (DIM + 2), double);
    if (n == "ghost_N") {
        if (dst) *(int *)dst = c->ghosts->N;
        return 1;
    }
    if (n == "ghost_x") RET(c->ghosts->x, c->ghosts->N, double);
    if (n == "ghost_y") RET(c->ghosts->y, c->ghosts->N, double);
    if (n == "ghost_parent") RET(c->ghosts->parent, c->ghosts->N, int);
#endif
#undef RET
    return -1;
}

/* wall-clock of `nsteps` full steps (phases 1-16), adaptive dt; returns seconds per step (median) */
double ref_time_steps(void *ctx, int nsteps, double *perStep) {
    RefCtx *c = (RefCtx *)ctx;
    int cap = c->capture;
    c->capture = 0;
    std::vector<double> ts;
    for (int s = 0; s < nsteps; ++s) {
        auto t0 = std::chrono::steady_clock::now();
        ref_step(ctx, -1., -1., 0);
        auto t1 = std::chrono::steady_clock::now();
        ts.push_back(std::chrono::duration<double>(t1 - t0).count());
        if (perStep) perStep[s] = ts.back();
    }
    c->capture = cap;
    std::vector<double> sorted(ts);
    std::sort(sorted.begin(), sorted.end());
    return sorted[sorted.size() / 2];
}

} // extern "C"
