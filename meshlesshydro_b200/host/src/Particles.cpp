// GPU-backed Particles -- see Particles.h.  No physics here: every loop of the reference's Particles.cpp runs in the
// kernels of libmlh_gpu.so; this file is the binding, the host mirrors and the reference's error behaviour.
#include "../include/Particles.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../../include/mlh_gpu.h"
#include "../include/H5Lite.h"
#include "../include/MultiGpu.h"

namespace {
bool g_pinned_ok = true;
template <typename T> T *hostArray(size_t n) {
    // multi-rank runs (MultiGpu.h): the public arrays are ONE copy shared by all rank processes
    if (mgpu::planned() > 1) return (T *)mgpu::sharedAlloc((n ? n : 1) * sizeof(T));
    if (g_pinned_ok) {
        void *p = nullptr;
        if (mlh_host_alloc((n ? n : 1) * sizeof(T), &p) == MLH_OK) {
            std::memset(p, 0, (n ? n : 1) * sizeof(T));
            return (T *)p;
        }
        g_pinned_ok = false;
    }
    T *p = (T *)std::calloc(n ? n : 1, sizeof(T));
    if (!p) throw std::bad_alloc();
    return p;
}
template <typename T> void hostFree(T *p) {
    if (!p || mgpu::planned() > 1) return; // shared mappings go with the process
    if (g_pinned_ok) mlh_host_free((void *)p);
    else std::free((void *)p);
}
int envInt(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}
[[noreturn]] void die(int code) {
    if (mgpu::active()) mgpu::fail(); // the other ranks may be blocked in a barrier or an NCCL call
    exit(code);
}
} // namespace

Particles::Particles(int numParticles, bool ghosts) : N{numParticles}, ghostHolder{ghosts} {
    // The ghost holder of the reference (Particles.cpp:71-107) stores copies of the parents' state for the periodic
    // images.  The device search generates the images on the fly (csrc/k2_neighbours.cu), so the holder is empty.
    const size_t n = ghosts ? 0 : (size_t)numParticles;
    if (ghosts) N = 0;
    matId = hostArray<int>(n);
    cell = hostArray<int>(n);
    noi = hostArray<int>(n);
    m = hostArray<double>(n);
    u = hostArray<double>(n);
    x = hostArray<double>(n);
    y = hostArray<double>(n);
    vx = hostArray<double>(n);
    vy = hostArray<double>(n);
    rho = hostArray<double>(n);
    P = hostArray<double>(n);
#if DIM == 3
    z = hostArray<double>(n);
    vz = hostArray<double>(n);
#endif
    rhoGrad = (double(*)[DIM])hostArray<double>(n * DIM);
    vxGrad = (double(*)[DIM])hostArray<double>(n * DIM);
    vyGrad = (double(*)[DIM])hostArray<double>(n * DIM);
    vzGrad = (double(*)[DIM])hostArray<double>(n * DIM);
    PGrad = (double(*)[DIM])hostArray<double>(n * DIM);
    for (int k = 0; k < 6; ++k) sumCache[k] = 0.;
    for (int k = 0; k < 2 * DIM; ++k) boxCfg[k] = 0.;
}

Particles::~Particles() {
    if (gpu) mlh_destroy(gpu);
    hostFree(matId); hostFree(cell); hostFree(noi); hostFree(m); hostFree(u); hostFree(x); hostFree(y);
    hostFree(vx); hostFree(vy); hostFree(rho); hostFree(P);
#if DIM == 3
    hostFree(z); hostFree(vz);
#endif
    hostFree((double *)rhoGrad); hostFree((double *)vxGrad); hostFree((double *)vyGrad); hostFree((double *)vzGrad);
    hostFree((double *)PGrad);
}

void Particles::check(int rc, const char *what) {
    if (rc == MLH_OK) return;
    Logger(ERROR) << what << " failed (" << rc << "): " << mlh_last_error(gpu) << " - Aborting.";
    die(rc == MLH_E_NO_DEVICE ? 20 : 21);
}

// device error flags -> the reference's messages and exit codes
void Particles::checkFlags() {
    const unsigned f = mlh_error_flags(gpu);
    if (f & MLH_F_MAX_INTERACTIONS) {
        Logger(ERROR) << "MAX_NUM_INTERACTIONS exceeded for at least one particle (device list capacity " << listCapacity
                      << ", set MLH_MAX_INTERACTIONS to change it) - Aborting."; // Particles.cpp:348-352 / :2249-2253
        die(1);
    }
    if (f & (MLH_F_MIGRATION | MLH_F_HALO_OVERFLOW)) {
        // a particle that crossed more than one cell layer of a slab in one step is no longer owned by any rank, a halo
        // buffer that overflowed dropped particles: mass and energy would change silently
        Logger(ERROR) << ((f & MLH_F_MIGRATION) ? "A particle moved across more than one cell layer of its slab in one step"
                                                : "Halo / migration buffer overflow")
                      << " (multi-GPU run; raise the capacity or lower the time step) - Aborting.";
        die(23);
    }
    if (f & MLH_F_OUT_OF_GRID) {
        Logger(ERROR) << "Particle outside of the search grid. - Aborting.";
        die(2);
    }
#if DEBUG_LVL
    if (f & MLH_F_NEG_GHOST_PRESSURE) {
        Logger(ERROR) << "Negative pressure encountered @ghost face. Very bad :( !!"; // Particles.cpp:1873-1881
        die(6);
    }
#endif
    if (f & MLH_F_VACUUM) Logger(WARN) << "  > Vacuum state sampled. This is not expected."; // Riemann.cpp:113,126
}

void Particles::configureDevice(const double &kernelSize, const double &gamma, const double *periodicBoxLimits) {
    hCfg = kernelSize;
    gammaCfg = gamma;
    if (periodicBoxLimits)
        for (int k = 0; k < 2 * DIM; ++k) boxCfg[k] = periodicBoxLimits[k];
    configured = true;
}

// The particles a rank owns: whole cell layers of the search grid along the slowest-varying axis (y in 2D, z in 3D:
// cell id = iX + iY cellsX + iZ cellsX cellsY, Particles.cpp:298-302), the same split the device uses (mlh_slab_range).
// Grid as Domain::createGrid builds it (Domain.cpp:9-54) from the periodic box or from getDomainLimits.
std::vector<int> Particles::slabParticles(int rank, int nranks) {
    double lim[2 * DIM];
#if PERIODIC_BOUNDARIES
    for (int k = 0; k < 2 * DIM; ++k) lim[k] = boxCfg[k];
#else
    getDomainLimits(lim); // host loop while nothing is on the device yet
#endif
    const int k = DIM - 1;
    const double *coord = y;
#if DIM == 3
    coord = z;
#endif
    const double lo = lim[k], hi = lim[DIM + k];
    const int layers = (int)std::floor((hi - lo) / hCfg);
    if (layers < 2 * nranks) {
        Logger(ERROR) << "slab decomposition needs >= 2 cell layers per rank (" << layers << " layers, " << nranks
                      << " ranks) - Aborting.";
        die(21);
    }
    const double size = (hi - lo) / layers;
    int first = 0, last = 0;
    if (mlh_slab_range(layers, nranks, rank, &first, &last) != MLH_OK) {
        Logger(ERROR) << "mlh_slab_range failed - Aborting.";
        die(21);
    }
    std::vector<int> mine;
    for (int i = 0; i < N; ++i) {
        int layer = (int)std::floor((coord[i] - lo) / size);
        if (layer == layers) layer -= 1; // Particles.cpp:283-295
        if (layer >= first && layer < last) mine.push_back(i);
    }
    return mine;
}

void Particles::selectShard() { shardIds = slabParticles(mgpu::rank(), mgpu::nranks()); }

void Particles::ensure(int target) {
    if (ghostHolder) return;
    if (!gpu) {
        if (!configured) {
            Logger(ERROR) << "Particles: configureDevice(kernelSize, gamma, box) must precede the first phase. - Aborting.";
            die(21);
        }
        mlh_config cfg;
        mlh_default_config(&cfg);
        cfg.dim = DIM;
        cfg.periodic = PERIODIC_BOUNDARIES;
        // list capacity: the reference reserves MAX_NUM_INTERACTIONS (+ MAX_NUM_GHOST_INTERACTIONS) slots per particle
        // (~100 kB per particle); the device lists hold 16 B per slot and default to the same capacity (MLH_MAX_INTERACTIONS
        // lowers it for large runs)
        int cap = MAX_NUM_INTERACTIONS + (PERIODIC_BOUNDARIES ? MAX_NUM_GHOST_INTERACTIONS : 0);
        if (cap > 1023) cap = 1023; // limit of the device's slot -> face map
        cfg.max_interactions = envInt("MLH_MAX_INTERACTIONS", cap);
        listCapacity = cfg.max_interactions;
        cfg.slope_limiting = SLOPE_LIMITING;
        cfg.pairwise_limiter = PAIRWISE_LIMITER;
        cfg.meshless_finite_mass = MESHLESS_FINITE_MASS;
        cfg.move_particles = MOVE_PARTICLES;
        cfg.first_order_quad_point = FIRST_ORDER_QUAD_POINT;
        cfg.abs_mode = envInt("MLH_ABS_MODE", MLH_ABS_FABS);
        cfg.q13_mode = envInt("MLH_Q13_MODE", MLH_Q13_ZERO_Z);
        cfg.q3_mode = envInt("MLH_Q3_MODE", MLH_Q3_REFERENCE);
        cfg.symmetric_seam = envInt("MLH_SYMMETRIC_SEAM", 0);
        cfg.cfl = CFL;
        cfg.beta = BETA;
        cfg.psi1 = PSI_1;
        cfg.psi2 = PSI_2;
        cfg.kernel_size = hCfg;
        cfg.gamma = gammaCfg;
        for (int k = 0; k < 2 * DIM; ++k) cfg.box[k] = boxCfg[k];
        cfg.device = envInt("MLH_DEVICE", 0);
        if (mgpu::active()) { // one rank per GPU, slab decomposition (MultiGpu.h)
            cfg.rank = mgpu::rank();
            cfg.nranks = mgpu::nranks();
            cfg.device = envInt("MLH_DEVICE", 0) + mgpu::rank();
            selectShard();
            const long share = (long)((double)N / mgpu::nranks() * 1.6) + 8192, mine = (long)((double)shardIds.size() * 1.6) + 8192;
            cfg.capacity = share > mine ? share : mine;
        }
        int rc = mlh_create(&cfg, &gpu);
        if (rc != MLH_OK) {
            Logger(ERROR) << "mlh_create failed (" << rc << "): " << mlh_last_error(nullptr) << " - Aborting.";
            die(rc == MLH_E_NO_DEVICE ? 20 : 21);
        }
        if (mgpu::active()) {
            if (mgpu::rank() == 0) check(mlh_comm_unique_id(mgpu::ncclId()), "mlh_comm_unique_id");
            mgpu::barrier();
            check(mlh_comm_init(gpu, mgpu::ncclId()), "mlh_comm_init");
        }
    }
    if (hostDirty && mgpu::active()) {
        // upload the particles of this rank's slab with their original indices
        if (shardIds.empty() && N > 0) selectShard();
        const size_t n = shardIds.size();
        std::vector<double> buf[8];
        const double *src[8] = {x, y, nullptr, vx, vy, nullptr, m, u};
#if DIM == 3
        src[2] = z;
        src[5] = vz;
#endif
        for (int f = 0; f < 8; ++f) {
            if (!src[f]) continue;
            buf[f].resize(n);
            for (size_t k = 0; k < n; ++k) buf[f][k] = src[f][shardIds[k]];
        }
        check(mlh_upload(gpu, (long)n, buf[0].data(), buf[1].data(), src[2] ? buf[2].data() : nullptr, buf[3].data(), buf[4].data(),
                         src[5] ? buf[5].data() : nullptr, buf[6].data(), buf[7].data(), shardIds.data()),
              "mlh_upload");
        shardIds.clear(); // ownership moves with the particles from now on
        shardIds.shrink_to_fit();
        hostDirty = false;
        hostStale = false;
        phase = PH_STATE;
        sumsValid = false;
    }
    if (hostDirty) {
        const double *zz = nullptr, *vzz = nullptr;
#if DIM == 3
        zz = z;
        vzz = vz;
#endif
        check(mlh_upload(gpu, N, x, y, zz, vx, vy, vzz, m, u, nullptr), "mlh_upload");
        hostDirty = false;
        hostStale = false;
        phase = PH_STATE;
        sumsValid = false;
    }
    if (phase < PH_GRID && target >= PH_GRID) {
        check(mlh_build_grid(gpu), "mlh_build_grid");
        phase = PH_GRID;
    }
    if (phase < PH_NEIGHBOURS && target >= PH_NEIGHBOURS) {
        check(mlh_neighbours(gpu), "mlh_neighbours");
        phase = PH_NEIGHBOURS;
        // the reference exits inside gridNNS / ghostNNS (Particles.cpp:348-352, :2249-2253), i.e. before anything is
        // computed or dumped from a truncated list.  Reading the flag word synchronises, so it is done every step only
        // when MLH_CHECK_EVERY_STEP is set, otherwise on the first step and then after each update.
        if (!flagsCheckedOnce || envInt("MLH_CHECK_EVERY_STEP", 0)) {
            checkFlags();
            flagsCheckedOnce = true;
        }
    }
    if (phase < PH_DENSITY && target >= PH_DENSITY) {
        check(mlh_density_matrix(gpu), "mlh_density_matrix");
        phase = PH_DENSITY;
        sumsValid = false;
    }
    if (phase < PH_GRADIENTS && target >= PH_GRADIENTS) {
        check(mlh_gradients_limit(gpu), "mlh_gradients_limit");
        phase = PH_GRADIENTS;
    }
}

void Particles::assignParticlesAndCells(Domain &domain) {
    ensure(PH_GRID);
    // the Domain object keeps describing the grid the device uses
    int cells[3];
    double size[3], bounds[6];
    mlh_grid_info(gpu, cells, size, bounds);
    domain.cellsX = cells[0];
    domain.cellsY = cells[1];
    domain.cellSizeX = size[0];
    domain.cellSizeY = size[1];
    domain.numGridCells = cells[0] * cells[1];
#if DIM == 3
    domain.cellsZ = cells[2];
    domain.cellSizeZ = size[2];
    domain.numGridCells *= cells[2];
#endif
}

void Particles::gridNNS(Domain &, const double &) { ensure(PH_NEIGHBOURS); }
void Particles::createGhostParticles(Domain &, Particles &ghostParticles, const double &) {
    ghostParticles.N = 0; // periodic images are generated inside the neighbour search
}
void Particles::ghostNNS(Domain &, const Particles &, const double &) { ensure(PH_NEIGHBOURS); }
void Particles::compDensity(const double &) { ensure(PH_DENSITY); }
void Particles::compDensity(const Particles &, const double &) { ensure(PH_DENSITY); }
void Particles::compPressure(const double &) { ensure(PH_DENSITY); }

double Particles::compGlobalTimestep(const double &, const double &) {
    ensure(PH_GRADIENTS);
    double dt = 0.;
    check(mlh_timestep(gpu, &dt), "mlh_timestep");
    return dt;
}

void Particles::updateGhostState(Particles &) {}
void Particles::updateGhostGradients(Particles &) {}
void Particles::compPsijTilde(Helper &, const double &) { ensure(PH_GRADIENTS); }
void Particles::compPsijTilde(Helper &, const Particles &, const double &) { ensure(PH_GRADIENTS); }

void Particles::gradient(double *f, double (*grad)[DIM]) {
    // the device computes the gradients of rho, vx, vy, (vz), P in one pass; other fields are not part of the path
    const bool known = (f == rho && grad == rhoGrad) || (f == vx && grad == vxGrad) || (f == vy && grad == vyGrad) ||
                       (f == P && grad == PGrad)
#if DIM == 3
                       || (f == vz && grad == vzGrad)
#endif
        ;
    if (!known) {
        Logger(ERROR) << "Particles::gradient: only the gradients of rho, vx, vy, vz, P are computed on the device. - Aborting.";
        die(21);
    }
    ensure(PH_GRADIENTS);
}
void Particles::gradient(double *f, double (*grad)[DIM], double *, const Particles &) { gradient(f, grad); }
void Particles::slopeLimiter(const double &, Particles *) { ensure(PH_GRADIENTS); }
void Particles::compEffectiveFace() { ensure(PH_GRADIENTS); }
void Particles::compEffectiveFace(const Particles &) { ensure(PH_GRADIENTS); }
void Particles::compRiemannStatesLR(const double &, const double &, const double &) { ensure(PH_GRADIENTS); }
void Particles::compRiemannStatesLR(const double &, const double &, const double &, const Particles &) { ensure(PH_GRADIENTS); }
void Particles::solveRiemannProblems(const double &, const Particles &) { ensure(PH_GRADIENTS); }
void Particles::collectFluxes(Helper &, const Particles &) { ensure(PH_GRADIENTS); }
// Particles::checkFluxSymmetry (reference Particles.cpp:2888-2976, called when DEBUG_LVL > 1): every pair of the lists is
// ONE stored flux that both endpoints add with opposite signs, so Fij + Fji == 0 exactly; what can still go wrong is the
// slot -> face map, which the device verifies (mlh_debug_fetch "flux_symmetry").  One-sided periodic pairs (quirk Q9) are
// where the reference prints "fluxes are NOT symmetric".
void Particles::checkFluxSymmetry(Particles *) {
    if (ghostHolder || !gpu || phase < PH_NEIGHBOURS) return;
    // the reference pays 11 % of its step for this check whenever DEBUG_LVL is set; here it costs a kernel and a
    // read-back, spent only in verbose runs (-v) or when MLH_CHECK_FLUX_SYMMETRY is set
    if (!(LOGCFG.level <= DEBUG || envInt("MLH_CHECK_FLUX_SYMMETRY", 0))) return;
    int r[4] = {0, 0, 0, 0};
    if (mlh_debug_fetch(gpu, "flux_symmetry", r, 4) != 4) return;
    if (r[1] > 0) Logger(WARN) << "  > Fluxes are NOT symmetric for " << r[1] << " list slots (slot -> face map inconsistent)";
#if PERIODIC_BOUNDARIES
    if (r[3] > 0 && !mgpu::active())
        Logger(WARN) << "  > Ghosts: " << r[3] << " one-sided periodic pairs (the reference's fluxes are NOT symmetric there)";
#endif
}

void Particles::updateStateAndPosition(const double &dt, const Domain &) {
    ensure(PH_GRADIENTS);
    check(mlh_flux_update(gpu, dt), "mlh_flux_update");
    phase = PH_STATE;
    hostStale = true;
    sumsValid = false;
    checkFlags();
}

// Particles::pairwiseLimiter, reference Particles.cpp:1735-1785, host version for callers of the class (the device
// copy is pairwise_limiter in csrc/k4_flux.cu); fabs semantics of quirk Q1
double Particles::pairwiseLimiter(double phi0, double phi_i, double phi_j, double xijxi_abs, double xjxi_abs) {
    double phi_ = phi_i;
    const double phi_ij = phi_i + xijxi_abs / xjxi_abs * (phi_j - phi_i);
    const double phiMin = phi_i < phi_j ? phi_i : phi_j, phiMax = phi_i < phi_j ? phi_j : phi_i;
    const double delta1 = PSI_1 * std::fabs(phi_i - phi_j), delta2 = PSI_2 * std::fabs(phi_i - phi_j);
    const bool sameMax = (phiMax + delta1 >= 0. && phiMax >= 0.) || (phiMax + delta1 < 0. && phiMax < 0.);
    const bool sameMin = (phiMin - delta1 >= 0. && phiMin >= 0.) || (phiMin - delta1 < 0. && phiMin < 0.);
    const double phiPlus = sameMax ? phiMax + delta1 : phiMax / (1. + delta1 / std::fabs(phiMax));
    const double phiMinus = sameMin ? phiMin - delta1 : phiMin / (1. + delta1 / std::fabs(phiMin));
    if (phi_i < phi_j) {
        const double t = (phi_ij + delta2 < phi0) ? phi_ij + delta2 : phi0;
        phi_ = phiMinus > t ? phiMinus : t;
    } else if (phi_i > phi_j) {
        const double t = (phi_ij - delta2 > phi0) ? phi_ij - delta2 : phi0;
        phi_ = phiPlus < t ? phiPlus : t;
    }
    return phi_;
}

// Particles::getDomainLimits (reference Particles.cpp:228-267) incl. its semantics (quirks Q2, Q8): the running
// maximum starts at the smallest positive normal and a particle that lowers the minimum is not tested against it.
void Particles::getDomainLimits(double *domainLimits) {
    if (gpu && !hostDirty) {
        // the state lives on the device: the bounding box is phase 0 of the step (K0), the grid follows from it
        ensure(PH_GRID);
        int cells[3];
        double size[3];
        mlh_grid_info(gpu, cells, size, domainLimits);
        return;
    }
    const double *c[3] = {x, y,
#if DIM == 3
                          z
#else
                          nullptr
#endif
    };
    for (int k = 0; k < DIM; ++k) {
        double lo = std::numeric_limits<double>::max(), hi = std::numeric_limits<double>::min();
        for (int i = 0; i < N; ++i) {
            if (c[k][i] < lo) lo = c[k][i];
            else if (c[k][i] > hi) hi = c[k][i];
        }
        domainLimits[k] = lo;
        domainLimits[DIM + k] = hi;
    }
}

void Particles::sums() {
    if (sumsValid) return;
    ensure(PH_DENSITY); // sumVolume needs omega
    check(mlh_sums(gpu, sumCache), "mlh_sums");
    sumsValid = true;
}
double Particles::sumVolume() { sums(); return sumCache[0]; }
double Particles::sumMass() { sums(); return sumCache[1]; }
double Particles::sumEnergy() { sums(); return sumCache[2]; }
double Particles::sumMomentumX() { sums(); return sumCache[3]; }
double Particles::sumMomentumY() { sums(); return sumCache[4]; }
#if DIM == 3
double Particles::sumMomentumZ() { sums(); return sumCache[5]; }
#endif

void Particles::syncHost() {
    if (ghostHolder || !gpu || hostDirty) return;
    if (mgpu::active()) {
        // every rank scatters its owned particles (device order + original ids) into the shared host arrays
        const long n = mlh_num_particles(gpu);
        std::vector<double> b[8];
        for (auto &v : b) v.resize((size_t)(n > 0 ? n : 1));
        std::vector<int> ids((size_t)(n > 0 ? n : 1));
        const bool three = DIM == 3;
        check(mlh_download_state(gpu, b[0].data(), b[1].data(), three ? b[2].data() : nullptr, b[3].data(), b[4].data(),
                                 three ? b[5].data() : nullptr, b[6].data(), b[7].data(), ids.data()),
              "mlh_download_state");
        for (long q = 0; q < n; ++q) {
            const int i = ids[(size_t)q];
            x[i] = b[0][q]; y[i] = b[1][q]; vx[i] = b[3][q]; vy[i] = b[4][q]; m[i] = b[6][q]; u[i] = b[7][q];
#if DIM == 3
            z[i] = b[2][q]; vz[i] = b[5][q];
#endif
        }
        if (phase >= PH_GRADIENTS) {
            std::vector<double> g((size_t)(n > 0 ? n : 1) * DIM);
            std::vector<int> cnt((size_t)(n > 0 ? n : 1));
            check(mlh_download_diag(gpu, b[0].data(), b[1].data(), g.data(), cnt.data()), "mlh_download_diag");
            for (long q = 0; q < n; ++q) {
                const int i = ids[(size_t)q];
                rho[i] = b[0][q];
                P[i] = b[1][q];
                noi[i] = cnt[(size_t)q];
                for (int a = 0; a < DIM; ++a) rhoGrad[i][a] = g[(size_t)q * DIM + a];
            }
        }
        mgpu::barrier(); // all slabs are in
        hostStale = false;
        return;
    }
    double *zz = nullptr, *vzz = nullptr;
#if DIM == 3
    zz = z;
    vzz = vz;
#endif
    check(mlh_download_state(gpu, x, y, zz, vx, vy, vzz, m, u, nullptr), "mlh_download_state");
    if (phase >= PH_GRADIENTS) {
        check(mlh_download_diag(gpu, rho, P, &rhoGrad[0][0], noi), "mlh_download_diag");
        mlh_debug_fetch(gpu, "cell", cell, N);
        mlh_debug_fetch(gpu, "vxGrad", &vxGrad[0][0], (long)N * DIM);
        mlh_debug_fetch(gpu, "vyGrad", &vyGrad[0][0], (long)N * DIM);
#if DIM == 3
        mlh_debug_fetch(gpu, "vzGrad", &vzGrad[0][0], (long)N * DIM);
#endif
        mlh_debug_fetch(gpu, "PGrad", &PGrad[0][0], (long)N * DIM);
    }
    hostStale = false;
}

long Particles::kernelLaunches() const { return gpu ? mlh_launch_count(gpu) : 0; }

// snapshot with the dataset names, shapes and types of the reference (Particles.cpp:2978-3076):
// /time /totalMass /energy /xMomentum /yMomentum (/zMomentum) f64[1]; /rho /m /u /P f64[N]; /x /v /rhoGrad f64[N][DIM];
// /noi i32[N]
void Particles::dump2file(std::string filename, double simTime) {
    ensure(PH_GRADIENTS);
    syncHost();
    const double t = simTime, mass = sumMass(), energy = sumEnergy(), px = sumMomentumX(), py = sumMomentumY();
#if DIM == 3
    const double pz = sumMomentumZ();
#endif
    if (mgpu::active() && mgpu::rank() != 0) { // the sums above are collective; the file is rank 0's job
        mgpu::barrier();
        return;
    }
    std::vector<double> pos((size_t)N * DIM), vel((size_t)N * DIM);
    for (int i = 0; i < N; ++i) {
        pos[(size_t)i * DIM] = x[i];
        pos[(size_t)i * DIM + 1] = y[i];
        vel[(size_t)i * DIM] = vx[i];
        vel[(size_t)i * DIM + 1] = vy[i];
#if DIM == 3
        pos[(size_t)i * DIM + 2] = z[i];
        vel[(size_t)i * DIM + 2] = vz[i];
#endif
    }
    try {
        H5Lite::Writer w(filename);
        const std::vector<uint64_t> one{1}, n1{(uint64_t)N}, n2{(uint64_t)N, (uint64_t)DIM};
        w.write("/time", one, &t);
        w.write("/totalMass", one, &mass);
        w.write("/energy", one, &energy);
        w.write("/xMomentum", one, &px);
        w.write("/yMomentum", one, &py);
#if DIM == 3
        w.write("/zMomentum", one, &pz);
#endif
        w.write("/rho", n1, rho);
        w.write("/m", n1, m);
        w.write("/u", n1, u);
        w.write("/x", n2, pos.data());
        w.write("/v", n2, vel.data());
        w.write("/rhoGrad", n2, &rhoGrad[0][0]);
        w.write("/P", n1, P);
        w.write("/noi", n1, (const int32_t *)noi);
        w.close();
    } catch (const std::exception &e) {
        Logger(ERROR) << "dump2file(" << filename << "): " << e.what();
        if (mgpu::active()) die(21);
        throw;
    }
    mgpu::barrier(); // the other ranks wait until the shared arrays have been written out
}
