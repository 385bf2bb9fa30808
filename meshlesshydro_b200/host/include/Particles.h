// GPU-backed Particles: the public surface of the reference's class
// (/root/reference/demonstrator/include/Particles.h:27-198, MFV part) over the C ABI in include/mlh_gpu.h.
//
// The reference keeps everything in host arrays and every method is a serial loop.  Here the particle state lives
// in HBM (SoA FP64, see csrc/mlh_internal.cuh); the public host arrays are page-locked MIRRORS that are uploaded
// before the first phase after they were edited (markHostStateChanged) and refreshed by syncHost() / dump2file().
// Each method maps to the device phase that contains the reference's loop; a method whose work is already part of
// an earlier fused phase is a no-op.  Calling a method out of MeshlessScheme::run()'s order runs the missing
// phases first, so the reference's driver works unchanged:
//
//   assignParticlesAndCells                         -> mlh_build_grid      (K0 bbox, K1 sort)
//   createGhostParticles, gridNNS, ghostNNS         -> mlh_neighbours      (K2 + face list; images are implicit)
//   compDensity(+ghost overload), compPressure      -> mlh_density_matrix  (K3)
//   compGlobalTimestep, updateGhostState, compPsijTilde, gradient x5,
//   updateGhostGradients, slopeLimiter              -> mlh_gradients_limit (K3b) [+ mlh_timestep]
//   compEffectiveFace, compRiemannStatesLR, solveRiemannProblems, collectFluxes,
//   updateStateAndPosition                          -> mlh_flux_update     (K4a, K4b, K4c/K5), issued by
//                                                      updateStateAndPosition (the first call that knows dt AND
//                                                      follows the snapshot point of the loop)
#ifndef DEMONSTRATOR_PARTICLES_H
#define DEMONSTRATOR_PARTICLES_H

#include <string>
#include <vector>

#include "parameter.h"
#include "Domain.h"
#include "Helper.h"
#include "Logger.h"

struct mlh_ctx;

class Particles {
public:
    Particles(int numParticles, bool ghosts = false);
    ~Particles();
    Particles(const Particles &) = delete;
    Particles &operator=(const Particles &) = delete;

    int N;
    int *matId;
    int *cell; // search-grid cell of each particle (refreshed by syncHost)
    double *m, *u, *x, *y, *vx, *vy, *rho, *P;
#if DIM == 3
    double *z, *vz;
#endif
    double (*rhoGrad)[DIM], (*vxGrad)[DIM], (*vyGrad)[DIM], (*vzGrad)[DIM], (*PGrad)[DIM];
    int *noi; // regular neighbours per particle (written to snapshots; private in the reference)

    void assignParticlesAndCells(Domain &domain);
    void gridNNS(Domain &domain, const double &kernelSize);
    void compDensity(const double &kernelSize); // also computes omega
    void compPsijTilde(Helper &helper, const double &kernelSize);
    void gradient(double *f, double (*grad)[DIM]);
    void slopeLimiter(const double &kernelSize, Particles *ghostParticles = nullptr);
    void compPressure(const double &gamma);
    void compEffectiveFace();
    double compGlobalTimestep(const double &gamma, const double &kernelSize);
    void compRiemannStatesLR(const double &dt, const double &kernelSize, const double &gamma);
    void solveRiemannProblems(const double &gamma, const Particles &ghostParticles);
    void collectFluxes(Helper &helper, const Particles &ghostParticles);
    void updateStateAndPosition(const double &dt, const Domain &domain);
    double pairwiseLimiter(double phi_0, double phi_i, double phi_j, double xijxi_abs, double xjxi_abs);

    // periodic-boundary overloads (declared in every build; the reference guards them with PERIODIC_BOUNDARIES)
    void createGhostParticles(Domain &domain, Particles &ghostParticles, const double &kernelSize);
    void ghostNNS(Domain &domain, const Particles &ghostParticles, const double &kernelSize);
    void compDensity(const Particles &ghostParticles, const double &kernelSize);
    void compPsijTilde(Helper &helper, const Particles &ghostParticles, const double &kernelSize);
    void gradient(double *f, double (*grad)[DIM], double *fGhost, const Particles &ghostParticles);
    void compEffectiveFace(const Particles &ghostParticles);
    void compRiemannStatesLR(const double &dt, const double &kernelSize, const double &gamma, const Particles &ghostParticles);
    void updateGhostState(Particles &ghostParticles);
    void updateGhostGradients(Particles &ghostParticles);
    void getDomainLimits(double *domainLimits);

    double sumVolume();
    double sumMass();
    double sumEnergy();
    double sumMomentumX();
    double sumMomentumY();
#if DIM == 3
    double sumMomentumZ();
#endif
    void checkFluxSymmetry(Particles *ghostParticles = nullptr);
    void dump2file(std::string filename, double simTime);

    // ---- additions of the GPU build ----
    /// run-time part of the configuration the device needs before the first phase (MeshlessScheme's ctor calls it)
    void configureDevice(const double &kernelSize, const double &gamma, const double *periodicBoxLimits);
    void markHostStateChanged() { hostDirty = true; }
    /// device -> host mirrors: state always; rho, P, rhoGrad, noi, cell when the step is past the gradient phase
    void syncHost();
    mlh_ctx *context() { return gpu; }
    /// original indices of the particles slab `rank` of `nranks` owns (whole cell layers of the slowest axis; what the
    /// multi-rank launcher uploads per rank) -- host arithmetic only
    std::vector<int> slabParticles(int rank, int nranks);
    long kernelLaunches() const;

private:
    enum Phase { PH_STATE = 0, PH_GRID = 1, PH_NEIGHBOURS = 2, PH_DENSITY = 3, PH_GRADIENTS = 4 };
    void ensure(int phase); // run the device phases up to `phase`
    void selectShard();     // multi-rank runs: shardIds = slabParticles(this rank)
    void check(int rc, const char *what);
    void checkFlags();
    void sums();
    bool ghostHolder;
    mlh_ctx *gpu{nullptr};
    bool hostDirty{true};   // host mirrors edited since the last upload
    bool hostStale{false};  // device state newer than the host mirrors
    int phase{PH_STATE};
    double sumCache[6];
    bool sumsValid{false};
    int listCapacity{0};          // neighbour-list capacity the device context was created with (for the exit(1) message)
    bool flagsCheckedOnce{false}; // device flags read after the first neighbour search (reference: exit inside gridNNS)
    double hCfg{0.}, gammaCfg{0.};
    double boxCfg[2 * DIM];
    bool configured{false};
    std::vector<int> shardIds;
};

#endif // DEMONSTRATOR_PARTICLES_H
