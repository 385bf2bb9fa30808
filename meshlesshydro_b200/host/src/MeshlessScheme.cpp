// MeshlessScheme::run -- the time loop of /root/reference/demonstrator/src/MeshlessScheme.cpp:21-254: same phase order,
// same log lines, same time-step policy and dump schedule (quirk Q7: after file 000000 the first dump target is
// dumpTimes[2]; the read past the end of dumpTimes is guarded).  The body of the reference's loop is split into one
// method per stage (MeshlessScheme.h); every `particles->` call keeps the reference's method name, what it launches on
// the device is listed in Particles.h.
#include "../include/MeshlessScheme.h"

#include <sstream>

MeshlessScheme::MeshlessScheme(Configuration config_, Particles *particles_, Domain::Cell bounds)
    : config{config_}, particles{particles_}, ghostParticles(DIM * particles_->N, true), domain(bounds),
      timeStep{config_.timeStep} {
    Logger(INFO) << "    > Creating grid ... ";
    domain.createGrid(config.kernelSize);
    Logger(INFO) << "    > ... got " << domain.numGridCells << " cells";
    particles->configureDevice(config.kernelSize, config.gamma, config.periodicBoxLimits);
}

MeshlessScheme::~MeshlessScheme() {}

// ---- stage 1 (non-periodic runs): the search grid follows the particles ----
void MeshlessScheme::rebuildGrid() {
#if !PERIODIC_BOUNDARIES
    Logger(INFO) << "    > Computing domain limits ...";
    double limits[DIM * 2];
    particles->getDomainLimits(limits);
    domain.bounds = Domain::Cell{limits};
    domain.printout();
    Logger(DEBUG) << "      > ... creating grid ...";
    domain.createGrid(config.kernelSize);
    Logger(INFO) << "    > ... done.";
#endif
}

// ---- stage 2: cells, periodic images, neighbour lists ----
void MeshlessScheme::searchNeighbours() {
    Logger(INFO) << "    > Assigning particles ...";
    particles->assignParticlesAndCells(domain);
    Logger(INFO) << "    > ... done.";
#if PERIODIC_BOUNDARIES
    Logger(INFO) << "    > Creating ghost particles ...";
    particles->createGhostParticles(domain, ghostParticles, config.kernelSize);
    Logger(INFO) << "    > ... done.";
#endif
    Logger(INFO) << "    > Nearest neighbor search";
    particles->gridNNS(domain, config.kernelSize);
#if PERIODIC_BOUNDARIES
    Logger(DEBUG) << "      > Ghosts NNS";
    particles->ghostNNS(domain, ghostParticles, config.kernelSize);
#endif
}

// ---- stage 3: omega, rho, P (+ the conservation sums when they are going to be printed) ----
void MeshlessScheme::densityAndPressure() {
    Logger(INFO) << "    > Computing density";
    particles->compDensity(config.kernelSize);
#if PERIODIC_BOUNDARIES
    particles->compDensity(ghostParticles, config.kernelSize);
#endif
    Logger(INFO) << "    > Computing pressure";
    particles->compPressure(config.gamma);
    if (LOGCFG.level > DEBUG) return; // the sums cost a device reduction + read-back: only when they are printed
    Logger(DEBUG) << "      SANITY CHECK > V_tot = " << particles->sumVolume();
    Logger(DEBUG) << "      SANITY CHECK > M_tot = " << particles->sumMass();
    Logger(DEBUG) << "      SANITY CHECK > E_tot = " << particles->sumEnergy();
    Logger(DEBUG) << "      SANITY CHECK > px_tot = " << particles->sumMomentumX();
    Logger(DEBUG) << "      SANITY CHECK > py_tot = " << particles->sumMomentumY();
#if DIM == 3
    Logger(DEBUG) << "      SANITY CHECK > pz_tot = " << particles->sumMomentumZ();
#endif
}

// ---- stage 4: dt (CFL, clipped to the next dump time) ----
void MeshlessScheme::chooseTimeStep(double t) {
#if ADAPTIVE_TIMESTEP
    Logger(INFO) << "    > Selecting global timestep ... ";
    timeStep = particles->compGlobalTimestep(config.gamma, config.kernelSize);
    if (dumpStep >= numDumpTimes) {
        Logger(ERROR) << "Simulation did not abort after reaching timeEnd. Exiting.";
        exit(9);
    }
    // (the reference reads dumpTimes[dumpStep+1] unguarded, one past the end for the last target)
    if (dumpStep + 1 < numDumpTimes && t + timeStep >= dumpTimes[dumpStep + 1]) {
        dumpNext = true;
        timeStep = dumpTimes[dumpStep + 1] - t;
    }
    Logger(INFO) << "Time  > dt = " << timeStep << " selected.";
#else
    (void)t;
    timeStep = config.timeStep;
#endif
}

// ---- stage 5: least-squares gradients of rho, v, P and their slope limiter ----
void MeshlessScheme::gradientsAndLimiter() {
    Logger(INFO) << "    > Computing gradients";
#if PERIODIC_BOUNDARIES
    Particles &gh = ghostParticles;
    particles->updateGhostState(gh);
    particles->compPsijTilde(helper, gh, config.kernelSize);
    particles->gradient(particles->rho, particles->rhoGrad, gh.rho, gh);
    particles->gradient(particles->vx, particles->vxGrad, gh.vx, gh);
    particles->gradient(particles->vy, particles->vyGrad, gh.vy, gh);
#if DIM == 3
    particles->gradient(particles->vz, particles->vzGrad, gh.vz, gh);
#endif
    particles->gradient(particles->P, particles->PGrad, gh.P, gh);
    Logger(DEBUG) << "      > Update ghost gradients";
    particles->updateGhostGradients(gh);
#if SLOPE_LIMITING
    Logger(DEBUG) << "      > Limiting slopes";
    particles->slopeLimiter(config.kernelSize, &gh);
    Logger(DEBUG) << "      > Update limited ghost gradients";
    particles->updateGhostGradients(gh);
#endif
#else
    particles->compPsijTilde(helper, config.kernelSize);
    particles->gradient(particles->rho, particles->rhoGrad);
    particles->gradient(particles->vx, particles->vxGrad);
    particles->gradient(particles->vy, particles->vyGrad);
#if DIM == 3
    particles->gradient(particles->vz, particles->vzGrad);
#endif
    particles->gradient(particles->P, particles->PGrad);
#if SLOPE_LIMITING
    Logger(DEBUG) << "      > Limiting slopes";
    particles->slopeLimiter(config.kernelSize);
#endif
#endif
}

// ---- stage 6: effective faces and reconstructed states ----
void MeshlessScheme::prepareRiemannProblems() {
    Logger(INFO) << "    > Preparing Riemann solver";
    Logger(DEBUG) << "      > Computing effective faces";
    particles->compEffectiveFace();
#if PERIODIC_BOUNDARIES
    particles->compEffectiveFace(ghostParticles);
#endif
    Logger(DEBUG) << "      > Computing fluxes";
    particles->compRiemannStatesLR(timeStep, config.kernelSize, config.gamma);
#if PERIODIC_BOUNDARIES
    Logger(DEBUG) << "      > Computing ghost fluxes";
    particles->compRiemannStatesLR(timeStep, config.kernelSize, config.gamma, ghostParticles);
#endif
}

// ---- snapshots ----
bool MeshlessScheme::snapshotDue(int step) {
#if ADAPTIVE_TIMESTEP
    (void)step;
    const bool due = dump;
    dump = false;
    return due;
#else
    return step % config.h5DumpInterval == 0;
#endif
}

void MeshlessScheme::writeSnapshot(double t, int step) {
    const auto io0 = Clock::now();
    Logger(INFO) << "   > Dump particle distribution";
    std::stringstream name;
#if ADAPTIVE_TIMESTEP
    (void)step;
    name << std::setw(6) << std::setfill('0') << dumpStep;
#else
    name << std::setw(6) << std::setfill('0') << step;
#endif
    Logger(INFO) << "      > Dump particles to file";
    particles->dump2file(config.outDir + "/" + name.str() + std::string(".h5"), t);
    ++dumpStep;
    ioSeconds += std::chrono::duration<double>(Clock::now() - io0).count();
}

// ---- stage 7: fluxes and the conserved update ----
void MeshlessScheme::solveAndUpdate() {
    Logger(INFO) << "    > Solving Riemann problems";
    particles->solveRiemannProblems(config.gamma, ghostParticles);
#if DEBUG_LVL
    Logger(DEBUG) << "    > Checking flux symmetry";
    particles->checkFluxSymmetry(&ghostParticles);
#endif
    Logger(INFO) << "    > Collecting fluxes";
    particles->collectFluxes(helper, ghostParticles);
    Logger(INFO) << "    > Updating state";
    particles->updateStateAndPosition(timeStep, domain);
}

void MeshlessScheme::run() {
    double t = 0;
    int step = 0;
    ioSeconds = 0.;
    const auto tStart = Clock::now();

#if ADAPTIVE_TIMESTEP
    numDumpTimes = (int)(config.timeEnd / config.timeStep) / config.h5DumpInterval + 1;
    Logger(DEBUG) << "      > Times for file dump: " << numDumpTimes;
    dumpTimes.assign(numDumpTimes > 0 ? numDumpTimes : 1, 0.);
    for (int iDump = 0; iDump < numDumpTimes; ++iDump) {
        dumpTimes[iDump] = iDump * config.timeStep * config.h5DumpInterval;
        Logger(DEBUG) << "        dumpTimes[" << iDump << "] = " << dumpTimes[iDump];
    }
    dump = true;
    dumpNext = false;
#endif
    dumpStep = 0;

    do {
        Logger(INFO) << "  > TIME: " << t << ", STEP: " << step;
        rebuildGrid();
        searchNeighbours();
        densityAndPressure();
        chooseTimeStep(t);
        gradientsAndLimiter();
        prepareRiemannProblems();
        if (snapshotDue(step)) writeSnapshot(t, step);
        if (t >= config.timeEnd) {
            Logger(INFO) << "    > t = " << t << " -> FINISHED!";
            break;
        }
        solveAndUpdate();
        t += timeStep;
        ++step;
#if ADAPTIVE_TIMESTEP
        if (dumpNext) {
            dump = true;
            dumpNext = false;
        }
#endif
    } while (t < config.timeEnd + timeStep);
    steps = step;
    loopSeconds = std::chrono::duration<double>(Clock::now() - tStart).count() - ioSeconds;
}
