// MeshlessScheme::run -- the loop of /root/reference/demonstrator/src/MeshlessScheme.cpp:21-254 with the same phase
// order, log lines, time-step policy and dump schedule (quirk Q7: after file 000000 the first dump target is
// dumpTimes[2]; the read past the end of dumpTimes is guarded).  Every `particles->` call is one of the reference's
// method names; what it launches on the device is listed in Particles.h.
#include "../include/MeshlessScheme.h"

#include <chrono>
#include <sstream>
#include <vector>

MeshlessScheme::MeshlessScheme(Configuration config_, Particles *particles_, Domain::Cell bounds)
    : config{config_}, timeStep{config_.timeStep}, particles{particles_}, ghostParticles(DIM * particles_->N, true), domain(bounds) {
    Logger(INFO) << "    > Creating grid ... ";
    domain.createGrid(config.kernelSize);
    Logger(INFO) << "    > ... got " << domain.numGridCells << " cells";
    particles->configureDevice(config.kernelSize, config.gamma, config.periodicBoxLimits);
}

MeshlessScheme::~MeshlessScheme() {}

void MeshlessScheme::run() {
    double t = 0;
    int step = 0;
    using clock = std::chrono::steady_clock;
    double ioSeconds = 0.;
    const auto tStart = clock::now();

#if ADAPTIVE_TIMESTEP
    const int numDumpTimes = (int)(config.timeEnd / config.timeStep) / config.h5DumpInterval + 1;
    Logger(DEBUG) << "      > Times for file dump: " << numDumpTimes;
    std::vector<double> dumpTimes(numDumpTimes > 0 ? numDumpTimes : 1);
    for (int iDump = 0; iDump < numDumpTimes; ++iDump) {
        dumpTimes[iDump] = iDump * config.timeStep * config.h5DumpInterval;
        Logger(DEBUG) << "        dumpTimes[" << iDump << "] = " << dumpTimes[iDump];
    }
    bool dump = true;
    bool dumpNext = false;
#endif
    int dumpStep = 0;

    do {
        Logger(INFO) << "  > TIME: " << t << ", STEP: " << step;
#if !PERIODIC_BOUNDARIES
        Logger(INFO) << "    > Computing domain limits ...";
        double domainLimits[DIM * 2];
        particles->getDomainLimits(domainLimits);
        Domain::Cell boundingBox{domainLimits};
        domain.bounds = boundingBox;
        domain.printout();
        Logger(DEBUG) << "      > ... creating grid ...";
        domain.createGrid(config.kernelSize);
        Logger(INFO) << "    > ... done.";
#endif
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
        Logger(INFO) << "    > Computing density";
        particles->compDensity(config.kernelSize);
#if PERIODIC_BOUNDARIES
        particles->compDensity(ghostParticles, config.kernelSize);
#endif
        Logger(INFO) << "    > Computing pressure";
        particles->compPressure(config.gamma);

        if (LOGCFG.level <= DEBUG) { // the sums cost a device reduction + read-back: only when they are printed
            Logger(DEBUG) << "      SANITY CHECK > V_tot = " << particles->sumVolume();
            Logger(DEBUG) << "      SANITY CHECK > M_tot = " << particles->sumMass();
            Logger(DEBUG) << "      SANITY CHECK > E_tot = " << particles->sumEnergy();
            Logger(DEBUG) << "      SANITY CHECK > px_tot = " << particles->sumMomentumX();
            Logger(DEBUG) << "      SANITY CHECK > py_tot = " << particles->sumMomentumY();
#if DIM == 3
            Logger(DEBUG) << "      SANITY CHECK > pz_tot = " << particles->sumMomentumZ();
#endif
        }

#if ADAPTIVE_TIMESTEP
        Logger(INFO) << "    > Selecting global timestep ... ";
        timeStep = particles->compGlobalTimestep(config.gamma, config.kernelSize);
        if (dumpStep >= numDumpTimes) {
            Logger(ERROR) << "Simulation did not abort after reaching timeEnd. Exiting.";
            exit(9);
        } else if (dumpStep + 1 < numDumpTimes && t + timeStep >= dumpTimes[dumpStep + 1]) {
            // (the reference reads dumpTimes[dumpStep+1] unguarded, one past the end for the last target)
            dumpNext = true;
            timeStep = dumpTimes[dumpStep + 1] - t;
        }
        Logger(INFO) << "Time  > dt = " << timeStep << " selected.";
#else
        timeStep = config.timeStep;
#endif

        Logger(INFO) << "    > Computing gradients";
#if PERIODIC_BOUNDARIES
        particles->updateGhostState(ghostParticles);
        particles->compPsijTilde(helper, ghostParticles, config.kernelSize);
        particles->gradient(particles->rho, particles->rhoGrad, ghostParticles.rho, ghostParticles);
        particles->gradient(particles->vx, particles->vxGrad, ghostParticles.vx, ghostParticles);
        particles->gradient(particles->vy, particles->vyGrad, ghostParticles.vy, ghostParticles);
#if DIM == 3
        particles->gradient(particles->vz, particles->vzGrad, ghostParticles.vz, ghostParticles);
#endif
        particles->gradient(particles->P, particles->PGrad, ghostParticles.P, ghostParticles);
        Logger(DEBUG) << "      > Update ghost gradients";
        particles->updateGhostGradients(ghostParticles);
#if SLOPE_LIMITING
        Logger(DEBUG) << "      > Limiting slopes";
        particles->slopeLimiter(config.kernelSize, &ghostParticles);
        Logger(DEBUG) << "      > Update limited ghost gradients";
        particles->updateGhostGradients(ghostParticles);
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

#if ADAPTIVE_TIMESTEP
        if (dump) {
            dump = false;
#else
        if (step % config.h5DumpInterval == 0) {
#endif
            const auto io0 = clock::now();
            std::stringstream stepss;
            Logger(INFO) << "   > Dump particle distribution";
            stepss << std::setw(6) << std::setfill('0')
#if ADAPTIVE_TIMESTEP
                   << dumpStep;
#else
                   << step;
#endif
            Logger(INFO) << "      > Dump particles to file";
            particles->dump2file(config.outDir + "/" + stepss.str() + std::string(".h5"), t);
            ++dumpStep;
            ioSeconds += std::chrono::duration<double>(clock::now() - io0).count();
        }
        if (t >= config.timeEnd) {
            Logger(INFO) << "    > t = " << t << " -> FINISHED!";
            break;
        }

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
    loopSeconds = std::chrono::duration<double>(clock::now() - tStart).count() - ioSeconds;
}
