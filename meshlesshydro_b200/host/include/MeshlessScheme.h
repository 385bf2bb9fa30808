// Time-loop driver of the GPU build.  Public surface = the reference's class (/root/reference/demonstrator/include/
// MeshlessScheme.h:16-44): `MeshlessScheme::Configuration`, the three-argument constructor and `run()`, so that the
// reference's main() compiles against it unchanged.  Everything else here is this build's own organisation: the loop
// body of the reference's run() (MeshlessScheme.cpp:39-253) is split into one private method per stage of the step.
#ifndef DEMONSTRATOR_MESHLESSSCHEME_H
#define DEMONSTRATOR_MESHLESSSCHEME_H

#include <chrono>
#include <iomanip>
#include <string>
#include <vector>

#include "parameter.h"
#include "Domain.h"
#include "Helper.h"
#include "InitialDistribution.h"
#include "Logger.h"

class MeshlessScheme {
public:
    /// run-time settings; one field per key of config.info (main.cpp fills it)
    struct Configuration {
        std::string initFile;              ///< `initFile`: HDF5 initial distribution
        std::string outDir;                ///< `outDir`: where the snapshots go
        double timeStep;                   ///< `timeStep`: fixed dt, or the dump cadence when ADAPTIVE_TIMESTEP
        double timeEnd;                    ///< `timeEnd`
        int h5DumpInterval;                ///< `h5DumpInterval`: steps (resp. multiples of timeStep) between snapshots
        double periodicBoxLimits[2 * DIM]; ///< `periodicBoxLimits`: [lowerX, lowerY(, lowerZ), upperX, upperY(, upperZ)]
        double kernelSize;                 ///< `kernelSize`: support radius h of the cubic spline
        double gamma;                      ///< `gamma`: adiabatic index of the ideal gas
    };

    MeshlessScheme(Configuration config, Particles *particles, Domain::Cell domain);
    ~MeshlessScheme();

    void run();

    // ---- additions of the GPU build ----
    int stepsDone() const { return steps; }
    double secondsInLoop() const { return loopSeconds; } ///< wall time of run() without snapshot I/O

private:
    using Clock = std::chrono::steady_clock;

    // one method per stage of the reference's loop body, in call order
    void rebuildGrid();                   // non-periodic runs: bounding box -> Domain (MeshlessScheme.cpp:41-51)
    void searchNeighbours();              // cell assignment, ghosts, neighbour lists (:52-71)
    void densityAndPressure();            // :72-89
    void chooseTimeStep(double t);        // :91-105, dump schedule of quirk Q7
    void gradientsAndLimiter();           // :107-141
    void prepareRiemannProblems();        // :143-162
    bool snapshotDue(int step);           // :164-170
    void writeSnapshot(double t, int step); // :171-199
    void solveAndUpdate();                // :205-228

    Configuration config;
    Particles *particles;     // borrowed from main(), as in the reference
    Particles ghostParticles; // placeholder: periodic images are implicit on the device
    Domain domain;
    Helper helper{};
    double timeStep;

    // dump schedule (ADAPTIVE_TIMESTEP)
    std::vector<double> dumpTimes;
    int numDumpTimes{0};
    int dumpStep{0};
    bool dump{true}, dumpNext{false};

    // bookkeeping of this build
    int steps{0};
    double loopSeconds{0.}, ioSeconds{0.};
};

#endif // DEMONSTRATOR_MESHLESSSCHEME_H
