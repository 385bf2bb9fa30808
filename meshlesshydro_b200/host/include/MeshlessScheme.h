// Time-loop driver, interface of /root/reference/demonstrator/include/MeshlessScheme.h:16-44.
#ifndef DEMONSTRATOR_MESHLESSSCHEME_H
#define DEMONSTRATOR_MESHLESSSCHEME_H

#include <iomanip>
#include <string>

#include "parameter.h"
#include "InitialDistribution.h"
#include "Logger.h"
#include "Domain.h"
#include "Helper.h"

class MeshlessScheme {
public:
    struct Configuration {
        std::string initFile;
        std::string outDir;
        double timeStep;
        double timeEnd;
        int h5DumpInterval;
        double periodicBoxLimits[2 * DIM];
        double kernelSize;
        double gamma; // adiabatic index
    };

    MeshlessScheme(Configuration config, Particles *particles, Domain::Cell domain);
    ~MeshlessScheme();

    void run();
    int stepsDone() const { return steps; }
    double secondsInLoop() const { return loopSeconds; } // wall time of run() without snapshot I/O

private:
    Configuration config;
    double timeStep;
    Particles *particles;
    Particles ghostParticles; // placeholder: periodic images are implicit on the device
    Domain domain;
    Helper helper{};
    int steps{0};
    double loopSeconds{0.};
};

#endif // DEMONSTRATOR_MESHLESSSCHEME_H
