// Entry point with the command line and config keys of /root/reference/demonstrator/src/main.cpp:16-117
// (-c/--config, -v/--verbose, -s/--silent, -h/--help; initFile, outDir, timeStep, timeEnd, h5DumpInterval,
// kernelSize, gamma, periodicBoxLimits{lowerX..upperZ}).  No cxxopts / Boost.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "../include/ConfigParser.h"
#include "../include/Logger.h"
#include "../include/MeshlessScheme.h"
#include "../include/MultiGpu.h"

structlog LOGCFG = {};

static void usage() {
    std::cout << "Demonstrator for the meshless hydrodynamic simulation methods MFV and MFM (B200 build)\n"
                 "Usage:\n  mlh [OPTION...]\n\n"
                 "  -c, --config arg  Path to config file (default: config.info)\n"
                 "  -v, --verbose     More printouts for debugging\n"
                 "  -s, --silent      Suppress normal printouts\n"
                 "  -n, --ranks arg   GPUs of this node to use, one process each (default: $MLH_RANKS or 1)\n"
                 "  -h, --help        Show this help\n";
}

int main(int argc, char *argv[]) {
    std::string configFile = "config.info";
    bool verbose = false, silent = false;
    int ranks = std::getenv("MLH_RANKS") ? std::atoi(std::getenv("MLH_RANKS")) : 1;
    for (int a = 1; a < argc; ++a) {
        const std::string arg = argv[a];
        if (arg == "-h" || arg == "--help") {
            usage();
            return 0;
        } else if (arg == "-v" || arg == "--verbose") {
            verbose = true;
        } else if (arg == "-s" || arg == "--silent") {
            silent = true;
        } else if ((arg == "-c" || arg == "--config") && a + 1 < argc) {
            configFile = argv[++a];
        } else if (arg.rfind("--config=", 0) == 0) {
            configFile = arg.substr(9);
        } else if ((arg == "-n" || arg == "--ranks") && a + 1 < argc) {
            ranks = std::atoi(argv[++a]);
        } else if (arg.rfind("--ranks=", 0) == 0) {
            ranks = std::atoi(arg.substr(8).c_str());
        } else {
            std::cerr << "Option '" << arg << "' does not exist" << std::endl;
            return 1;
        }
    }
    ConfigParser confP{configFile};

    LOGCFG.headers = true;
    LOGCFG.level = verbose ? DEBUG : INFO;
    if (silent) {
        if (verbose) throw std::invalid_argument("Command line options -s and -v are incompatible");
        LOGCFG.level = WARN;
    }

    Logger(INFO) << "Reading configuration ... ";
    MeshlessScheme::Configuration config;
    config.initFile = confP.getVal<std::string>("initFile");
    Logger(INFO) << "    > Initial distribution: " << config.initFile;
    config.outDir = confP.getVal<std::string>("outDir");
    Logger(INFO) << "    > Output directory: " << config.outDir;
    config.timeStep = confP.getVal<double>("timeStep");
    Logger(INFO) << "    > Time step: " << config.timeStep;
    config.timeEnd = confP.getVal<double>("timeEnd");
    Logger(INFO) << "    > End of simulation: " << config.timeEnd;
    config.h5DumpInterval = confP.getVal<int>("h5DumpInterval");
    Logger(INFO) << "    > Dump data to h5 file every " << config.h5DumpInterval << " steps";
    config.kernelSize = confP.getVal<double>("kernelSize");
    Logger(INFO) << "    > Using global kernel size h = " << config.kernelSize;
    config.gamma = confP.getVal<double>("gamma");
    Logger(INFO) << "    > Adiabatic index for ideal gas EOS gamma = " << config.gamma;
    for (int k = 0; k < 2 * DIM; ++k) config.periodicBoxLimits[k] = 0.;
#if PERIODIC_BOUNDARIES
    auto periodicBoxLimits = confP.getObj("periodicBoxLimits");
    config.periodicBoxLimits[0] = periodicBoxLimits.getVal<double>("lowerX");
    config.periodicBoxLimits[DIM] = periodicBoxLimits.getVal<double>("upperX");
    config.periodicBoxLimits[1] = periodicBoxLimits.getVal<double>("lowerY");
    config.periodicBoxLimits[DIM + 1] = periodicBoxLimits.getVal<double>("upperY");
#if DIM == 3
    config.periodicBoxLimits[2] = periodicBoxLimits.getVal<double>("lowerZ");
    config.periodicBoxLimits[DIM + 2] = periodicBoxLimits.getVal<double>("upperZ");
#endif
    std::string periodicBoxStr = "[";
    for (int i = 0; i < 2 * DIM; i++) {
        periodicBoxStr.append(std::to_string(config.periodicBoxLimits[i]));
        if (i < 2 * DIM - 1) periodicBoxStr.append(", ");
    }
    Logger(INFO) << "    > Periodic boundaries within box: " << periodicBoxStr << "]";
#endif

    mgpu::plan(ranks); // before the particle arrays exist: with several ranks they live in shared memory
    Logger(INFO) << "    > Reading initial distribution ...";
    InitialDistribution initDist{config.initFile};
    Particles particles{initDist.getNumberOfParticles()};
    initDist.getAllParticles(particles);
    Logger(INFO) << "    > N = " << particles.N;
    if (mgpu::planned() > 1) {
        Logger(INFO) << "    > Slab decomposition over " << mgpu::planned() << " GPUs (one process each)";
        mgpu::launch(); // no CUDA call so far; the parent continues as rank 0
        // every rank keeps the SAME verbosity level: it decides which collective sanity sums the time loop evaluates
        // (MeshlessScheme::densityAndPressure), so it must not differ between ranks; only rank 0 prints (errors: all)
        LOGCFG.myRank = mgpu::rank();
        LOGCFG.outputRank = 0;
    }
    Logger(INFO) << "... done. Initializing simulation ...";

#if PERIODIC_BOUNDARIES
    double *domainLimits = config.periodicBoxLimits;
#else
    double domainLimits[DIM * 2];
    particles.getDomainLimits(domainLimits);
#endif
    Domain::Cell boundingBox{domainLimits};
    MeshlessScheme algorithm{config, &particles, boundingBox};
    Logger(INFO) << "... done.";

    Logger(INFO) << "Starting time integration ...";
    algorithm.run();
    Logger(INFO) << "... done.";
    if (algorithm.stepsDone() > 0)
        Logger(INFO) << "    > " << algorithm.stepsDone() << " steps, " << algorithm.secondsInLoop() << " s in the time loop ("
                     << (double)particles.N * algorithm.stepsDone() / algorithm.secondsInLoop() << " particle-updates/s, "
                     << particles.kernelLaunches() << " kernel launches" << (mgpu::active() ? " on rank 0)" : ")");
    return mgpu::finish(0);
}
