// Host-only self test of the C++ mirror classes (no GPU needed): prints "key value" lines that
// tests/test_host_cpu.py checks.  usage: host_selftest <config.info>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <iostream>

#include "../include/ConfigParser.h"
#include "../include/Domain.h"
#include "../include/Helper.h"
#include "../include/InitialDistribution.h"
#include "../include/Logger.h"
#include "../include/Particles.h"

structlog LOGCFG = {};

// host_selftest shard <init.h5> <kernelSize> <nranks> [box limits...]: the slab decomposition of the multi-rank launcher
static int shardMode(int argc, char **argv) {
    InitialDistribution init{argv[2]};
    Particles p{init.getNumberOfParticles()};
    init.getAllParticles(p);
    const double h = std::atof(argv[3]);
    const int nranks = std::atoi(argv[4]);
    double box[2 * DIM] = {0.};
    for (int k = 0; k < 2 * DIM && 5 + k < argc; ++k) box[k] = std::atof(argv[5 + k]);
    p.configureDevice(h, 5. / 3., box);
    for (int r = 0; r < nranks; ++r) {
        std::printf("shard%d", r);
        for (int i : p.slabParticles(r, nranks)) std::printf(" %d", i);
        std::printf("\n");
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 4 && std::string(argv[1]) == "shard") return shardMode(argc, argv);
    std::printf("DIM %d\n", DIM);
    if (argc > 1) {
        ConfigParser c{argv[1]};
        std::printf("initFile %s\n", c.getVal<std::string>("initFile").c_str());
        std::printf("timeStep %.17g\n", c.getVal<double>("timeStep"));
        std::printf("h5DumpInterval %d\n", c.getVal<int>("h5DumpInterval"));
        std::printf("gamma %.17g\n", c.getVal<double>("gamma"));
        ConfigParser box = c.getObj("periodicBoxLimits");
        std::printf("upperX %.17g\n", box.getVal<double>("upperX"));
        std::printf("nested %.17g\n", c.getVal<double>("periodicBoxLimits.lowerY"));
        try {
            c.getVal<double>("doesNotExist");
            std::printf("missing no-throw\n");
        } catch (const std::exception &e) {
            std::printf("missing throws\n");
        }
        int n = 0;
        for (double v : c.getList<double>("someList")) std::printf("list%d %.17g\n", n++, v);
        n = 0;
        for (ConfigParser o : c.getObjList("objects")) std::printf("obj%d %s\n", n++, o.getVal<std::string>("name").c_str());
    }
    // Helper
    {
        Helper h;
#if DIM == 2
        double A[4] = {4., 1., 1., 3.};
        h.inverseMatrix(A, 2);
        std::printf("inv %.17g %.17g %.17g %.17g\n", A[0], A[1], A[2], A[3]);
        double a[2] = {0.6, 0.8}, b[2] = {1., 0.}, L[4];
        Helper::rotationMatrix2D(a, b, L);
        std::printf("rot %.17g %.17g %.17g %.17g\n", L[0], L[1], L[2], L[3]);
#else
        double A[9] = {4., 1., .5, 1., 3., .25, .5, .25, 2.};
        h.inverseMatrix(A, 3);
        std::printf("inv");
        for (double v : A) std::printf(" %.17g", v);
        std::printf("\n");
        double a[3] = {0.36, 0.48, 0.8}, b[3] = {1., 0., 0.}, L[9];
        Helper::rotationMatrix3D(a, b, L);
        std::printf("rot");
        for (double v : L) std::printf(" %.17g", v);
        std::printf("\n");
        double c3[3];
        Helper::crossProduct(a, b, c3);
        std::printf("cross %.17g %.17g %.17g\n", c3[0], c3[1], c3[2]);
#endif
        std::printf("dot %.17g\n", Helper::dotProduct(a, a));
    }
    // Domain
    {
#if DIM == 2
        double lim[4] = {0., 0., 1., 1.};
#else
        double lim[6] = {-.5, -.5, -.5, .5, .5, .5};
#endif
        Domain d{Domain::Cell{lim}};
        d.createGrid(0.07);
        std::printf("cells %d %d numGridCells %d cellSizeX %.17g\n", d.cellsX, d.cellsY, d.numGridCells, d.cellSizeX);
        int nb[27];
        d.getNeighborCells(0, nb);
        std::printf("nb0");
        for (int k = 0; k < (DIM == 2 ? 9 : 27); ++k) std::printf(" %d", nb[k]);
        std::printf("\n");
        std::printf("cell1 %.17g %.17g\n", d.grid[1].minX, d.grid[1].maxX);
    }
    // Particles: host-side pieces
    {
        Particles p{4};
        double xs[4] = {0.1, -0.3, 0.7, 0.2};
        for (int i = 0; i < 4; ++i) {
            p.x[i] = xs[i];
            p.y[i] = -xs[i];
#if DIM == 3
            p.z[i] = 2. * xs[i];
#endif
        }
        double lim[2 * DIM];
        p.getDomainLimits(lim);
        std::printf("limits");
        for (double v : lim) std::printf(" %.17g", v);
        std::printf("\n");
        std::printf("pairwise %.17g\n", p.pairwiseLimiter(1.3, 1.0, 2.0, 0.5, 1.0));
    }
    // Logger
    LOGCFG.level = INFO;
    Logger(DEBUG) << "hidden";
    Logger(INFO) << "logger " << 42;
    return 0;
}
