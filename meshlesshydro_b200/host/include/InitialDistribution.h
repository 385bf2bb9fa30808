// HDF5 initial-condition reader, interface of /root/reference/demonstrator/include/InitialDistribution.h:12-25
// (datasets /m, /x, /v, /u, /materialId; InitialDistribution.cpp:7-30), on H5Lite instead of HighFive.
#ifndef MESHLESSHYDRO_INITIALDISTRIBUTION_H
#define MESHLESSHYDRO_INITIALDISTRIBUTION_H

#include <string>
#include <vector>

#include "Particles.h"

class InitialDistribution {
public:
    InitialDistribution(const std::string &file);
    int getNumberOfParticles() const { return numberOfParticles; };
    void getAllParticles(Particles &particles);

private:
    std::vector<double> m{}, u{};
    std::vector<std::vector<double>> x{}, v{};
    std::vector<int> matId{};
    int numberOfParticles{0};
};

#endif // MESHLESSHYDRO_INITIALDISTRIBUTION_H
