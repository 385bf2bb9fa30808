// HDF5 initial-condition reader, interface of /root/reference/demonstrator/include/InitialDistribution.h:12-25
// (datasets /m, /x, /v, /u, /materialId; InitialDistribution.cpp:7-30), on H5Lite instead of HighFive.
#ifndef MESHLESSHYDRO_INITIALDISTRIBUTION_H
#define MESHLESSHYDRO_INITIALDISTRIBUTION_H

#include <string>
#include <vector>

#include "Particles.h"

class InitialDistribution {
public:
    /// reads the whole file; throws std::length_error when the datasets disagree in length
    InitialDistribution(const std::string &file);
    int getNumberOfParticles() const { return numberOfParticles; };
    /// fills the host mirrors of `particles` (m, u, matId, x, y[, z], vx, vy[, vz]) and marks them for upload
    void getAllParticles(Particles &particles);

private:
    int numberOfParticles{0};
    std::vector<double> mass{}, energy{};                      // /m, /u          f64[N]
    std::vector<std::vector<double>> position{}, velocity{};   // /x, /v          f64[N][>=DIM]
    std::vector<int> material{};                               // /materialId     i8 | i32 [N]
};

#endif // MESHLESSHYDRO_INITIALDISTRIBUTION_H
