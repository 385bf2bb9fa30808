#include "../include/InitialDistribution.h"

#include <stdexcept>

#include "../include/H5Lite.h"

InitialDistribution::InitialDistribution(const std::string &file) {
    H5Lite::File h5file(file);
    h5file.read("/m", mass);
    h5file.read("/x", position);
    h5file.read("/v", velocity);
    h5file.read("/u", energy);
    h5file.read("/materialId", material);
    if (position.size() == velocity.size() && position.size() == mass.size() && position.size() == energy.size() &&
        position.size() == material.size()) {
        numberOfParticles = (int)position.size();
    } else {
        throw std::length_error("Length mismatch between mass, position and/or velocity vectors.");
    }
    if (numberOfParticles > 0 && ((int)position[0].size() < DIM || (int)velocity[0].size() < DIM))
        throw std::length_error("Initial distribution has fewer than DIM coordinates per particle.");
}

void InitialDistribution::getAllParticles(Particles &p) {
    for (int i = 0; i < numberOfParticles; ++i) {
        const std::vector<double> &xi = position[i], &vi = velocity[i];
        p.m[i] = mass[i];
        p.u[i] = energy[i];
        p.matId[i] = material[i];
        p.x[i] = xi[0];
        p.y[i] = xi[1];
        p.vx[i] = vi[0];
        p.vy[i] = vi[1];
#if DIM == 3
        p.z[i] = xi[2];
        p.vz[i] = vi[2];
#endif
    }
    p.markHostStateChanged();
}
