#include "../include/InitialDistribution.h"

#include <stdexcept>

#include "../include/H5Lite.h"

InitialDistribution::InitialDistribution(const std::string &file) {
    H5Lite::File h5file(file);
    h5file.read("/m", m);
    h5file.read("/x", x);
    h5file.read("/v", v);
    h5file.read("/u", u);
    h5file.read("/materialId", matId);
    if (x.size() == v.size() && x.size() == m.size() && x.size() == u.size() && x.size() == matId.size()) {
        numberOfParticles = (int)x.size();
    } else {
        throw std::length_error("Length mismatch between mass, position and/or velocity vectors.");
    }
    if (numberOfParticles > 0 && ((int)x[0].size() < DIM || (int)v[0].size() < DIM))
        throw std::length_error("Initial distribution has fewer than DIM coordinates per particle.");
}

void InitialDistribution::getAllParticles(Particles &p) {
    for (int i = 0; i < numberOfParticles; ++i) {
        p.m[i] = m[i];
        p.u[i] = u[i];
        p.matId[i] = matId[i];
        p.x[i] = x[i][0];
        p.vx[i] = v[i][0];
        p.y[i] = x[i][1];
        p.vy[i] = v[i][1];
#if DIM == 3
        p.z[i] = x[i][2];
        p.vz[i] = v[i][2];
#endif
    }
    p.markHostStateChanged();
}
