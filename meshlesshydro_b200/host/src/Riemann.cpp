#include "../include/Riemann.h"

#include <cstdlib>

#include "../../../include/mlh_gpu.h"

Riemann::Riemann(double *WR_, double *WL_, double *vFrame_, double *Aij_, int i_) : i{i_}, WR{WR_}, WL{WL_}, vFrame{vFrame_}, Aij{Aij_} {
    for (int k = 0; k < DIM + 2; ++k) {
        WR0[k] = WR[k];
        WL0[k] = WL[k];
    }
    const double AijNorm = sqrt(Helper::dotProduct(Aij, Aij));
    for (int k = 0; k < DIM; ++k) hatAij[k] = 1. / AijNorm * Aij[k];
    double Lambda[DIM * DIM];
#if DIM == 2
    Helper::rotationMatrix2D(hatAij, unitX, Lambda);
#else
    Helper::rotationMatrix3D(hatAij, unitX, Lambda);
#endif
    double bR[DIM], bL[DIM];
    for (int k = 0; k < DIM; ++k) {
        bR[k] = WR[2 + k];
        bL[k] = WL[2 + k];
    }
    for (int a = 0; a < DIM; ++a) {
        double r = 0., l = 0.;
        for (int b = 0; b < DIM; ++b) {
            r += Lambda[DIM * a + b] * bR[b];
            l += Lambda[DIM * a + b] * bL[b];
        }
        WR[2 + a] = r;
        WL[2 + a] = l;
    }
}

namespace {
mlh_ctx *g_ctx = nullptr; // context without particles, one per gamma in use
double g_gamma = 0.;
mlh_ctx *solverContext(double gamma) {
    if (g_ctx && g_gamma == gamma) return g_ctx;
    if (g_ctx) mlh_destroy(g_ctx);
    mlh_config cfg;
    mlh_default_config(&cfg);
    cfg.dim = DIM;
    cfg.periodic = 0;
    cfg.meshless_finite_mass = MESHLESS_FINITE_MASS;
    cfg.gamma = gamma;
    cfg.kernel_size = 1.;
    const char *dev = std::getenv("MLH_DEVICE");
    cfg.device = dev ? std::atoi(dev) : 0;
    int rc = mlh_create(&cfg, &g_ctx);
    if (rc != MLH_OK) {
        Logger(ERROR) << "Riemann: mlh_create failed (" << rc << "): " << mlh_last_error(nullptr) << " - Aborting.";
        exit(rc == MLH_E_NO_DEVICE ? 20 : 21);
    }
    g_gamma = gamma;
    return g_ctx;
}
} // namespace

void Riemann::exactBatch(long n, const double *WR, const double *WL, const double *vFrame, const double *Aij, double *Fij,
                         const double &gamma) {
    mlh_ctx *c = solverContext(gamma);
    int rc = mlh_riemann_faces(c, n, WR, WL, vFrame, Aij, Fij);
    if (rc != MLH_OK) {
        Logger(ERROR) << "mlh_riemann_faces failed (" << rc << "): " << mlh_last_error(c) << " - Aborting.";
        exit(21);
    }
}

void Riemann::exact(double *Fij, const double &gamma) { exactBatch(1, WR0, WL0, vFrame, Aij, Fij, gamma); }
