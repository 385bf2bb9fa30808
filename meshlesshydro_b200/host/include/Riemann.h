// One face's Riemann problem, interface of /root/reference/demonstrator/include/Riemann.h:14-44.
// The constructor normalises the effective face and rotates both velocity vectors into the face frame IN PLACE
// (as the reference does, Riemann.cpp:7-81); exact() hands the face to the device solver (mlh_riemann_faces ->
// csrc/k4_flux.cu k_face_setup / k_face_iterate / k_face_finish: rotation, exact Riemann solver, rotation back,
// projection on Aij).  Riemann::exactBatch solves many faces in one call -- what a GPU caller should use.
#ifndef MESHLESSHYDRO_RIEMANN_H
#define MESHLESSHYDRO_RIEMANN_H

#include <cmath>

#include "Helper.h"
#include "parameter.h"
#include "Logger.h"

class Riemann {
public:
    /// WR, WL and Aij must be pre-allocated; W = [rho, P, vx, vy(, vz)]
    Riemann(double *WR, double *WL, double *vFrame, double *Aij, int i);

    /// flux through the face [mass, energy, px, py(, pz)], Fij pre-allocated (DIM+2)
    void exact(double *Fij, const double &gamma);

    /// n faces at once; arrays as in mlh_riemann_faces (n x (DIM+2), n x DIM), unrotated states
    static void exactBatch(long n, const double *WR, const double *WL, const double *vFrame, const double *Aij, double *Fij,
                           const double &gamma);

private:
    int i;
    double *WR, *WL, *vFrame, *Aij;
    double WR0[DIM + 2], WL0[DIM + 2]; // states as handed in (the device solver rotates them itself)
    double hatAij[DIM];
    double unitX[DIM] = {1, 0
#if DIM == 3
                         , 0
#endif
    };
};

#endif // MESHLESSHYDRO_RIEMANN_H
