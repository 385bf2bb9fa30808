// Small dense linear algebra of the demonstrator, interface of /root/reference/demonstrator/include/Helper.h:23-47.
// The hot path does these per particle / per face inside the kernels (csrc/k3_density.cu inverse_lu,
// csrc/k4_flux.cu face_rotate); the host versions exist for callers of the class (tests, Riemann).
#ifndef MESHLESSHYDRO_HELPER_H
#define MESHLESSHYDRO_HELPER_H

#include <cmath>

#include "parameter.h"
#include "Logger.h"

class Helper {
public:
    /// in-place inverse of the N x N matrix A (N <= 3): LU with partial pivoting + inversion, the
    /// algorithm of LAPACK dgetrf_/dgetri_ which the reference links (Helper.cpp:7-18)
    void inverseMatrix(double *A, int N);
    static double dotProduct(double *a, double *b);
    static void crossProduct(double *a, double *b, double *crossProduct);
    /// rotation taking the unit vector a onto the unit vector b; Lambda[j + DIM*i] = lambda_ij
    static void rotationMatrix2D(double *a, double *b, double *Lambda);
#if DIM == 3
    static void rotationMatrix3D(double *a, double *b, double *Lambda);
#endif
};

#endif // MESHLESSHYDRO_HELPER_H
