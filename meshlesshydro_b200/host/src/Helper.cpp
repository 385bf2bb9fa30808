#include "../include/Helper.h"

#include <cstdlib>

// column-major N x N, unblocked dgetf2 + dtrti2/dgetri (the matrices here are symmetric, so the storage order
// only matters for round-off; same operation order as the device routine inverse_lu in csrc/k3_density.cu)
void Helper::inverseMatrix(double *A, int n) {
    if (n < 1 || n > 3) {
        Logger(ERROR) << "Helper::inverseMatrix supports N <= 3. - Aborting.";
        exit(9);
    }
    int ipiv[3];
#define AA(r, c) A[(r) + (c) * n]
    for (int j = 0; j < n; ++j) {
        int pv = j;
        double amax = std::fabs(AA(j, j));
        for (int i = j + 1; i < n; ++i)
            if (std::fabs(AA(i, j)) > amax) {
                amax = std::fabs(AA(i, j));
                pv = i;
            }
        ipiv[j] = pv;
        if (amax != 0.) {
            if (pv != j)
                for (int k = 0; k < n; ++k) std::swap(AA(j, k), AA(pv, k));
            const double rcp = 1. / AA(j, j);
            for (int i = j + 1; i < n; ++i) AA(i, j) *= rcp;
        }
        for (int k = j + 1; k < n; ++k)
            for (int i = j + 1; i < n; ++i) AA(i, k) -= AA(i, j) * AA(j, k);
    }
    for (int j = 0; j < n; ++j) { // inverse of U in place
        AA(j, j) = 1. / AA(j, j);
        const double ajj = -AA(j, j);
        for (int k = 0; k < j; ++k) {
            if (AA(k, j) != 0.) {
                const double t = AA(k, j);
                for (int i = 0; i < k; ++i) AA(i, j) += t * AA(i, k);
                AA(k, j) *= AA(k, k);
            }
        }
        for (int i = 0; i < j; ++i) AA(i, j) *= ajj;
    }
    for (int j = n - 2; j >= 0; --j) { // solve inv(A) L = inv(U)
        double work[3];
        for (int i = j + 1; i < n; ++i) {
            work[i] = AA(i, j);
            AA(i, j) = 0.;
        }
        for (int k = j + 1; k < n; ++k)
            for (int i = 0; i < n; ++i) AA(i, j) -= AA(i, k) * work[k];
    }
    for (int j = n - 2; j >= 0; --j)
        if (ipiv[j] != j)
            for (int i = 0; i < n; ++i) std::swap(AA(i, j), AA(i, ipiv[j]));
#undef AA
}

double Helper::dotProduct(double *a, double *b) {
    double res = 0.;
    for (int k = 0; k < DIM; ++k) res += a[k] * b[k];
    return res;
}

void Helper::crossProduct(double *a, double *b, double *c) {
#if DIM == 3
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
#else
    (void)a; (void)b; (void)c;
    Logger(ERROR) << "Cross product not defined for 2D. - Aborting.";
    exit(9);
#endif
}

void Helper::rotationMatrix2D(double *a, double *b, double *L) {
    const double c = a[0] * b[0] + a[1] * b[1], s = a[0] * b[1] - a[1] * b[0];
    L[0] = c;
    L[1] = -s;
    L[2] = s;
    L[3] = c;
}

#if DIM == 3
// Rodrigues' formula with n = 1/(1 + cos); singular for a = -b, as the reference (Helper.cpp:48-77)
void Helper::rotationMatrix3D(double *a, double *b, double *L) {
    double v[3];
    crossProduct(a, b, v);
    const double n = 1. / (1. + dotProduct(a, b));
    L[0] = 1. - n * (v[2] * v[2] + v[1] * v[1]);
    L[1] = -v[2] + n * v[0] * v[1];
    L[2] = v[1] + n * v[0] * v[2];
    L[3] = v[2] + n * v[0] * v[1];
    L[4] = 1. - n * (v[2] * v[2] + v[0] * v[0]);
    L[5] = -v[0] + n * v[1] * v[2];
    L[6] = -v[1] + n * v[0] * v[2];
    L[7] = v[0] + n * v[1] * v[2];
    L[8] = 1. - n * (v[1] * v[1] + v[0] * v[0]);
}
#endif
