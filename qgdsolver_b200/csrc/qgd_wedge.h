// wedge patches: the face transformation tensor, shared by the boundary kernels (qgd_kernels.cu) and the CPU test shim
#pragma once
#include <cmath>

#ifndef QGD_HD
#ifdef __CUDACC__
#define QGD_HD __host__ __device__ __forceinline__
#else
#define QGD_HD inline
#endif
#endif

namespace qgd {

// ---- wedge patches [OF-v2312 wedgePolyPatch::calcGeometry, rotationTensor]: faceT = rotationTensor(centreNormal, n) with the patch
// normal n and the centre-plane normal = the coordinate axis n is closest to, sign(n_i) (max(|n_i|, 0.5) - 0.5) normalised:
//   faceT = s I + (1 - s) n3 n3 / |n3|^2 + (n n1 - n1 n),  n1 = centreNormal, s = n1 . n, n3 = n1 x n
QGD_HD void wedgeFaceT(const double (&n)[3], double (&T)[9])
{
    double n1[3];
    for (int i = 0; i < 3; ++i) n1[i] = (n[i] >= 0.0 ? 1.0 : -1.0) * (fmax(fabs(n[i]), 0.5) - 0.5);
    const double m1 = sqrt(n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2]);
    for (int i = 0; i < 3; ++i) n1[i] /= m1;
    const double s = n1[0] * n[0] + n1[1] * n[1] + n1[2] * n[2];
    const double n3[3] = {n1[1] * n[2] - n1[2] * n[1], n1[2] * n[0] - n1[0] * n[2], n1[0] * n[1] - n1[1] * n[0]};
    const double m3 = n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = (i == j) ? s : 0.0;
            if (m3 > 1.0e-15) v += (1.0 - s) * n3[i] * n3[j] / m3 + (n[i] * n1[j] - n1[i] * n[j]);
            else v = (i == j) ? 1.0 : 0.0;
            T[3 * i + j] = v;
        }
}


} // namespace qgd
