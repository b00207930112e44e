// QHDFoam on the device (explicit branch): QHDFoam.C:83-139.
//
//   k_qhd_bnd_eval       correctBoundaryConditions() of U,T or p (fixedValue | zeroGradient | fixedGradient; qhdFlux is a
//                        fixed-gradient patch in this solver)
//   k_qhd_points<W>      volPointInterpolation of (U,T) or p                          [OF-v2312]
//   k_qhd_face_pre       updateFields.H:36-73 + updateFluxes.H:33-38 fused: gradUf, Uf, BdFrcf -> F0 = phiu - phiwo ;
//                        QHDCourantNo.H:37-57
//   k_qhd_cell_pre       pEqn source (QHDpEqn.H:36-43: -V(div phiu - div phiwo) + boundaryCoeffs + setReference) and the
//                        cell-centred Gauss gradient of U used by QHDUEqn.H:76
//   PcgMatrix::solve     QHDpEqn.H:45, one cooperative kernel (qgd_pcg.cu)
//   k_qhd_face_post      QHDpEqn.H:47 phi ; QHDUEqn.H:36-43 ; the four explicit face fluxes of QHDUEqn.H:68-84 and the
//                        three of QHDTEqn.H:83-91 summed into FU(3), FT
//   k_qhd_cell_update    Euler update of U and T, reference shift of p (QHDFoam.C:123-131)
// The pressure matrix -laplacian(tauQGDf/rhof) is constant in time (thermo.correct() is only called from createFields.H:39,
// so tauQGD, rho, mu, alpha never change): it is assembled once, together with the DIC factor.
#include <cfloat>
#include <cstring>
#include <memory>

#include "qgd_kernels.cuh"
#include "qgd_pcg.cuh"

namespace qgd {

struct QhdConsts {
    double rho0, nu, Hi, beta, g[3];
    int needRef, refCell;
    double refValue;
    int implicit;                // QGD::implicitDiffusion: the Laplacians of U and T are solved implicitly (QHDUEqn.H:46-65, QHDTEqn.H:69-80)
    int scalarTransport;         // scalarTransportQHDFoam.C:70-135: U, p frozen; phi = phiu; T equation with -fvc::Sp(fvc::div(phiu),T); Co from mag(Uf)
};

struct QhdView {
    int nCells, nPoints, nPatchPoints;
    int nOwned;                  // cells solved by this rank; [nOwned, nCells) are halo copies refreshed by the exchanges (multi-GPU)
    double* Q;                   // [5][nCells]  Ux,Uy,Uz,T,p
    double* P;                   // [5][nPoints]
    int pcEllW; const int* pcEll; const double* pcEllWt; const int* pcCount;
    const int* pcTailOff; const int* pcTailCell; const double* pcTailW;
    const int* patchPoints; const int* ppOff; const int* ppFace; const double* ppW;
    int cfEllW; const int* cfEll; const int* cfTailOff; const int* cfTailEnc; const double* V;
    const double* tauf;          // [nF]  tauQGDf (device face order)
    const double* upper;         // [nI]  pEqn upper coefficient  -(tauQGDf/rhof) |Sf| nonOrthDeltaCoeffs
    double* F0; double* FU; double* FT; double* phi;      // face fluxes: [nF], [3][nF], [nF], [nF]
    double* GU;                  // [9][nCells]  fvc::grad(U) (Gauss linear), cell values
    double* UB; double* TB; double* pB;                   // boundary values: [3][nB] SoA, [nB], [nB]
    const int* bcU; const int* bcT; const int* bcP;       // per boundary face
    const double* bvU; const double* bvT; const double* bvP;   // value | gradient per boundary face ([3][nB] SoA for U)
    const double* intC; const double* bouC;               // pEqn internalCoeffs / boundaryCoeffs per boundary face
    const double* srcBnd;        // [nCells] sum of boundaryCoeffs of the cell's boundary faces
    const double* diag0;         // [nCells] pEqn diagonal before boundary contributions and setReference
    double* pcgB;                // [nCells] pEqn source handed to the PCG
    // implicit branch: time-constant Laplacian row sums / boundary sources, per-step diagonals and right-hand sides
    const double* sumAU; const double* sumAT; const double* srcBU; const double* srcBT;   // [nC], [nC], [3][nC], [nC]
    double* diagU; double* diagT; double* bU; double* bT;                                 // [nC], [nC], [3][nC], [nC]
    double* shift;               // [1] reference shift of p
    StepScalars* sc;
};

namespace {

constexpr int kB = 256;
inline int nblk(long n, int b = kB) { return (int)((n + b - 1) / b); }

__device__ __forceinline__ unsigned long long dbitsQ(double v) { return (unsigned long long)__double_as_longlong(v); }

template <int BLOCK>
__device__ __forceinline__ void blockReduceCoQ(double coMax, double tauMin, StepScalars* sc)
{
    __shared__ double sMax[BLOCK / 32];
    __shared__ double sMin[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        coMax = fmax(coMax, __shfl_xor_sync(0xffffffffu, coMax, o));
        tauMin = fmin(tauMin, __shfl_xor_sync(0xffffffffu, tauMin, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sMax[wid] = coMax; sMin[wid] = tauMin; }
    __syncthreads();
    if (wid == 0) {
        coMax = (lane < BLOCK / 32) ? sMax[lane] : 0.0;
        tauMin = (lane < BLOCK / 32) ? sMin[lane] : DBL_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            coMax = fmax(coMax, __shfl_xor_sync(0xffffffffu, coMax, o));
            tauMin = fmin(tauMin, __shfl_xor_sync(0xffffffffu, tauMin, o));
        }
        if (lane == 0) { atomicMax(&sc->coMaxBits, dbitsQ(coMax)); atomicMin(&sc->tauMinBits, dbitsQ(tauMin)); }
    }
}

// ---- boundary conditions.  which: 0 -> U and T, 1 -> p.  addShift: p_b += shift (QHDFoam.C:125-130)
__global__ void k_qhd_bnd_eval(FaceView fv, QhdView q, int which, int addShift)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    if (fv.bKind[b] == QGD_PATCH_EMPTY) return;
    const int f = fv.nI + b;
    const int P = fv.own[f];
    const size_t n = q.nCells, nB = fv.nB;
    const double dc = fv.dC[f];
    auto eval = [&](int kind, double cell, double bv) {
        return kind == QGD_BC_FIXED_VALUE ? bv : (kind == QGD_BC_ZERO_GRADIENT ? cell : cell + bv / dc);
    };
    if (which == 0) {
        const int kU = q.bcU[b];
#pragma unroll
        for (int j = 0; j < 3; ++j) q.UB[j * nB + b] = eval(kU, q.Q[j * n + P], q.bvU[j * nB + b]);
        q.TB[b] = eval(q.bcT[b], q.Q[3 * n + P], q.bvT[b]);
    } else if (addShift) {
        q.pB[b] += *q.shift;
    } else {
        const int kP = q.bcP[b];
        q.pB[b] = eval(kP, q.Q[4 * n + P], q.bvP[b]);
    }
}

// ---- cell -> point gather of fields [K0, K0+K)
template <int W, int K0, int K>
__global__ void __launch_bounds__(kB) k_qhd_points(QhdView q)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= q.nPoints) return;
    const int cnt = __ldg(&q.pcCount[p]);
    if (cnt == 0) return;
    const size_t nP = q.nPoints, n = q.nCells;
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = 0.0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int id = __ldg(&q.pcEll[j * nP + p]);
        const double wq = __ldg(&q.pcEllWt[j * nP + p]);
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] += wq * q.Q[(K0 + k) * n + id];
    }
    if (cnt > W)
        for (int t = __ldg(&q.pcTailOff[p]); t < __ldg(&q.pcTailOff[p + 1]); ++t) {
            const int id = __ldg(&q.pcTailCell[t]);
            const double w1 = __ldg(&q.pcTailW[t]);
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] += w1 * q.Q[(K0 + k) * n + id];
        }
#pragma unroll
    for (int k = 0; k < K; ++k) q.P[(K0 + k) * nP + p] = a[k];
}

__global__ void k_qhd_patch_points(QhdView q, int nB, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.nPatchPoints) return;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int t = q.ppOff[i]; t < q.ppOff[i + 1]; ++t) {
        const int b = q.ppFace[t];
        const double wq = q.ppW[t];
        if (which == 0) { a0 += wq * q.UB[b]; a1 += wq * q.UB[(size_t)nB + b]; a2 += wq * q.UB[2 * (size_t)nB + b]; a3 += wq * q.TB[b]; }
        else a0 += wq * q.pB[b];
    }
    const size_t nP = q.nPoints;
    double* o = q.P + q.patchPoints[i];
    if (which == 0) { o[0] = a0; o[nP] = a1; o[2 * nP] = a2; o[3 * nP] = a3; }
    else o[4 * nP] = a0;
}

// ---- shared face algebra
struct FaceGeo { double g1[3], g2[3], gp[3], Sf[3]; };

__device__ __forceinline__ FaceGeo loadGeo(const FaceView& fv, int f)
{
    FaceGeo o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o.g1[i] = __ldg(&fv.G[(size_t)(0 + i) * fv.fs + f]);
        o.g2[i] = __ldg(&fv.G[(size_t)(3 + i) * fv.fs + f]);
        o.Sf[i] = __ldg(&fv.Sf[(size_t)i * fv.fs + f]);
        o.gp[i] = __ldg(&fv.G[6 * (size_t)fv.fs + f]) * o.Sf[i];            // GP = gpS * Sf
    }
    return o;
}

// vertex differences (phi[v1]-phi[v3], phi[v2]-phi[v4]) of point field k
__device__ __forceinline__ void ptDiff(const QhdView& q, int k, const int4& v, int flags, double& d1, double& d2)
{
    if (flags & FF_POINTS) {
        const double* p = q.P + (size_t)k * q.nPoints;
        d1 = p[v.x] - p[v.z];
        d2 = p[v.y] - p[v.w];
    } else { d1 = 0.0; d2 = 0.0; }
}

// gradU[3*i+j] = d_i U_j from the three difference triples (GaussVolPointBase3D.C:831-854)
__device__ __forceinline__ void gradVec(const FaceGeo& ge, int flags, const double (&d1)[3], const double (&d2)[3], const double (&dP)[3], double (&G)[9])
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) G[3 * i + j] = ge.g1[i] * d1[j] + ge.g2[i] * d2[j] + ge.gp[i] * dP[j];
    if (flags & FF_TRI_QUIRK) {
        const double dxx = G[0], dyy = G[4], dzz = G[8];
#pragma unroll
        for (int row = 0; row < 3; ++row) { G[3 * row] = dxx; G[3 * row + 1] = dyy; G[3 * row + 2] = dzz; }
    }
}

// leastSquares gradient of state field k (Q[k]) on internal face f: sum_s c_s (phi_s - phi_f), phi_f = linearInterpolate(phi)
// (extendedFaceStencilScalarGrad.C:52-72); vectors component-wise (leastSquaresStencil.C:145-202)
__device__ __forceinline__ void lsqGradQ(const FaceView& fv, const QhdView& q, int f, int k, double phiP, double phiN, double w, double (&g)[3])
{
    const size_t nI = fv.nI;
    const double* phi = q.Q + (size_t)k * q.nCells;
    const double sF = w * (phiP - phiN) + phiN;
    g[0] = g[1] = g[2] = 0.0;
    for (int s = 0; s < fv.lsqW; ++s) {
        const int c = __ldg(&fv.lsqCells[(size_t)s * nI + f]);
        const double d = phi[c] - sF;
#pragma unroll
        for (int i = 0; i < 3; ++i) g[i] += __ldg(&fv.lsqCoef[((size_t)s * 3 + i) * nI + f]) * d;
    }
}
// gradU[3*i+j] = d_i U_j with the least-squares stencil
__device__ __forceinline__ void lsqGradVecQ(const FaceView& fv, const QhdView& q, int f, const double (&uP)[3], const double (&uN)[3], double w, double (&G)[9])
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double g[3];
        lsqGradQ(fv, q, f, j, uP[j], uN[j], w, g);
        G[j] = g[0]; G[3 + j] = g[1]; G[6 + j] = g[2];
    }
}

// the part of the face evaluation shared by the pre- and post-solve kernels
struct FaceCommon { double Uf[3], Tf, Bf[3], G[9], UgU[3], phiu, tau; };

__device__ __forceinline__ void faceCommon(const QhdConsts& k, const FaceGeo& ge, int flags, const double (&uP)[3], const double (&uN)[3],
                                           double TP, double TN, double w, const double (&d1)[3], const double (&d2)[3], double tau, FaceCommon& o,
                                           bool haveG = false)
{
    const double dP[3] = {uP[0] - uN[0], uP[1] - uN[1], uP[2] - uN[2]};
    if (!haveG) gradVec(ge, flags, d1, d2, dP, o.G);          // haveG: o.G already holds the leastSquares gradient
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.Uf[j] = w * (uP[j] - uN[j]) + uN[j];
        const double bP = (k.beta * TP) * k.g[j], bN = (k.beta * TN) * k.g[j];        // updateFields.H:66-67
        o.Bf[j] = w * (bP - bN) + bN;
    }
    o.Tf = w * (TP - TN) + TN;
#pragma unroll
    for (int j = 0; j < 3; ++j) o.UgU[j] = o.Uf[0] * o.G[j] + o.Uf[1] * o.G[3 + j] + o.Uf[2] * o.G[6 + j];
    o.phiu = ge.Sf[0] * o.Uf[0] + ge.Sf[1] * o.Uf[1] + ge.Sf[2] * o.Uf[2];           // updateFluxes.H:33
    o.tau = tau;
}

__device__ __forceinline__ double phiwoOf(const FaceGeo& ge, const FaceCommon& c)
{
    double wv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) wv[j] = c.tau * (c.UgU[j] - c.Bf[j]);
    return ge.Sf[0] * wv[0] + ge.Sf[1] * wv[1] + ge.Sf[2] * wv[2];                    // updateFluxes.H:35
}

template <bool ADJUST>
__global__ void __launch_bounds__(kB) k_qhd_face_pre(QhdConsts k, FaceView fv, QhdView q)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    double coMax = 0.0, tauMin = DBL_MAX;
    if (f < fv.nI) {
        const size_t n = q.nCells;
        const int P = __ldg(&fv.own[f]), N = __ldg(&fv.nei[f]), flags = __ldg(&fv.flags[f]);
        const int4 v = __ldg(&fv.vtx[f]);
        const FaceGeo ge = loadGeo(fv, f);
        double uP[3], uN[3], d1[3], d2[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { uP[j] = q.Q[j * n + P]; uN[j] = q.Q[j * n + N]; ptDiff(q, j, v, flags, d1[j], d2[j]); }
        const double TP = q.Q[3 * n + P], TN = q.Q[3 * n + N];
        const double tau = __ldg(&q.tauf[f]);
        FaceCommon c;
        const bool lsq = (flags & FF_LSQ) != 0;
        if (lsq) lsqGradVecQ(fv, q, f, uP, uN, __ldg(&fv.w[f]), c.G);
        faceCommon(k, ge, flags, uP, uN, TP, TN, __ldg(&fv.w[f]), d1, d2, tau, c, lsq);
        q.F0[f] = c.phiu - phiwoOf(ge, c);
        if (ADJUST && f < fv.nIActive) {                            // QHDCourantNo.H:39-54 (faces owned by halo cells belong to another rank)
            const double ms = __ldg(&fv.magSf[f]);
            const double Unf = c.Uf[0] * (ge.Sf[0] / ms) + c.Uf[1] * (ge.Sf[1] / ms) + c.Uf[2] * (ge.Sf[2] / ms);
            coMax = (k.scalarTransport ? sqrt(c.Uf[0] * c.Uf[0] + c.Uf[1] * c.Uf[1] + c.Uf[2] * c.Uf[2])       // scalarTransportQHDFoam.C:88-96
                                       : fabs(Unf)) / __ldg(&fv.hf[f]);
            tauMin = tau;
        }
    }
    if (ADJUST) blockReduceCoQ<kB>(coMax, tauMin, q.sc);
}

// boundary-face evaluation shared by the pre- and post-solve boundary kernels
struct BndCommon { FaceCommon c; FaceGeo ge; double snU[3], snT, d1T, d2T; int flags; int4 v; };

__device__ __forceinline__ double bndSnGrad(int kind, double delta, double vb, double vc, double bv)
{
    return kind == QGD_BC_FIXED_VALUE ? delta * (vb - vc) : (kind == QGD_BC_ZERO_GRADIENT ? 0.0 : bv);
}

__device__ __forceinline__ void bndCommon(const QhdConsts& k, const FaceView& fv, const QhdView& q, int b, BndCommon& o)
{
    const int f = fv.nI + b;
    const int P = fv.own[f];
    const size_t n = q.nCells, nB = fv.nB;
    o.flags = fv.flags[f];
    o.v = fv.vtx[f];
    o.ge = loadGeo(fv, f);
    const double delta = fv.dC[f], hd = fv.halfDist[b];
    double uP[3], uB[3], d1[3], d2[3], dP[3];
    const int kU = q.bcU[b];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uP[j] = q.Q[j * n + P];
        uB[j] = q.UB[j * nB + b];
        o.snU[j] = bndSnGrad(kU, delta, uB[j], uP[j], q.bvU[j * nB + b]);
        ptDiff(q, j, o.v, o.flags, d1[j], d2[j]);
        dP[j] = uP[j] - (uB[j] + o.snU[j] * hd);                    // GaussVolPointBase3D.C:790-793
    }
    const double TP = q.Q[3 * n + P], TB = q.TB[b];
    o.snT = bndSnGrad(q.bcT[b], delta, TB, TP, q.bvT[b]);
    ptDiff(q, 3, o.v, o.flags, o.d1T, o.d2T);
    FaceCommon& c = o.c;
    if (o.flags & FF_NORMAL_ONLY) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) c.G[3 * i + j] = o.ge.gp[i] * o.snU[j];
    } else gradVec(o.ge, o.flags, d1, d2, dP, c.G);
#pragma unroll
    for (int j = 0; j < 3; ++j) { c.Uf[j] = uB[j]; c.Bf[j] = (k.beta * TB) * k.g[j]; }
    c.Tf = TB;
#pragma unroll
    for (int j = 0; j < 3; ++j) c.UgU[j] = c.Uf[0] * c.G[j] + c.Uf[1] * c.G[3 + j] + c.Uf[2] * c.G[6 + j];
    c.phiu = o.ge.Sf[0] * c.Uf[0] + o.ge.Sf[1] * c.Uf[1] + o.ge.Sf[2] * c.Uf[2];
    c.tau = q.tauf[f];
}

__global__ void k_qhd_bnd_pre(QhdConsts k, FaceView fv, QhdView q, int adjust)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double coMax = 0.0, tauMin = DBL_MAX;
    if (b < fv.nB) {
        const int f = fv.nI + b;
        if (fv.bKind[b] == QGD_PATCH_EMPTY) q.F0[f] = 0.0;
        else {
            BndCommon bc;
            bndCommon(k, fv, q, b, bc);
            q.F0[f] = bc.c.phiu - phiwoOf(bc.ge, bc.c);
            if (fv.own[f] < q.nOwned) {
                const double ms = fv.magSf[f];
                const double Unf = bc.c.Uf[0] * (bc.ge.Sf[0] / ms) + bc.c.Uf[1] * (bc.ge.Sf[1] / ms) + bc.c.Uf[2] * (bc.ge.Sf[2] / ms);
                coMax = (k.scalarTransport ? sqrt(bc.c.Uf[0] * bc.c.Uf[0] + bc.c.Uf[1] * bc.c.Uf[1] + bc.c.Uf[2] * bc.c.Uf[2]) : fabs(Unf)) / fv.hf[f];
                tauMin = bc.c.tau;
            }
        }
    }
    if (adjust) blockReduceCoQ<kB>(coMax, tauMin, q.sc);
}

// setDeltaT-QGDQHD.H:41-61 ; QHDCourantNo.H:54
__global__ void k_qhd_dt(StepScalars* sc)
{
    if (sc->adjust) {
        const double coNum = __longlong_as_double((long long)sc->coMaxBits) * sc->dt;
        const double tauMin = __longlong_as_double((long long)sc->tauMinBits);
        sc->coNum = coNum;
        const double maxDeltaTFact = sc->maxCo / (coNum + 1e-15);
        const double deltaTFact = fmin(fmin(maxDeltaTFact, 1.0 + 0.1 * maxDeltaTFact), 1.2);
        double maxDeltaT1 = sc->cTau * tauMin;
        maxDeltaT1 = fmin(sc->maxDeltaT, maxDeltaT1);
        sc->dt = fmin(deltaTFact * sc->dt, maxDeltaT1);
    }
    sc->time += sc->dt;
    sc->coMaxBits = 0ull;
    sc->tauMinBits = dbitsQ(DBL_MAX);
}

// iterate over the faces of cell c in ascending polyMesh order: fn(deviceFace, isNeighbourSide)
template <class F>
__device__ __forceinline__ void forCellFaces(const QhdView& q, int c, F fn)
{
    const int W = q.cfEllW;
    int last = -1;
    for (int j = 0; j < W; ++j) {
        const int e = __ldg(&q.cfEll[(size_t)j * q.nCells + c]);
        last = e;
        if (e >= 0) fn(e >> 1, e & 1);
    }
    if (last >= 0)
        for (int t = __ldg(&q.cfTailOff[c]); t < __ldg(&q.cfTailOff[c + 1]); ++t) {
            const int e = __ldg(&q.cfTailEnc[t]);
            fn(e >> 1, e & 1);
        }
}

// pEqn source + Gauss-linear cell gradient of U
__global__ void __launch_bounds__(kB) k_qhd_cell_pre(QhdConsts k, FaceView fv, QhdView q)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= q.nCells) return;
    const size_t n = q.nCells, nB = fv.nB;
    const double uc[3] = {q.Q[c], q.Q[n + c], q.Q[2 * n + c]};
    double s0 = 0.0, G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    forCellFaces(q, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        double uf[3];
        if (f < fv.nI) {
            const int o = side ? __ldg(&fv.own[f]) : __ldg(&fv.nei[f]);
            const double w = __ldg(&fv.w[f]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double uo = q.Q[j * n + o];
                uf[j] = side ? (w * (uo - uc[j]) + uc[j]) : (w * (uc[j] - uo) + uo);     // linearInterpolate(U)
            }
        } else {
            const int b = f - fv.nI;
            if (__ldg(&fv.bKind[b]) == QGD_PATCH_EMPTY) return;
#pragma unroll
            for (int j = 0; j < 3; ++j) uf[j] = q.UB[j * nB + b];
        }
        s0 += sgn * q.F0[f];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double Si = __ldg(&fv.Sf[(size_t)i * fv.fs + f]);
#pragma unroll
            for (int j = 0; j < 3; ++j) G[3 * i + j] += sgn * (Si * uf[j]);
        }
    });
    const double V = __ldg(&q.V[c]);
#pragma unroll
    for (int t = 0; t < 9; ++t) q.GU[t * n + c] = G[t] / V;
    double src = -(V * (s0 / V));
    if (k.needRef && c == k.refCell) src += __ldg(&q.diag0[c]) * q.Q[4 * n + c];       // fvMatrix::setReference
    q.pcgB[c] = src + __ldg(&q.srcBnd[c]);
}

// fluxes of the U and T equations on internal faces
__global__ void __launch_bounds__(kB) k_qhd_face_post(QhdConsts k, FaceView fv, QhdView q)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fv.nI) return;
    const size_t n = q.nCells, nF = fv.nF;
    const int P = __ldg(&fv.own[f]), N = __ldg(&fv.nei[f]), flags = __ldg(&fv.flags[f]);
    const int4 v = __ldg(&fv.vtx[f]);
    const FaceGeo ge = loadGeo(fv, f);
    double uP[3], uN[3], d1[3], d2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { uP[j] = q.Q[j * n + P]; uN[j] = q.Q[j * n + N]; ptDiff(q, j, v, flags, d1[j], d2[j]); }
    const double TP = q.Q[3 * n + P], TN = q.Q[3 * n + N], pP = q.Q[4 * n + P], pN = q.Q[4 * n + N];
    const double w = __ldg(&fv.w[f]), tau = __ldg(&q.tauf[f]);
    FaceCommon c;
    const bool lsq = (flags & FF_LSQ) != 0;
    if (lsq) lsqGradVecQ(fv, q, f, uP, uN, w, c.G);
    faceCommon(k, ge, flags, uP, uN, TP, TN, w, d1, d2, tau, c, lsq);
    double dT1, dT2, dp1, dp2;
    ptDiff(q, 3, v, flags, dT1, dT2);
    ptDiff(q, 4, v, flags, dp1, dp2);
    double gT[3], gPr[3];
    if (lsq) {
        lsqGradQ(fv, q, f, 3, TP, TN, w, gT);
        lsqGradQ(fv, q, f, 4, pP, pN, w, gPr);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gT[i] = ge.g1[i] * dT1 + ge.g2[i] * dT2 + ge.gp[i] * (TP - TN);
            gPr[i] = ge.g1[i] * dp1 + ge.g2[i] * dp2 + ge.gp[i] * (pP - pN);
        }
    }
    const double up = __ldg(&q.upper[f]);
    const double phi = k.scalarTransport ? c.phiu                                      // scalarTransportQHDFoam.C:110 qgdFlux(phiu,T,Tf)
                                         : q.F0[f] + (up * pN - up * pP);              // QHDpEqn.H:47
    q.phi[f] = phi;
    const double ms = __ldg(&fv.magSf[f]), nd = __ldg(&fv.ndC[f]);
    const double pf = w * (pP - pN) + pN;
    double Wf[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) Wf[j] = tau * (c.UgU[j] + gPr[j] / k.rho0 - c.Bf[j]);  // QHDUEqn.H:37
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double UW = ge.Sf[0] * (c.Uf[0] * Wf[j]) + ge.Sf[1] * (c.Uf[1] * Wf[j]) + ge.Sf[2] * (c.Uf[2] * Wf[j]);   // :39
        const double phiUf = phi * c.Uf[j] - UW;                                       // :41-43
        const double lap = k.implicit ? 0.0 : k.nu * (nd * (uN[j] - uP[j])) * ms;      // :74 fvc::laplacian(muf/rhof,U) (explicit branch)
        double tr = 0.0;                                                               // :76 (muf/rhof Sf) & I(T(grad U))
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double gP_ = q.GU[(size_t)(3 * j + i) * n + P], gN_ = q.GU[(size_t)(3 * j + i) * n + N];
            tr += (k.nu * ge.Sf[i]) * (w * (gP_ - gN_) + gN_);
        }
        q.FU[j * nF + f] = phiUf - lap - tr + ge.Sf[j] * pf / k.rho0;                  // + Gauss-linear grad(p)/rho (:79)
    }
    const double phiTf = phi * c.Tf;                                                   // QHDTEqn.H:65
    const double reg = tau * c.phiu * (c.Uf[0] * gT[0] + c.Uf[1] * gT[1] + c.Uf[2] * gT[2]);   // :66
    const double lapT = k.implicit ? 0.0 : k.Hi * (nd * (TN - TP)) * ms;               // :87 (explicit branch)
    q.FT[f] = phiTf - lapT - reg;
}

__global__ void k_qhd_bnd_post(QhdConsts k, FaceView fv, QhdView q)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    const int f = fv.nI + b;
    const size_t n = q.nCells, nF = fv.nF;
    if (fv.bKind[b] == QGD_PATCH_EMPTY) {
        q.phi[f] = 0.0; q.FT[f] = 0.0; q.FU[f] = 0.0; q.FU[nF + f] = 0.0; q.FU[2 * nF + f] = 0.0;
        return;
    }
    BndCommon bc;
    bndCommon(k, fv, q, b, bc);
    const FaceCommon& c = bc.c;
    const FaceGeo& ge = bc.ge;
    const int P = fv.own[f];
    const double delta = fv.dC[f], hd = fv.halfDist[b], ms = fv.magSf[f];
    const double pP = q.Q[4 * n + P], pb = q.pB[b], TP = q.Q[3 * n + P];
    const double snP = bndSnGrad(q.bcP[b], delta, pb, pP, q.bvP[b]);
    double gT[3], gPr[3];
    if (bc.flags & FF_NORMAL_ONLY) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { gT[i] = ge.gp[i] * bc.snT; gPr[i] = ge.gp[i] * snP; }
    } else {
        double dp1, dp2;
        ptDiff(q, 4, bc.v, bc.flags, dp1, dp2);
        const double dTP = TP - (c.Tf + bc.snT * hd), dpP = pP - (pb + snP * hd);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gT[i] = ge.g1[i] * bc.d1T + ge.g2[i] * bc.d2T + ge.gp[i] * dTP;
            gPr[i] = ge.g1[i] * dp1 + ge.g2[i] * dp2 + ge.gp[i] * dpP;
        }
    }
    const double phi = k.scalarTransport ? c.phiu : q.F0[f] + (q.intC[b] * pP - q.bouC[b]);   // fvMatrix::flux, boundary part
    q.phi[f] = phi;
    double nrm[3] = {ge.Sf[0] / ms, ge.Sf[1] / ms, ge.Sf[2] / ms};
    double Wf[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) Wf[j] = c.tau * (c.UgU[j] + gPr[j] / k.rho0 - c.Bf[j]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double UW = ge.Sf[0] * (c.Uf[0] * Wf[j]) + ge.Sf[1] * (c.Uf[1] * Wf[j]) + ge.Sf[2] * (c.Uf[2] * Wf[j]);
        const double phiUf = phi * c.Uf[j] - UW;
        const double lap = k.implicit ? 0.0 : k.nu * bc.snU[j] * ms;
        // boundary value of T(fvc::grad(U)): gaussGrad::correctBoundaryConditions  [OF-v2312]
        //   gradU_b = gradU_P + n (x) (snGrad(U)_b - n . gradU_P) ; the flux needs (gradU_b)_{j i}
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double nGi = 0.0;
#pragma unroll
            for (int m = 0; m < 3; ++m) nGi += nrm[m] * q.GU[(size_t)(3 * m + i) * n + P];
            const double gb = q.GU[(size_t)(3 * j + i) * n + P] + nrm[j] * (bc.snU[i] - nGi);   // (gradU_b)_{j i}
            tr += (k.nu * ge.Sf[i]) * gb;
        }
        q.FU[j * nF + f] = phiUf - lap - tr + ge.Sf[j] * pb / k.rho0;
    }
    const double reg = c.tau * c.phiu * (c.Uf[0] * gT[0] + c.Uf[1] * gT[1] + c.Uf[2] * gT[2]);
    q.FT[f] = phi * c.Tf - (k.implicit ? 0.0 : k.Hi * bc.snT * ms) - reg;
}

// implicit branch: diagonals and right-hand sides of the U and T systems (QHDUEqn.H:48-64, QHDTEqn.H:71-79)
__global__ void __launch_bounds__(kB) k_qhd_cell_sys(QhdConsts k, FaceView fv, QhdView q)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= q.nCells) return;
    const size_t n = q.nCells, nF = fv.nF;
    double su[3] = {0, 0, 0}, sT = 0.0, sPhi = 0.0;
    forCellFaces(q, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        su[0] += sgn * q.FU[f]; su[1] += sgn * q.FU[nF + f]; su[2] += sgn * q.FU[2 * nF + f];
        sT += sgn * q.FT[f];
        sPhi += sgn * q.phi[f];
    });
    const double V = __ldg(&q.V[c]);
    const double rDeltaT = 1.0 / q.sc->dt;
    const double T = q.Q[3 * n + c];
    if (k.scalarTransport) {     // scalarTransportQHDFoam.C:116-124: only T; - fvc::Sp(fvc::div(phiu),T) goes to the source (phi == phiu here)
        q.diagT[c] = rDeltaT * V + __ldg(&q.sumAT[c]);
        q.bT[c] = rDeltaT * T * V - V * (sT / V) + V * ((sPhi / V) * T) + __ldg(&q.srcBT[c]);
        return;
    }
    q.diagU[c] = rDeltaT * V + __ldg(&q.sumAU[c]);
    q.diagT[c] = rDeltaT * V + __ldg(&q.sumAT[c]);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        q.bU[j * n + c] = rDeltaT * q.Q[j * n + c] * V - V * (su[j] / V) + V * ((k.beta * T) * k.g[j]) + __ldg(&q.srcBU[j * n + c]);
    q.bT[c] = rDeltaT * T * V - V * (sT / V) + __ldg(&q.srcBT[c]);
}

__global__ void k_qhd_pshift(QhdView q)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < q.nCells) q.Q[4 * (size_t)q.nCells + c] += *q.shift;
}

__global__ void k_qhd_shift(QhdConsts k, QhdView q)
{
    // decomposed runs: the rank that owns the reference cell computes the shift (refCell = -1 elsewhere), the others receive it by all-reduce
    *q.shift = (k.needRef && k.refCell >= 0) ? (k.refValue - q.Q[4 * (size_t)q.nCells + k.refCell]) : 0.0;
}

__global__ void __launch_bounds__(kB) k_qhd_cell_update(QhdConsts k, FaceView fv, QhdView q)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= q.nOwned) return;
    const size_t n = q.nCells, nF = fv.nF;
    double su[3] = {0, 0, 0}, sT = 0.0;
    forCellFaces(q, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        su[0] += sgn * q.FU[f]; su[1] += sgn * q.FU[nF + f]; su[2] += sgn * q.FU[2 * nF + f];
        sT += sgn * q.FT[f];
    });
    const double V = __ldg(&q.V[c]);
    const double rDeltaT = 1.0 / q.sc->dt;
    const double diag = rDeltaT * V;
    const double T = q.Q[3 * n + c];
#pragma unroll
    for (int j = 0; j < 3; ++j) {                                                      // QHDUEqn.H:68-84
        const double src = rDeltaT * q.Q[j * n + c] * V - V * (su[j] / V) + V * ((k.beta * T) * k.g[j]);
        q.Q[j * n + c] = src / diag;
    }
    q.Q[3 * n + c] = (rDeltaT * T * V - V * (sT / V)) / diag;                          // QHDTEqn.H:83-91
    q.Q[4 * n + c] += *q.shift;                                                        // QHDFoam.C:123-131
}

} // namespace

} // namespace qgd

// ============================================================================ solver object + C ABI
using namespace qgd;

struct qgd_qhd_solver {
    qgd_mesh* mesh = nullptr;
    std::unique_ptr<qgd_fvsc> fvsc;
    qgd_qhdfoam_desc desc{};
    std::string model, precondName;
    QhdConsts k{};
    int precond = 2;
    DevBuf<double> Q, P, tauf, upper, F0, FU, FT, phi, GU, UB, TB, pB, bvU, bvT, bvP, intC, bouC, srcBnd, diag0, shift, stage;
    DevBuf<double> sumAU, sumAT, srcBU, srcBT, diagU, diagT, bU, bT;      // implicit branch
    PcgMatrix AU, AT;
    int diffPrecond = 2;
    std::vector<double> hbvU, hbvT;
    DevBuf<int> bcU, bcT, bcP;
    DevBuf<StepScalars> sc;
    PcgMatrix A;
    std::vector<int> hbcU, hbcT, hbcP;           // per boundary face
    std::vector<double> hbvP, tauCell, tauBnd;
    long long launches = 0;
    bool bcsSet = false, fieldsSet = false;
    // decomposed run (extended sub-mesh): vertex-ring lists for the state, face-neighbour lists for the PCG search direction and
    // the Gauss gradients; the pressure equation is solved by the stepwise PCG with NCCL exchange / all-reduce between its phases
    HaloLists halo, haloFace;
    StepwisePcg sw;
    DevBuf<PcgResult> swOut;
    bool fixesPLocal = false;
    bool multi() const { return mesh->h.nOwned != mesh->h.nCells; }
    QhdView view()
    {
        const qgd_mesh& m = *mesh;
        QhdView q;
        q.nCells = m.h.nCells; q.nPoints = m.h.nPoints; q.nPatchPoints = (int)m.h.patchPoints.size();
        q.nOwned = m.h.nOwned;
        q.Q = Q.p; q.P = P.p;
        q.pcEllW = m.pcEllW; q.pcEll = m.pcEll.p; q.pcEllWt = m.pcEllWt.p; q.pcCount = m.pcCount.p;
        q.pcTailOff = m.pcTailOff.p; q.pcTailCell = m.pcTailCell.p; q.pcTailW = m.pcTailW.p;
        q.patchPoints = m.patchPoints.p; q.ppOff = m.ppOff.p; q.ppFace = m.ppFace.p; q.ppW = m.ppW.p;
        q.cfEllW = m.cfEllW; q.cfEll = m.cfEll.p; q.cfTailOff = m.cfTailOff.p; q.cfTailEnc = m.cfTailEnc.p; q.V = m.V.p;
        q.tauf = tauf.p; q.upper = upper.p; q.F0 = F0.p; q.FU = FU.p; q.FT = FT.p; q.phi = phi.p; q.GU = GU.p;
        q.UB = UB.p; q.TB = TB.p; q.pB = pB.p; q.bcU = bcU.p; q.bcT = bcT.p; q.bcP = bcP.p;
        q.bvU = bvU.p; q.bvT = bvT.p; q.bvP = bvP.p; q.intC = intC.p; q.bouC = bouC.p; q.srcBnd = srcBnd.p; q.diag0 = diag0.p;
        q.pcgB = A.b.p; q.shift = shift.p; q.sc = sc.p;
        q.sumAU = sumAU.p; q.sumAT = sumAT.p; q.srcBU = srcBU.p; q.srcBT = srcBT.p; q.diagU = diagU.p; q.diagT = diagT.p; q.bU = bU.p; q.bT = bT.p;
        return q;
    }
};

namespace {

template <int K0, int K> void launchPoints(cudaStream_t st, const QhdView& q)
{
    const int g = nblk(q.nPoints);
    if (q.pcEllW == 4) k_qhd_points<4, K0, K><<<g, kB, 0, st>>>(q);
    else if (q.pcEllW == 6) k_qhd_points<6, K0, K><<<g, kB, 0, st>>>(q);
    else k_qhd_points<8, K0, K><<<g, kB, 0, st>>>(q);
}

void qhdRunSteps(qgd_qhd_solver* s, int nSteps)
{
    cudaStream_t st = runtimeStream();
    const FaceView fv = s->fvsc->view();
    const QhdView q = s->view();
    const QhdConsts& k = s->k;
    const bool pts = !s->fvsc->reduced && s->mesh->h.nD > 1;
    const bool adjust = s->desc.adjust_time_step != 0;
    const int nB = fv.nB;
    const bool multi = s->multi();
    const size_t nCs = q.nCells;
    PcgHooks hooks;
    if (multi) {
        hooks.exchange = [s](double* vec, cudaStream_t cs) { s->launches += commExchange(s->haloFace, vec, 0, 1, cs); };
        hooks.allreduceSum = [](double* dev, int count, cudaStream_t cs) { commAllReduce(dev, count, COMM_SUM, cs); };
    }
    for (int i = 0; i < nSteps; ++i) {
        int n = 0;
        if (nB) { k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 0, 0); ++n; }
        if (pts) {
            launchPoints<0, 4>(st, q); ++n;
            if (q.nPatchPoints) { k_qhd_patch_points<<<nblk(q.nPatchPoints), kB, 0, st>>>(q, nB, 0); ++n; }
        }
        if (fv.nI) {
            if (adjust) k_qhd_face_pre<true><<<nblk(fv.nI), kB, 0, st>>>(k, fv, q);
            else k_qhd_face_pre<false><<<nblk(fv.nI), kB, 0, st>>>(k, fv, q);
            ++n;
        }
        if (nB) { k_qhd_bnd_pre<<<nblk(nB), kB, 0, st>>>(k, fv, q, adjust ? 1 : 0); ++n; }
        if (multi && adjust) {       // gMax / gMin of QHDCourantNo.H:54, setDeltaT-QGDQHD.H:46 (bit patterns of non-negative doubles)
            commAllReduce(reinterpret_cast<double*>(&q.sc->coMaxBits), 1, COMM_MAX, st);
            commAllReduce(reinterpret_cast<double*>(&q.sc->tauMinBits), 1, COMM_MIN, st);
        }
        k_qhd_dt<<<1, 1, 0, st>>>(q.sc); ++n;
        if (!k.scalarTransport) {                                                        // scalarTransportQHDFoam has no pressure equation
        if (nB) { k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 1, 0); ++n; }           // QHDpEqn.H:35
        k_qhd_cell_pre<<<nblk(q.nCells), kB, 0, st>>>(k, fv, q); ++n;
        // QHDpEqn.H:45 — x is the p slice of the state
        s->A.xExternal = q.Q + 4 * (size_t)q.nCells;
        if (multi) {
            // fvc::grad(U) of the face-neighbour halo cells (their Gauss sums are incomplete locally), then the decomposed solve
            n += commExchange(s->haloFace, q.GU, nCs, 9, st);
            PcgResult r;
            n += s->sw.solve(s->A, s->A.b.p, s->A.xExternal, s->desc.p_tolerance, s->desc.p_rel_tol, s->desc.p_max_iter, s->precond, st, &hooks, &r);
            QGD_CUDA(cudaMemcpyAsync(s->A.out.p, &r, sizeof(r), cudaMemcpyHostToDevice, st));
            QGD_CUDA(cudaStreamSynchronize(st));
            n += commExchange(s->halo, s->A.xExternal, nCs, 1, st);          // p on the whole vertex ring (point interpolation of p)
        } else
        n += s->A.solve(s->desc.p_tolerance, s->desc.p_rel_tol, s->desc.p_max_iter, st);
        if (nB) { k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 1, 0); ++n; }           // psi.correctBoundaryConditions()
        if (pts) {
            launchPoints<4, 1>(st, q); ++n;
            if (q.nPatchPoints) { k_qhd_patch_points<<<nblk(q.nPatchPoints), kB, 0, st>>>(q, nB, 1); ++n; }
        }
        }
        if (fv.nI) { k_qhd_face_post<<<nblk(fv.nI), kB, 0, st>>>(k, fv, q); ++n; }
        if (nB) { k_qhd_bnd_post<<<nblk(nB), kB, 0, st>>>(k, fv, q); ++n; }
        k_qhd_shift<<<1, 1, 0, st>>>(k, q); ++n;
        if (multi) commAllReduce(q.shift, 1, COMM_SUM, st);
        if (k.implicit) {
            const size_t nC = q.nCells;
            const double tol = s->desc.diff_tolerance, rel = s->desc.diff_rel_tol;
            const int maxIter = s->desc.diff_max_iter > 0 ? s->desc.diff_max_iter : 1000;
            k_qhd_cell_sys<<<nblk(q.nCells), kB, 0, st>>>(k, fv, q); ++n;
            if (!k.scalarTransport) {
            s->AU.refresh(nullptr, q.diagU, st);
            for (int j = 0; j < 3; ++j) {                           // QHDUEqn.H:48-64, segregated components
                s->AU.bExternal = q.bU + j * nC;
                s->AU.xExternal = q.Q + j * nC;
                n += 1 + s->AU.solve(tol, rel, maxIter, st);
            }
            }
            s->AT.refresh(nullptr, q.diagT, st);
            s->AT.bExternal = q.bT;
            s->AT.xExternal = q.Q + 3 * nC;
            n += 2 + s->AT.solve(tol, rel, maxIter, st);               // QHDTEqn.H:71-79
            k_qhd_pshift<<<nblk(q.nCells), kB, 0, st>>>(q); ++n;
        } else if (!k.scalarTransport) {                             // scalarTransportQHDFoam.C:114: nothing is solved without implicitDiffusion
            k_qhd_cell_update<<<nblk(q.nOwned), kB, 0, st>>>(k, fv, q); ++n;
        }
        if (multi) n += commExchange(s->halo, q.Q, nCs, 5, st);              // U, T, p of the halo cells for the next step
        if (nB) {
            k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 0, 0);                        // U, T correctBoundaryConditions
            k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 1, 1);                        // p_b += shift
            n += 2;
        }
        s->launches += n;
    }
    QGD_CUDA(cudaGetLastError());
}

} // namespace

extern "C" {

int qgd_qhdfoam_create(qgd_mesh* mesh, const qgd_qhdfoam_desc* d, qgd_qhd_solver** out)
{
    return guarded([&] {
        requireInit();
        if (!mesh || !d || !out || !d->fvsc_scheme || !d->qgd_coeffs_model) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_create: null argument");
        const std::string model = d->qgd_coeffs_model;
        if (!isCoeffsModel(model))                   // QGDCoeffs.C:70-79
            throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown QGD coeffs evaluation approach type " + model +
                                                   "\n\nValid model types are:\n" + coeffsModelToc());
        if (model != "constTau" && model != "H2bynuQHD" && model != "HbyUQHD" && model != "T0byGr")
            throw Error(QGD_ERR_UNSUPPORTED, "QGDCoeffs model " + model + " is not a QHD model available on the device (constTau, H2bynuQHD, HbyUQHD, T0byGr)");
        if (!mesh->h.wedgePts.empty())               // their vertex constraint (pointConstraints) lives in the QGDFoam step only
            throw Error(QGD_ERR_UNSUPPORTED, "wedge / symmetryPlane patches are not available in QHDFoam on the device");
        auto precondOf = [](const char* name) {
            const std::string pc = name ? name : "DIC";
            if (pc == "DIC") return 2;
            if (pc == "diagonal") return 1;
            if (pc == "none") return 0;
            throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown symmetric matrix preconditioner " + pc + "\n\nValid symmetric matrix preconditioners are:\n3\n(\nDIC\ndiagonal\nnone\n)\n");
        };
        const std::string pc = d->p_preconditioner ? d->p_preconditioner : "DIC";
        const int precond = precondOf(d->p_preconditioner);
        const int diffPrecond = d->implicit_diffusion ? precondOf(d->diff_preconditioner) : 2;
        for (int pk : mesh->h.patchKind)
            if (pk == QGD_PATCH_PROCESSOR) throw Error(QGD_ERR_UNSUPPORTED, "QHDFoam: processor patches (multi-GPU) are not available yet");
        const bool sub = mesh->h.nOwned != mesh->h.nCells;
        if (sub) {          // decomposed run on an extended sub-mesh (qgd_qhdfoam_set_halo): explicit branch, PCG + diagonal | none
            if (d->implicit_diffusion || d->scalar_transport)
                throw Error(QGD_ERR_UNSUPPORTED, "QHDFoam on extended sub-meshes (multi-GPU): implicitDiffusion / scalarTransportQHDFoam are not available yet");
            if (precond == 2 && mesh->h.pcgBlock.empty())
                throw Error(QGD_ERR_UNSUPPORTED, "QHDFoam on extended sub-meshes (multi-GPU): DIC is local to a block in a decomposed run - set DIC "
                                                 "blocks first (qgd_mesh_make_pcg_blocks) or use diagonal | none");
        }
        if (!(d->delta_t > 0.0) || !(d->rho0 > 0.0) || !(d->Pr > 0.0)) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_create: deltaT, rho and Pr must be positive");
        // decomposed runs: the local id of the global pRefCell on the rank that owns it, -1 on every other rank
        if (d->p_ref_cell >= mesh->h.nOwned || (d->p_ref_cell < 0 && !(sub && d->p_ref_cell == -1)))
            throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_create: pRefCell out of range");
        std::unique_ptr<qgd_qhd_solver> s(new qgd_qhd_solver());
        s->mesh = mesh; s->desc = *d; s->model = model; s->precondName = pc; s->precond = precond;
        s->desc.fvsc_scheme = nullptr; s->desc.qgd_coeffs_model = nullptr; s->desc.p_preconditioner = nullptr; s->desc.diff_preconditioner = nullptr;
        s->diffPrecond = diffPrecond;
        s->fvsc.reset(new qgd_fvsc());
        fvscBuild(*s->fvsc, mesh, d->fvsc_scheme);
        QhdConsts& k = s->k;
        k.rho0 = d->rho0; k.nu = d->mu / d->rho0; k.Hi = (d->mu / d->Pr) / d->rho0; k.beta = d->beta;
        for (int j = 0; j < 3; ++j) k.g[j] = d->g[j];
        k.needRef = 0; k.refCell = d->p_ref_cell; k.refValue = d->p_ref_value; k.implicit = d->implicit_diffusion ? 1 : 0;
        k.scalarTransport = d->scalar_transport ? 1 : 0;
        const HostMesh& h = mesh->h;
        cudaStream_t st = runtimeStream();
        s->Q.alloc(5 * (size_t)h.nCells); s->P.alloc(5 * (size_t)h.nPoints); s->P.zero(st);
        s->F0.alloc(h.nFaces); s->FU.alloc(3 * (size_t)h.nFaces); s->FT.alloc(h.nFaces); s->phi.alloc(h.nFaces); s->GU.alloc(9 * (size_t)h.nCells);
        s->F0.zero(st); s->FU.zero(st); s->FT.zero(st); s->phi.zero(st);
        s->UB.alloc(3 * (size_t)h.nBnd + 1); s->TB.alloc(h.nBnd + 1); s->pB.alloc(h.nBnd + 1);
        s->UB.zero(st); s->TB.zero(st); s->pB.zero(st);
        s->shift.alloc(1); s->shift.zero(st);
        StepScalars sc{};
        sc.dt = d->delta_t; sc.time = 0.0; sc.coNum = -1.0; sc.coMaxBits = 0ull;
        const double big = DBL_MAX;
        std::memcpy(&sc.tauMinBits, &big, sizeof(double));
        sc.maxCo = d->max_co; sc.maxDeltaT = d->max_delta_t; sc.cTau = d->c_tau; sc.adjust = d->adjust_time_step;
        s->sc.upload(std::vector<StepScalars>(1, sc), st);
        *out = s.release();
    });
}

int qgd_qhdfoam_destroy(qgd_qhd_solver* s) { return guarded([&] { delete s; }); }

int qgd_qhdfoam_set_bcs(qgd_qhd_solver* s, const int* bc_U, const int* bc_T, const int* bc_p, const double* val_U,
                        const double* val_T, const double* val_p)
{
    return guarded([&] {
        requireInit();
        if (!s || !bc_U || !bc_T || !bc_p) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_set_bcs: null argument");
        const HostMesh& h = s->mesh->h;
        cudaStream_t st = runtimeStream();
        const int nB = h.nBnd;
        s->hbcU.assign(nB, 1); s->hbcT.assign(nB, 1); s->hbcP.assign(nB, 1);
        bool fixesP = false;
        auto norm = [&](int kind, const char* what) {
            if (kind == QGD_BC_QGD_FLUX || kind == QGD_BC_QHD_FLUX) kind = QGD_BC_FIXED_GRADIENT;
            if (kind != QGD_BC_FIXED_VALUE && kind != QGD_BC_ZERO_GRADIENT && kind != QGD_BC_FIXED_GRADIENT)
                throw Error(QGD_ERR_UNSUPPORTED, std::string(what) + " boundary condition outside the device-native set (fixedValue, zeroGradient, fixedGradient, qhdFlux)");
            return kind;
        };
        for (int b = 0; b < nB; ++b) {
            const int pi = h.bfacePatch[b];
            if (h.patchKind[pi] == QGD_PATCH_EMPTY) continue;
            s->hbcU[b] = norm(bc_U[pi], "U"); s->hbcT[b] = norm(bc_T[pi], "T"); s->hbcP[b] = norm(bc_p[pi], "p");
            if (s->hbcP[b] == QGD_BC_FIXED_VALUE) fixesP = true;
            if ((s->hbcU[b] != QGD_BC_ZERO_GRADIENT && !val_U) || (s->hbcT[b] != QGD_BC_ZERO_GRADIENT && !val_T) ||
                (s->hbcP[b] != QGD_BC_ZERO_GRADIENT && !val_p))
                throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_set_bcs: fixedValue / fixedGradient patch without values");
        }
        s->fixesPLocal = fixesP;
        s->k.needRef = (fixesP || s->k.scalarTransport) ? 0 : 1;        // p.needReference(); scalarTransportQHDFoam never touches p
        s->bcU.upload(s->hbcU.empty() ? std::vector<int>(1, 1) : s->hbcU, st);
        s->bcT.upload(s->hbcT.empty() ? std::vector<int>(1, 1) : s->hbcT, st);
        s->bcP.upload(s->hbcP.empty() ? std::vector<int>(1, 1) : s->hbcP, st);
        std::vector<double> u(3 * (size_t)nB + 1, 0.0), t(nB + 1, 0.0), p(nB + 1, 0.0);
        for (int b = 0; b < nB; ++b) {
            if (val_U) for (int j = 0; j < 3; ++j) u[(size_t)j * nB + b] = val_U[3 * (size_t)b + j];
            if (val_T) t[b] = val_T[b];
            if (val_p) p[b] = val_p[b];
        }
        s->hbvP.assign(p.begin(), p.begin() + nB);
        s->hbvT.assign(t.begin(), t.begin() + nB);
        s->hbvU.assign(val_U ? val_U : u.data(), (val_U ? val_U : u.data()) + 3 * (size_t)nB);     // AoS as passed in
        s->bvU.upload(u, st); s->bvT.upload(t, st); s->bvP.upload(p, st);
        s->bcsSet = true;
    });
}

int qgd_qhdfoam_init_fields(qgd_qhd_solver* s, const double* U, const double* T, const double* p, const double* alphaQGD)
{
    return guarded([&] {
        requireInit();
        if (!s || !U || !T || !p) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_init_fields: null argument");
        if (!s->bcsSet) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_init_fields: call qgd_qhdfoam_set_bcs first");
        const HostMesh& h = s->mesh->h;
        cudaStream_t st = runtimeStream();
        const int nC = h.nCells, nF = h.nFaces, nI = h.nInternal, nB = h.nBnd;
        const qgd_qhdfoam_desc& d = s->desc;
        const bool multi = s->multi();
        if (multi && !s->halo.active()) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_init_fields: extended sub-mesh without exchange lists (call qgd_qhdfoam_set_halo first)");
        if (multi) {        // p.needReference() is a global property: a rank may hold no face of the fixedValue patch
            double flag = s->fixesPLocal ? 1.0 : 0.0;
            QGD_CUDA(cudaMemcpyAsync(s->shift.p, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
            commAllReduce(s->shift.p, 1, COMM_MAX, st);
            QGD_CUDA(cudaMemcpyAsync(&flag, s->shift.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            QGD_CUDA(cudaStreamSynchronize(st));
            s->k.needRef = flag > 0.5 ? 0 : 1;
            s->shift.zero(st);
        }
        // ---- state (SoA)
        {
            std::vector<double> q(5 * (size_t)nC);
            for (int c = 0; c < nC; ++c) {
                for (int j = 0; j < 3; ++j) q[(size_t)j * nC + c] = U[3 * (size_t)c + j];
                q[3 * (size_t)nC + c] = T[c];
                q[4 * (size_t)nC + c] = p[c];
            }
            QGD_CUDA(cudaMemcpyAsync(s->Q.p, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            QGD_CUDA(cudaStreamSynchronize(st));
        }
        // ---- QGDCoeffs (QHD family): tauQGD on cells and patches, tauQGDf = linearInterpolate(tauQGD); constant in time
        auto tauOf = [&](double a, double hq) {
            if (s->model == "constTau") return d.Tau;                                   // constTau.C:71-73
            if (s->model == "H2bynuQHD") { const double nu = d.mu / d.rho0; return a * (hq * hq) / nu; }   // H2bynuQHD.C:80-81
            if (s->model == "HbyUQHD") return a * hq / d.UQHD;                          // HbyUQHD.C:82
            return d.T0 / d.Gr;                                                         // T0byGr.C:86
        };
        s->tauCell.resize(nC); s->tauBnd.assign(nB, 0.0);
        for (int c = 0; c < nC; ++c) s->tauCell[c] = tauOf(alphaQGD ? alphaQGD[c] : 0.5, h.hQGD[c]);
        if (multi) {        // hQGD of a halo cell is incomplete on the sub-mesh (its outer faces are cut): take the owners' tauQGD
            DevBuf<double> t;
            t.upload(s->tauCell, st);
            commExchange(s->halo, t.p, (size_t)nC, 1, st);
            QGD_CUDA(cudaMemcpyAsync(s->tauCell.data(), t.p, (size_t)nC * sizeof(double), cudaMemcpyDeviceToHost, st));
            QGD_CUDA(cudaStreamSynchronize(st));
        }
        for (int b = 0; b < nB; ++b) {
            if (h.patchKind[h.bfacePatch[b]] == QGD_PATCH_EMPTY) continue;
            const int P = h.owner[nI + b];
            s->tauBnd[b] = tauOf(alphaQGD ? alphaQGD[P] : 0.5, h.hQGDf[nI + b]);        // alphaQGD zeroGradient, hQGD_b = hQGDf_b
        }
        std::vector<double> tauf(nF, 0.0);
        for (int f = 0; f < nI; ++f) {
            const double tP = s->tauCell[h.owner[f]], tN = s->tauCell[h.neighbour[f]];
            tauf[f] = h.w[f] * (tP - tN) + tN;
        }
        for (int b = 0; b < nB; ++b) tauf[nI + b] = s->tauBnd[b];
        // ---- pEqn matrix  -fvm::laplacian(tauQGDf/rhof, p)  (QHDpEqn.H:40) and its boundary coefficients
        std::vector<double> upper(std::max(nI, 1), 0.0), diag(nC, 0.0), intC(nB + 1, 0.0), bouC(nB + 1, 0.0), srcBnd(nC, 0.0);
        for (int f = 0; f < nI; ++f) {
            const double up = h.ndC[f] * ((tauf[f] / d.rho0) * h.magSf[f]);
            upper[f] = -up;
            diag[h.owner[f]] += up; diag[h.neighbour[f]] += up;
        }
        for (int b = 0; b < nB; ++b) {
            if (h.patchKind[h.bfacePatch[b]] == QGD_PATCH_EMPTY) continue;
            const int f = nI + b;
            const double gS = (tauf[f] / d.rho0) * h.magSf[f];
            if (s->hbcP[b] == QGD_BC_FIXED_VALUE) { intC[b] = gS * h.ndC[f]; bouC[b] = gS * (h.ndC[f] * s->hbvP[b]); }
            else if (s->hbcP[b] == QGD_BC_FIXED_GRADIENT) bouC[b] = gS * s->hbvP[b];
        }
        std::vector<double> diagT(diag);
        if (s->k.needRef && s->k.refCell >= 0) diagT[s->k.refCell] += diagT[s->k.refCell];   // fvMatrix::setReference
        for (int b = 0; b < nB; ++b) { const int P = h.owner[nI + b]; diagT[P] += intC[b]; srcBnd[P] += bouC[b]; }
        for (int c = h.nOwned; c < nC; ++c) diagT[c] = 1.0;                             // halo rows are never solved; keep 1/diag finite
        s->A.build(h, diagT.data(), upper.data(), s->precond, st);
        if (multi) s->sw.alloc(s->A, h.nOwned);
        // device face order for the per-face arrays
        const std::vector<int>& perm = s->mesh->facePerm;
        std::vector<double> taufDev(nF), upperDev(std::max(nI, 1), 0.0);
        for (int f = 0; f < nF; ++f) taufDev[f] = tauf[perm[f]];
        for (int f = 0; f < nI; ++f) upperDev[f] = upper[perm[f]];
        s->tauf.upload(taufDev, st); s->upper.upload(upperDev, st);
        s->intC.upload(intC, st); s->bouC.upload(bouC, st); s->srcBnd.upload(srcBnd, st); s->diag0.upload(diag, st);
        if (s->k.implicit) {
            // U and T systems: -laplacian(muf/rhof, .) and -laplacian(Hif, .) are time-constant; only V/deltaT changes
            std::vector<double> upU(std::max(nI, 1), 0.0), upT(std::max(nI, 1), 0.0), sAU(nC, 0.0), sAT(nC, 0.0), sBU(3 * (size_t)nC, 0.0), sBT(nC, 0.0), d0(nC, 1.0);
            const double nu = d.mu / d.rho0, Hi = (d.mu / d.Pr) / d.rho0;
            for (int f = 0; f < nI; ++f) {
                const double aU = h.ndC[f] * (nu * h.magSf[f]), aT = h.ndC[f] * (Hi * h.magSf[f]);
                upU[f] = -aU; upT[f] = -aT;
                sAU[h.owner[f]] += aU; sAU[h.neighbour[f]] += aU; sAT[h.owner[f]] += aT; sAT[h.neighbour[f]] += aT;
            }
            for (int b = 0; b < nB; ++b) {
                if (h.patchKind[h.bfacePatch[b]] == QGD_PATCH_EMPTY) continue;
                const int f = nI + b, P = h.owner[f];
                const double gU = nu * h.magSf[f], gT = Hi * h.magSf[f];
                if (s->hbcU[b] == QGD_BC_FIXED_VALUE) { sAU[P] += gU * h.ndC[f]; for (int j = 0; j < 3; ++j) sBU[(size_t)j * nC + P] += gU * (h.ndC[f] * s->hbvU[3 * (size_t)b + j]); }
                else if (s->hbcU[b] == QGD_BC_FIXED_GRADIENT) for (int j = 0; j < 3; ++j) sBU[(size_t)j * nC + P] += gU * s->hbvU[3 * (size_t)b + j];
                if (s->hbcT[b] == QGD_BC_FIXED_VALUE) { sAT[P] += gT * h.ndC[f]; sBT[P] += gT * (h.ndC[f] * s->hbvT[b]); }
                else if (s->hbcT[b] == QGD_BC_FIXED_GRADIENT) sBT[P] += gT * s->hbvT[b];
            }
            s->AU.build(h, d0.data(), upU.data(), s->diffPrecond, st);
            s->AT.build(h, d0.data(), upT.data(), s->diffPrecond, st);
            s->sumAU.upload(sAU, st); s->sumAT.upload(sAT, st); s->srcBU.upload(sBU, st); s->srcBT.upload(sBT, st);
            s->diagU.alloc(nC); s->diagT.alloc(nC); s->bU.alloc(3 * (size_t)nC); s->bT.alloc(nC);
        }
        // boundary values of the fields as read
        const FaceView fv = s->fvsc->view();
        const QhdView q = s->view();
        if (multi) commExchange(s->halo, s->Q.p, (size_t)nC, 5, st);     // the caller's halo values may be placeholders
        if (nB) {
            k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 0, 0);
            k_qhd_bnd_eval<<<nblk(nB), kB, 0, st>>>(fv, q, 1, 0);
        }
        QGD_CUDA(cudaGetLastError());
        QGD_CUDA(cudaStreamSynchronize(st));
        s->fieldsSet = true;
    });
}

int qgd_qhdfoam_set_halo(qgd_qhd_solver* s, int nn, const int* nbr_rank, const int* send_off, const int* send_cells, const int* recv_off,
                         const int* recv_cells, int nn_face, const int* nbr_rank_face, const int* fsend_off, const int* fsend_cells,
                         const int* frecv_off, const int* frecv_cells)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_set_halo: null solver");
        if (s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_set_halo: call before qgd_qhdfoam_init_fields");
        const HostMesh& h = s->mesh->h;
        cudaStream_t st = runtimeStream();
        s->halo.set(nn, nbr_rank, send_off, send_cells, recv_off, recv_cells, h.nCells, h.nOwned, 9, st);
        s->haloFace.set(nn_face, nbr_rank_face, fsend_off, fsend_cells, frecv_off, frecv_cells, h.nCells, h.nOwned, 9, st);
    });
}

int qgd_qhdfoam_step(qgd_qhd_solver* s, int n_steps)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_step: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_step: call qgd_qhdfoam_init_fields first");
        qhdRunSteps(s, n_steps);
    });
}

int qgd_qhdfoam_get(qgd_qhd_solver* s, int field, double* cells, double* bnd)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_get: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_get: no fields yet");
        const HostMesh& h = s->mesh->h;
        cudaStream_t st = runtimeStream();
        const size_t n = h.nCells, nB = h.nBnd;
        if (field < 0 || field > 3) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_get: unknown field id");
        if (field == 3) {
            if (cells) std::copy(s->tauCell.begin(), s->tauCell.end(), cells);
            if (bnd) std::copy(s->tauBnd.begin(), s->tauBnd.end(), bnd);
            return;
        }
        if (cells) {
            if (field == 0) {
                std::vector<double> t(3 * n);
                QGD_CUDA(cudaMemcpyAsync(t.data(), s->Q.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
                QGD_CUDA(cudaStreamSynchronize(st));
                for (size_t c = 0; c < n; ++c) for (int j = 0; j < 3; ++j) cells[3 * c + j] = t[j * n + c];
            } else {
                QGD_CUDA(cudaMemcpyAsync(cells, s->Q.p + (field == 1 ? 3 : 4) * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
                QGD_CUDA(cudaStreamSynchronize(st));
            }
        }
        if (bnd && nB) {
            if (field == 0) {
                std::vector<double> t(3 * nB);
                QGD_CUDA(cudaMemcpyAsync(t.data(), s->UB.p, 3 * nB * sizeof(double), cudaMemcpyDeviceToHost, st));
                QGD_CUDA(cudaStreamSynchronize(st));
                for (size_t b = 0; b < nB; ++b) for (int j = 0; j < 3; ++j) bnd[3 * b + j] = t[j * nB + b];
            } else {
                QGD_CUDA(cudaMemcpyAsync(bnd, field == 1 ? s->TB.p : s->pB.p, nB * sizeof(double), cudaMemcpyDeviceToHost, st));
                QGD_CUDA(cudaStreamSynchronize(st));
            }
        }
    });
}

int qgd_qhdfoam_get_flux(qgd_qhd_solver* s, double* phi)
{
    return guarded([&] {
        requireInit();
        if (!s || !phi) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_get_flux: null argument");
        const size_t nF = s->mesh->h.nFaces;
        const std::vector<int>& perm = s->mesh->facePerm;
        std::vector<double> t(nF);
        QGD_CUDA(cudaMemcpyAsync(t.data(), s->phi.p, nF * sizeof(double), cudaMemcpyDeviceToHost, runtimeStream()));
        QGD_CUDA(cudaStreamSynchronize(runtimeStream()));
        for (size_t f = 0; f < nF; ++f) phi[perm[f]] = t[f];
    });
}

int qgd_qhdfoam_get_scalars(qgd_qhd_solver* s, double* delta_t, double* courant, double* time)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_get_scalars: null solver");
        StepScalars sc;
        QGD_CUDA(cudaMemcpyAsync(&sc, s->sc.p, sizeof(sc), cudaMemcpyDeviceToHost, runtimeStream()));
        QGD_CUDA(cudaStreamSynchronize(runtimeStream()));
        if (delta_t) *delta_t = sc.dt;
        if (courant) *courant = sc.coNum;
        if (time) *time = sc.time;
    });
}

int qgd_qhdfoam_solver_info(qgd_qhd_solver* s, int* iters, double* initial_residual, double* final_residual)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qhdfoam_solver_info: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qhdfoam_solver_info: no solve yet");
        PcgResult r;
        QGD_CUDA(cudaMemcpyAsync(&r, s->A.out.p, sizeof(r), cudaMemcpyDeviceToHost, runtimeStream()));
        QGD_CUDA(cudaStreamSynchronize(runtimeStream()));
        if (iters) *iters = r.iters;
        if (initial_residual) *initial_residual = r.res0;
        if (final_residual) *final_residual = r.res;
    });
}

long long qgd_qhdfoam_launch_count(qgd_qhd_solver* s) { return s ? s->launches : 0; }

} // extern "C"
