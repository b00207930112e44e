// Hand-written FP64 CUDA for sm_100a: fvsc face-centre derivative operators and the fused QGDFoam step.
//
// Step structure (DESIGN.md "Kernels"):
//   k_points          cell -> point inverse-distance gather of (rho,U,e,p)        [volPointInterpolation, OF]
//   k_patch_points    boundary points from boundary-face values
//   k_bnd_pre         phiwStar on boundary faces, qgdFlux re-evaluation of p_b     (QGDFoam/updateFluxes.H:54-65)
//   k_face_flux       fused: 11 interpolations + 4 face gradients + QGD flux algebra (updateFields.H:45-80,
//                     updateFluxes.H:41-139, QGDCourantNo.H:38-50) -> 5 flux scalars per face
//   k_bnd_flux        same on boundary faces (ghost values, GaussVolPointBase3D.C:771-819)
//   k_dt              setDeltaT-QGDQHD.H:41-61 on the device (no host sync)
//   k_cell_update     atomic-free face->cell reduction over the signed cell-face CSR + Euler update
//                     (QGDRhoEqn/UEqn/EEqn.H) + hePsiQGDThermo::calculate + constScPrModel1::correct
//   k_bnd_post        boundary state after the update (correctBoundaryConditions sequence of QGDFoam.C:133-156)
#include <cfloat>
#include <cstdio>

#include "qgd_kernels.cuh"
#include "qgd_wedge.h"

namespace qgd {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ RecA loadA(const SolverView& sv, int i)
{
    const double* p = sv.S + i;
    const size_t n = sv.nCells;
    return RecA{__ldg(p), __ldg(p + n), __ldg(p + 2 * n), __ldg(p + 3 * n), __ldg(p + 4 * n), __ldg(p + 5 * n),
                __ldg(p + 6 * n), __ldg(p + 7 * n)};
}
__device__ __forceinline__ RecB loadB(const SolverView& sv, int i)
{
    const double* p = sv.S + 8 * (size_t)sv.nCells + i;
    const size_t n = sv.nCells;
    return RecB{__ldg(p), __ldg(p + n), __ldg(p + 2 * n), __ldg(p + 3 * n), __ldg(p + 4 * n), __ldg(p + 5 * n),
                __ldg(p + 6 * n), __ldg(p + 7 * n)};
}
__device__ __forceinline__ RecP loadP(const SolverView& sv, int i)
{
    const double* p = sv.P + i;
    const size_t n = sv.nPoints;
    return RecP{__ldg(p), __ldg(p + n), __ldg(p + 2 * n), __ldg(p + 3 * n), __ldg(p + 4 * n), __ldg(p + 5 * n)};
}
// L2 cache-policy variants of the gathers (k_face_flux_tma<.., HINT & 2>): cell / point state is re-read by later tiles
// (the y- and z-neighbours of a cell come back as owners 256 / 65536 cells later on a hex box) while the face constants
// and the fluxes stream through once, so the state is tagged evict_last and the streams evict_first.
__device__ __forceinline__ unsigned long long l2PolicyEvictLast()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2PolicyEvictFirst()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ldgKeep(const double* p, unsigned long long pol)
{
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ RecA loadAKeep(const SolverView& sv, int i, unsigned long long pol)
{
    const double* p = sv.S + i;
    const size_t n = sv.nCells;
    return RecA{ldgKeep(p, pol), ldgKeep(p + n, pol), ldgKeep(p + 2 * n, pol), ldgKeep(p + 3 * n, pol), ldgKeep(p + 4 * n, pol),
                ldgKeep(p + 5 * n, pol), ldgKeep(p + 6 * n, pol), ldgKeep(p + 7 * n, pol)};
}
__device__ __forceinline__ RecB loadBKeep(const SolverView& sv, int i, unsigned long long pol)
{
    const double* p = sv.S + 8 * (size_t)sv.nCells + i;
    const size_t n = sv.nCells;
    return RecB{ldgKeep(p, pol), ldgKeep(p + n, pol), ldgKeep(p + 2 * n, pol), ldgKeep(p + 3 * n, pol), ldgKeep(p + 4 * n, pol),
                ldgKeep(p + 5 * n, pol), ldgKeep(p + 6 * n, pol), ldgKeep(p + 7 * n, pol)};
}
__device__ __forceinline__ RecP loadPKeep(const SolverView& sv, int i, unsigned long long pol)
{
    const double* p = sv.P + i;
    const size_t n = sv.nPoints;
    return RecP{ldgKeep(p, pol), ldgKeep(p + n, pol), ldgKeep(p + 2 * n, pol), ldgKeep(p + 3 * n, pol), ldgKeep(p + 4 * n, pol),
                ldgKeep(p + 5 * n, pol)};
}
// 8 fields of cell i starting at field k0
__device__ __forceinline__ void storeRec(const SolverView& sv, int k0, int i, const double (&v)[8])
{
    double* p = sv.S + (size_t)k0 * sv.nCells + i;
    const size_t n = sv.nCells;
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k * n] = v[k];
}

// ---- thermo: perfectGas + hConst + sensibleInternalEnergy  [OF-v2312; SURVEY 8c item 9]
__device__ __forceinline__ double thermoEs(const Consts& k, double T)
{
    return k.eConst ? k.Cv * (T - k.Tref) + k.Esref             // eConstThermo::Es [OF-v2312]
                    : k.Cp * (T - k.Tref) + k.Hsref - k.R * T;  // hConstThermo::Hs - p/rho
}
// molecular viscosity and alphah of the transport model at temperature T (hePsiQGDThermo.C:59-62)
__device__ __forceinline__ double muMol(const Consts& k, double T)
{
    if (k.transport == 1) return k.mu0 * pow(T / k.T0, k.kExp);          // powerLawTransportI.H:121-128
    if (k.transport == 2) return k.As * sqrt(T) / (1.0 + k.Ts / T);      // sutherlandTransport::mu [OF-v2312]
    return k.mu;
}
__device__ __forceinline__ double alphahMol(const Consts& k, double muT)
{
    if (k.transport == 1) return muT * k.rPr;                            // powerLawTransportI.H:143-150
    if (k.transport == 2) return muT * k.Cv * (1.32 + 1.77 * k.R / k.Cv) / k.Cp;   // kappa/Cp, modified Eucken [OF-v2312]
    return k.mu / k.Pr;                                                  // constTransport
}
__device__ __forceinline__ double thermoTHE(const Consts& k, double e, double T0)
{
    double Test = T0, Tnew = T0;
    const double Ttol = T0 * 1e-4;
    int iter = 0;
    do {
        Test = Tnew;
        Tnew = Test - (thermoEs(k, Test) - e) / k.Cv;
        if (iter++ > 100) break;
    } while (fabs(Tnew - Test) > Ttol);
    return Tnew;
}

// ============================================================================ generic fvsc operator kernels
template <int K>
__global__ void k_point_gather_generic(int nPoints, int W, const int* __restrict__ ell, const double* __restrict__ ellW,
                                       const int* __restrict__ cnt, const int* __restrict__ tailOff,
                                       const int* __restrict__ tailCell, const double* __restrict__ tailW,
                                       const double* __restrict__ cell, double* __restrict__ pts)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPoints) return;
    const int n = cnt[p];
    if (n == 0) return;                       // patch point: written by k_patch_point_gather_generic
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.0;
    for (int q = 0; q < W && q < n; ++q) {
        const int c = ell[(size_t)q * nPoints + p];
        const double wq = ellW[(size_t)q * nPoints + p];
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] += wq * __ldg(&cell[(size_t)c * K + j]);
    }
    if (n > W)
        for (int q = tailOff[p]; q < tailOff[p + 1]; ++q) {
            const int c = tailCell[q];
            const double wq = tailW[q];
#pragma unroll
            for (int j = 0; j < K; ++j) acc[j] += wq * __ldg(&cell[(size_t)c * K + j]);
        }
#pragma unroll
    for (int j = 0; j < K; ++j) pts[(size_t)p * K + j] = acc[j];
}

template <int K>
__global__ void k_patch_point_gather_generic(int nPP, const int* __restrict__ patchPoints, const int* __restrict__ ppOff,
                                             const int* __restrict__ ppFace, const double* __restrict__ ppW,
                                             const double* __restrict__ bnd, double* __restrict__ pts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPP) return;
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.0;
    for (int q = ppOff[i]; q < ppOff[i + 1]; ++q) {
        const int b = ppFace[q];
        const double wq = ppW[q];
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] += wq * bnd[(size_t)b * K + j];
    }
    const int p = patchPoints[i];
#pragma unroll
    for (int j = 0; j < K; ++j) pts[(size_t)p * K + j] = acc[j];
}

// differences (phi[v1]-phi[v3], phi[v2]-phi[v4], phiP-phiN) of one face for K components
template <int K>
__device__ __forceinline__ void faceDiffs(const FaceView& fv, int f, const double* __restrict__ cell, const double* __restrict__ pts,
                                          const double* __restrict__ bnd, const double* __restrict__ bsg,
                                          const double* __restrict__ nbr, int flags, double (&d1)[K], double (&d2)[K], double (&dP)[K])
{
    const int P = fv.own[f];
    if (flags & FF_POINTS) {
        const int4 v = fv.vtx[f];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            d1[j] = pts[(size_t)v.x * K + j] - pts[(size_t)v.z * K + j];
            d2[j] = pts[(size_t)v.y * K + j] - pts[(size_t)v.w * K + j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
    }
    if (f < fv.nI) {
        const int N = fv.nei[f];
#pragma unroll
        for (int j = 0; j < K; ++j) dP[j] = cell[(size_t)P * K + j] - cell[(size_t)N * K + j];
    } else {
        const int b = f - fv.nI;
        const bool proc = fv.bKind[b] == QGD_PATCH_PROCESSOR;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double psiN = (proc && nbr) ? nbr[(size_t)b * K + j] : bnd[(size_t)b * K + j] + bsg[(size_t)b * K + j] * fv.halfDist[b];
            dP[j] = cell[(size_t)P * K + j] - psiN;
        }
    }
}

// leastSquares gradient of K components on internal face f: g[K*i + j] = sum_s c_s[i] (phi_s[j] - phi_f[j]),
// phi_f = linearInterpolate(phi)
template <int K>
__device__ __forceinline__ void lsqGradGeneric(const FaceView& fv, int f, const double* __restrict__ cell, double (&g)[3 * K])
{
    const int P = fv.own[f], N = fv.nei[f];
    const double w = fv.w[f];
    const size_t nI = fv.nI;
    double sF[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { const double a = cell[(size_t)P * K + j], b = cell[(size_t)N * K + j]; sF[j] = w * (a - b) + b; }
#pragma unroll
    for (int q = 0; q < 3 * K; ++q) g[q] = 0.0;
    for (int s = 0; s < fv.lsqW; ++s) {
        const int c = fv.lsqCells[(size_t)s * nI + f];
        const double cx = fv.lsqCoef[((size_t)s * 3 + 0) * nI + f], cy = fv.lsqCoef[((size_t)s * 3 + 1) * nI + f],
                     cz = fv.lsqCoef[((size_t)s * 3 + 2) * nI + f];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double d = cell[(size_t)c * K + j] - sF[j];
            g[j] += cx * d; g[K + j] += cy * d; g[2 * K + j] += cz * d;
        }
    }
}

// grad: out[f][3*i... ] tensor index K*i + j = d_i phi_j
template <int K>
__global__ void k_fvsc_grad(FaceView fv, const double* __restrict__ cell, const double* __restrict__ pts,
                            const double* __restrict__ bnd, const double* __restrict__ bsg, const double* __restrict__ nbr,
                            double* __restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fv.nF) return;
    double* o = out + (size_t)fv.perm[f] * 3 * K;
    const int flags = fv.flags[f];
    double g1[3], g2[3], gp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g1[i] = fv.G[(size_t)(0 + i) * fv.fs + f];
        g2[i] = fv.G[(size_t)(3 + i) * fv.fs + f];
        gp[i] = fv.G[6 * (size_t)fv.fs + f] * fv.Sf[(size_t)i * fv.fs + f];      // GP = gpS * Sf
    }
    if (f >= fv.nI) {
        const int b = f - fv.nI;
        if (fv.bKind[b] == QGD_PATCH_EMPTY) {
#pragma unroll
            for (int q = 0; q < 3 * K; ++q) o[q] = 0.0;
            return;
        }
        if (flags & FF_NORMAL_ONLY) {          // nf * snGrad  (GaussVolPointBase1D.C:49-63, GaussVolPointBase3D.C:808-817)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) o[K * i + j] = gp[i] * bsg[(size_t)b * K + j];
            return;
        }
    }
    if (flags & FF_LSQ) {                       // leastSquares::Grad  extendedFaceStencilScalarGrad.C:52-72
        double gl[3 * K];
        lsqGradGeneric<K>(fv, f, cell, gl);
#pragma unroll
        for (int q = 0; q < 3 * K; ++q) o[q] = gl[q];
        return;
    }
    double d1[K], d2[K], dP[K];
    faceDiffs<K>(fv, f, cell, pts, bnd, bsg, nbr, flags, d1, d2, dP);
    if (K == 3 && (flags & FF_TRI_QUIRK)) {     // GaussVolPointBase3D.C:844-854
#pragma unroll
        for (int row = 0; row < 3; ++row)
#pragma unroll
            for (int d = 0; d < 3; ++d) o[3 * row + d] = g1[d] * d1[d] + g2[d] * d2[d] + gp[d] * dP[d];
        return;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) o[K * i + j] = g1[i] * d1[j] + g2[i] * d2[j] + gp[i] * dP[j];
}

// div: K=3 -> scalar (sum_i d_i phi_i), K=9 -> vector_j = sum_i d_i T_ij
template <int K>
__global__ void k_fvsc_div(FaceView fv, const double* __restrict__ cell, const double* __restrict__ pts,
                           const double* __restrict__ bnd, const double* __restrict__ bsg, const double* __restrict__ nbr,
                           double* __restrict__ out)
{
    constexpr int OK = K / 3;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fv.nF) return;
    double* o = out + (size_t)fv.perm[f] * OK;
    const int flags = fv.flags[f];
    double g1[3], g2[3], gp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g1[i] = fv.G[(size_t)(0 + i) * fv.fs + f];
        g2[i] = fv.G[(size_t)(3 + i) * fv.fs + f];
        gp[i] = fv.G[6 * (size_t)fv.fs + f] * fv.Sf[(size_t)i * fv.fs + f];      // GP = gpS * Sf
    }
    if (f >= fv.nI) {
        const int b = f - fv.nI;
        if (fv.bKind[b] == QGD_PATCH_EMPTY) {
#pragma unroll
            for (int q = 0; q < OK; ++q) o[q] = 0.0;
            return;
        }
        if (flags & FF_NORMAL_ONLY) {          // nf & snGrad
#pragma unroll
            for (int j = 0; j < OK; ++j) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) s += gp[i] * bsg[(size_t)b * K + OK * i + j];
                o[j] = s;
            }
            return;
        }
    }
    if (flags & FF_LSQ) {                       // leastSquares::Div  leastSquaresStencil.C:204-275
        double gl[3 * K];
        lsqGradGeneric<K>(fv, f, cell, gl);
#pragma unroll
        for (int j = 0; j < OK; ++j) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) s += gl[K * i + OK * i + j];
            o[j] = s;
        }
        return;
    }
    double d1[K], d2[K], dP[K];
    faceDiffs<K>(fv, f, cell, pts, bnd, bsg, nbr, flags, d1, d2, dP);
#pragma unroll
    for (int j = 0; j < OK; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) s += g1[i] * d1[OK * i + j] + g2[i] * d2[OK * i + j] + gp[i] * dP[OK * i + j];
        o[j] = (K == 9 && j == fv.zeroDivCmpt) ? 0.0 : s;
    }
}

// ============================================================================ QGDFoam step kernels
struct FaceState {            // interpolated face quantities (QGDFoam/updateFields.H:45-80)
    double rho, U[3], rhoU[3], UrhoU[9], p, c, H, alpha, mu, tau;
};
struct FaceGrads { double U[9], e[3], rho[3], p[3]; };

// QGDFoam/updateFluxes.H:54-63 : rhoW* and phiwStar
__device__ __forceinline__ void qgdRhoWStar(const FaceState& s, const FaceGrads& g, double divU, double (&rhoW)[3])
{
    const double gRU = g.rho[0] * s.U[0] + g.rho[1] * s.U[1] + g.rho[2] * s.U[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double rUG = s.rhoU[0] * g.U[j] + s.rhoU[1] * g.U[3 + j] + s.rhoU[2] * g.U[6 + j];
        rhoW[j] = s.tau * (s.U[j] * gRU + s.rhoU[j] * divU + rUG);
    }
}

// QGDFoam/updateFluxes.H:67-139 (explicit branch): mass, momentum, energy face fluxes
__device__ __forceinline__ void qgdFluxes(const Consts& k, const FaceState& s, const FaceGrads& g, const double (&Sf)[3],
                                          double& Fm, double (&FU)[3], double& FE, double& phiw)
{
    const double divU = g.U[0] + g.U[4] + g.U[8];
    double rhoW[3];
    qgdRhoWStar(s, g, divU, rhoW);
    phiw = Sf[0] * rhoW[0] + Sf[1] * rhoW[1] + Sf[2] * rhoW[2];
    double jm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { rhoW[j] += s.tau * g.p[j]; jm[j] = s.rhoU[j] - rhoW[j]; }
    const double phiJm = Sf[0] * jm[0] + Sf[1] * jm[1] + Sf[2] * jm[2];
    Fm = phiJm;
    const double UgP = s.U[0] * g.p[0] + s.U[1] * g.p[1] + s.U[2] * g.p[2];
    const double iso = UgP + (k.gamma * s.p * divU);
    double Pi[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double UrUG = s.UrhoU[3 * i] * g.U[j] + s.UrhoU[3 * i + 1] * g.U[3 + j] + s.UrhoU[3 * i + 2] * g.U[6 + j];
            double v = s.tau * (UrUG + s.U[i] * g.p[j]) + s.tau * ((i == j) ? iso : 0.0);
            v += s.mu * (g.U[3 * i + j] + g.U[3 * j + i] - ((i == j) ? (2.0 / 3.0) * divU : 0.0));
            Pi[3 * i + j] = v;
        }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double phiPi = Sf[0] * Pi[j] + Sf[1] * Pi[3 + j] + Sf[2] * Pi[6 + j];
        FU[j] = phiJm * s.U[j] + Sf[j] * s.p - phiPi;
    }
    const double pr2 = s.p / s.rho / s.rho;
    double v[3], q[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = g.e[j] - pr2 * g.rho[j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        q[i] = -s.tau * (s.UrhoU[3 * i] * v[0] + s.UrhoU[3 * i + 1] * v[1] + s.UrhoU[3 * i + 2] * v[2]);
        q[i] -= s.alpha * g.e[i];
    }
    const double phiQ = Sf[0] * q[0] + Sf[1] * q[1] + Sf[2] * q[2];
    double PiU[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) PiU[i] = Pi[3 * i] * s.U[0] + Pi[3 * i + 1] * s.U[1] + Pi[3 * i + 2] * s.U[2];
    const double phiPiU = Sf[0] * PiU[0] + Sf[1] * PiU[1] + Sf[2] * PiU[2];
    FE = phiJm * s.H + phiQ - phiPiU;
}

__device__ __forceinline__ void gradsFromDiffs(const double (&g1)[3], const double (&g2)[3], const double (&gp)[3], int flags,
                                               const RecP& d1, const RecP& d2, const RecP& dP, FaceGrads& g)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g.rho[i] = g1[i] * d1.rho + g2[i] * d2.rho + gp[i] * dP.rho;
        g.e[i] = g1[i] * d1.e + g2[i] * d2.e + gp[i] * dP.e;
        g.p[i] = g1[i] * d1.p + g2[i] * d2.p + gp[i] * dP.p;
        g.U[3 * i + 0] = g1[i] * d1.Ux + g2[i] * d2.Ux + gp[i] * dP.Ux;
        g.U[3 * i + 1] = g1[i] * d1.Uy + g2[i] * d2.Uy + gp[i] * dP.Uy;
        g.U[3 * i + 2] = g1[i] * d1.Uz + g2[i] * d2.Uz + gp[i] * dP.Uz;
    }
    if (flags & FF_TRI_QUIRK) {                 // GaussVolPointBase3D.C:844-854
        const double dxx = g.U[0], dyy = g.U[4], dzz = g.U[8];
#pragma unroll
        for (int row = 0; row < 3; ++row) { g.U[3 * row] = dxx; g.U[3 * row + 1] = dyy; g.U[3 * row + 2] = dzz; }
    }
}

__device__ __forceinline__ RecP recDiff(const RecP& a, const RecP& b)
{
    return RecP{a.rho - b.rho, a.Ux - b.Ux, a.Uy - b.Uy, a.Uz - b.Uz, a.e - b.e, a.p - b.p};
}

__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

// block-wide max / min of non-negative doubles, one atomic per block
template <int BLOCK>
__device__ __forceinline__ void blockReduceCo(double coMax, double tauMin, StepScalars* sc)
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
        if (lane == 0) {
            atomicMax(&sc->coMaxBits, dbits(coMax));
            atomicMin(&sc->tauMinBits, dbits(tauMin));
        }
    }
}

// ---- cell -> point gather of (rho,U,e,p); ELL rows: every index/weight load of a warp is one coalesced line
template <int W>
__global__ void __launch_bounds__(kBlock) k_points(SolverView sv, const int* __restrict__ list, int nList)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nList) return;
    const int p = list ? __ldg(&list[i]) : i;
    const int cnt = __ldg(&sv.pcCount[p]);
    if (cnt == 0) return;                                   // patch point
    const size_t nP = sv.nPoints, n = sv.nCells;
    int ids[W];
    double wq[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        ids[j] = __ldg(&sv.pcEll[j * nP + p]);
        wq[j] = __ldg(&sv.pcEllWt[j * nP + p]);            // padded entries: weight 0, id = first cell of the row
    }
    double a[6] = {0, 0, 0, 0, 0, 0};
    double v[W][6];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const double* r = sv.S + ids[j];
#pragma unroll
        for (int k = 0; k < 6; ++k) v[j][k] = __ldg(r + k * n);
    }
#pragma unroll
    for (int j = 0; j < W; ++j)
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] += wq[j] * v[j][k];
    if (cnt > W)
        for (int q = __ldg(&sv.pcTailOff[p]); q < __ldg(&sv.pcTailOff[p + 1]); ++q) {
            const double* r = sv.S + __ldg(&sv.pcTailCell[q]);
            const double w1 = __ldg(&sv.pcTailW[q]);
#pragma unroll
            for (int k = 0; k < 6; ++k) a[k] += w1 * __ldg(r + k * n);
        }
    double* o = sv.P + p;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k * nP] = a[k];
}

// boundary points from boundary-face values; onlyP: refresh p only (after the qgdFlux re-evaluation)
__global__ void k_patch_points(SolverView sv, BndState bs, int onlyP)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.nPatchPoints) return;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
    for (int q = sv.ppOff[i]; q < sv.ppOff[i + 1]; ++q) {
        const int b = sv.ppFace[q];
        const double wq = sv.ppW[q];
        const RecA r = bs.A[b];
        a0 += wq * r.rho; a1 += wq * r.Ux; a2 += wq * r.Uy; a3 += wq * r.Uz; a4 += wq * r.e; a5 += wq * bs.pNew[b];
    }
    const size_t nP = sv.nPoints;
    double* o = sv.P + sv.patchPoints[i];
    o[5 * nP] = a5;
    if (onlyP) return;
    o[0] = a0; o[nP] = a1; o[2 * nP] = a2; o[3 * nP] = a3; o[4 * nP] = a4;
}

// [OF-v2312] pointConstraints::constrain on the vertices of the constraint patches (wedge, symmetryPlane): with the vertex's
// constraint tensor R (HostMesh::wedgeR: I - n n on one plane, d d where two planes meet) a vector becomes R.v, a tensor R.T.R^T.
// step form: the velocity rows (fields 1..3) of the SoA point array P[6][nPoints]
__global__ void k_wedge_points(int n, const int* __restrict__ pts, const double* __restrict__ R9, double* __restrict__ P, size_t nPoints)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pts[i];
    const double* R = R9 + 9 * (size_t)i;
    double* u = P + nPoints + p;
    const double u0 = u[0], u1 = u[nPoints], u2 = u[2 * nPoints];
    u[0] = R[0] * u0 + R[1] * u1 + R[2] * u2;
    u[nPoints] = R[3] * u0 + R[4] * u1 + R[5] * u2;
    u[2 * nPoints] = R[6] * u0 + R[7] * u1 + R[8] * u2;
}
// operator form: AoS point values pts[p*K + j], K = 3 (vector) or 9 (tensor)
template <int K>
__global__ void k_wedge_points_generic(int n, const int* __restrict__ list, const double* __restrict__ nrm, double* __restrict__ pv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = nrm[9 * (size_t)i + q];
    double* v = pv + (size_t)list[i] * K;
    if (K == 3) {
        const double u0 = v[0], u1 = v[1], u2 = v[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) v[j] = R[3 * j] * u0 + R[3 * j + 1] * u1 + R[3 * j + 2] * u2;
    } else {
        double t[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) { t[3 * a + b] = 0.0; for (int c = 0; c < 3; ++c) t[3 * a + b] += R[3 * a + c] * v[3 * c + b]; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) { double r = 0.0; for (int c = 0; c < 3; ++c) r += t[3 * a + c] * R[3 * b + c]; v[3 * a + b] = r; }
    }
}

// boundary-face inputs shared by k_bnd_pre and k_bnd_flux
struct BndFace { FaceState s; FaceGrads g; double Sf[3]; };

__device__ __forceinline__ void bndFaceSetup(const Consts& k, const FaceView& fv, const SolverView& sv, const BndState& bs,
                                             int b, bool usePNew, BndFace& o)
{
    const int f = fv.nI + b;
    const int P = fv.own[f];
    const RecA a = bs.A[b];
    const RecB bb = bs.B[b];
    const RecA cA = loadA(sv, P);
    const int flags = fv.flags[f];
    const double delta = fv.dC[f];
#pragma unroll
    for (int i = 0; i < 3; ++i) o.Sf[i] = fv.Sf[(size_t)i * fv.fs + f];
    // face values = boundary values
    o.s.rho = a.rho; o.s.U[0] = a.Ux; o.s.U[1] = a.Uy; o.s.U[2] = a.Uz;
    o.s.rhoU[0] = bb.rhoUx; o.s.rhoU[1] = bb.rhoUy; o.s.rhoU[2] = bb.rhoUz;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.s.UrhoU[3 * i + j] = o.s.U[i] * o.s.rhoU[j];
    o.s.p = a.p; o.s.c = bb.c; o.s.H = a.H; o.s.alpha = k.implicit ? 0.0 : bb.alphaEff; o.s.mu = k.implicit ? 0.0 : bb.mu;
    o.s.tau = k.tauMode == 1 ? bb.aByC : (k.tauMode == 2 ? bb.aByC * fv.hf[f] / bb.c : bb.aByC * fv.hf[f]);
    // patch snGrad per field [OF fvPatchField::snGrad / zeroGradient / fixedGradient]
    // slip [OF-v2312 basicSymmetryFvPatchField::snGrad]: (transform(I - 2nn, U_P) - U_P) deltaCoeffs/2 = deltaCoeffs (U_b - U_P)
    const bool fixU = bs.bcU[b] == QGD_BC_FIXED_VALUE || bs.bcU[b] == QGD_BC_SLIP, fixT = bs.bcT[b] == QGD_BC_FIXED_VALUE;
    const double pB = usePNew ? bs.pNew[b] : a.p;
    RecP sn;
    sn.rho = delta * (a.rho - cA.rho);
    sn.Ux = fixU ? delta * (a.Ux - cA.Ux) : 0.0;
    sn.Uy = fixU ? delta * (a.Uy - cA.Uy) : 0.0;
    sn.Uz = fixU ? delta * (a.Uz - cA.Uz) : 0.0;
    sn.e = fixT ? delta * (a.e - cA.e) : 0.0;
    const int bcP = bs.bcP[b];
    sn.p = (bcP == QGD_BC_FIXED_VALUE) ? delta * (pB - cA.p) : ((bcP == QGD_BC_ZERO_GRADIENT) ? 0.0 : bs.pGrad[b]);
    double g1[3], g2[3], gp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g1[i] = fv.G[(size_t)(0 + i) * fv.fs + f];
        g2[i] = fv.G[(size_t)(3 + i) * fv.fs + f];
        gp[i] = fv.G[6 * (size_t)fv.fs + f] * fv.Sf[(size_t)i * fv.fs + f];      // GP = gpS * Sf
    }
    if (flags & FF_NORMAL_ONLY) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o.g.rho[i] = gp[i] * sn.rho; o.g.e[i] = gp[i] * sn.e; o.g.p[i] = gp[i] * sn.p;
            o.g.U[3 * i] = gp[i] * sn.Ux; o.g.U[3 * i + 1] = gp[i] * sn.Uy; o.g.U[3 * i + 2] = gp[i] * sn.Uz;
        }
        return;
    }
    const double hd = fv.halfDist[b];
    RecP d1{0, 0, 0, 0, 0, 0}, d2{0, 0, 0, 0, 0, 0};
    if (flags & FF_POINTS) {
        const int4 v = fv.vtx[f];
        d1 = recDiff(loadP(sv, v.x), loadP(sv, v.z));
        d2 = recDiff(loadP(sv, v.y), loadP(sv, v.w));
    }
    // phiP - psiN, psiN = phi_b + snGrad_b*|d|/2   (GaussVolPointBase3D.C:790-793)
    RecP dP;
    dP.rho = cA.rho - (a.rho + sn.rho * hd);
    dP.Ux = cA.Ux - (a.Ux + sn.Ux * hd);
    dP.Uy = cA.Uy - (a.Uy + sn.Uy * hd);
    dP.Uz = cA.Uz - (a.Uz + sn.Uz * hd);
    dP.e = cA.e - (a.e + sn.e * hd);
    dP.p = cA.p - (pB + sn.p * hd);
    gradsFromDiffs(g1, g2, gp, flags, d1, d2, dP, o.g);
}

// phiwStar on boundary faces, then qgdFlux::updateCoeffs + fixedGradient::evaluate for p
__global__ void k_bnd_pre(Consts k, FaceView fv, SolverView sv, BndState bs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    const int f = fv.nI + b;
    if (fv.bKind[b] == QGD_PATCH_EMPTY || fv.own[f] >= sv.nOwned) return;      // halo faces: state comes from the owner rank
    BndFace bf;
    bndFaceSetup(k, fv, sv, bs, b, false, bf);
    const double divU = bf.g.U[0] + bf.g.U[4] + bf.g.U[8];
    double rhoW[3];
    qgdRhoWStar(bf.s, bf.g, divU, rhoW);
    const double phiw = bf.Sf[0] * rhoW[0] + bf.Sf[1] * rhoW[1] + bf.Sf[2] * rhoW[2];
    bs.phiw[b] = phiw;
    double pNew = bs.A[b].p;
    if (bs.bcP[b] == QGD_BC_QGD_FLUX && !k.reducedScheme) {
        const double grad = -(phiw / bf.s.tau / fv.magSf[f]);     // qgdFluxFvPatchScalarField.C:184-192
        bs.pGrad[b] = grad;
        pNew = loadA(sv, fv.own[f]).p + grad / fv.dC[f];
    }
    bs.pNew[b] = pNew;
}

__global__ void k_bnd_flux(Consts k, FaceView fv, SolverView sv, BndState bs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double coMax = 0.0, tauMin = DBL_MAX;
    if (b < fv.nB) {
        const int f = fv.nI + b;
        if (fv.bKind[b] == QGD_PATCH_EMPTY || fv.own[f] >= sv.nOwned) {
#pragma unroll
            for (int q = 0; q < 5; ++q) sv.FB[q][b] = 0.0;
        } else {
            BndFace bf;
            bndFaceSetup(k, fv, sv, bs, b, true, bf);
            double Fm, FU[3], FE, phiw;
            qgdFluxes(k, bf.s, bf.g, bf.Sf, Fm, FU, FE, phiw);
            sv.FB[0][b] = Fm; sv.FB[1][b] = FU[0]; sv.FB[2][b] = FU[1]; sv.FB[3][b] = FU[2]; sv.FB[4][b] = FE;
            const double ms = fv.magSf[f];
            const double Unf = bf.s.U[0] * (bf.Sf[0] / ms) + bf.s.U[1] * (bf.Sf[1] / ms) + bf.s.U[2] * (bf.Sf[2] / ms);
            coMax = fmax(fabs(Unf + bf.s.c), fabs(Unf - bf.s.c)) / fv.hf[f];
            tauMin = bf.s.tau;
        }
    }
    blockReduceCo<kBlock>(coMax, tauMin, sv.sc);
}

// ---- one internal face: 11 interpolations + 4 GaussVolPoint gradients + QGD flux algebra -> 5 flux doubles at `slot`
// SMEM: the streamed per-face constants (G1, G2, gpS | Sf | w, hQGDf, |Sf|) were staged by TMA into shared memory: sd[k*kTmaTile + li]
constexpr int kTmaTile = 256;
// leastSquares gradients of (rho, U, e, p) on an internal face of the step kernel
__device__ __forceinline__ void lsqGrads(const FaceView& fv, const SolverView& sv, int f, const RecA& aP, const RecA& aN, double w, FaceGrads& g)
{
    const size_t nI = fv.nI, n = sv.nCells;
    const double sF[6] = {w * (aP.rho - aN.rho) + aN.rho, w * (aP.Ux - aN.Ux) + aN.Ux, w * (aP.Uy - aN.Uy) + aN.Uy,
                          w * (aP.Uz - aN.Uz) + aN.Uz, w * (aP.e - aN.e) + aN.e, w * (aP.p - aN.p) + aN.p};
    double acc[3][6];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[i][q] = 0.0;
    for (int s = 0; s < fv.lsqW; ++s) {
        const int c = __ldg(&fv.lsqCells[(size_t)s * nI + f]);
        const double cf[3] = {__ldg(&fv.lsqCoef[((size_t)s * 3 + 0) * nI + f]), __ldg(&fv.lsqCoef[((size_t)s * 3 + 1) * nI + f]),
                              __ldg(&fv.lsqCoef[((size_t)s * 3 + 2) * nI + f])};
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double d = __ldg(sv.S + q * n + c) - sF[q];
#pragma unroll
            for (int i = 0; i < 3; ++i) acc[i][q] += cf[i] * d;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g.rho[i] = acc[i][0]; g.U[3 * i] = acc[i][1]; g.U[3 * i + 1] = acc[i][2]; g.U[3 * i + 2] = acc[i][3];
        g.e[i] = acc[i][4]; g.p[i] = acc[i][5];
    }
}

template <bool ADJUST, bool SMEM = false, bool LSQ = false, int HINT = 0>
__device__ __forceinline__ void faceFluxOne(const Consts& k, const FaceView& fv, const SolverView& sv, int f, size_t slot, int P, int N,
                                            int flagsCur, const int4& v, double& coMax, double& tauMin, const double* sd = nullptr, int li = 0,
                                            unsigned long long polKeep = 0ull)
{
    (void)f;
    const RecA aP = (HINT & 2) ? loadAKeep(sv, P, polKeep) : loadA(sv, P), aN = (HINT & 2) ? loadAKeep(sv, N, polKeep) : loadA(sv, N);
    const RecB bP = (HINT & 2) ? loadBKeep(sv, P, polKeep) : loadB(sv, P), bN = (HINT & 2) ? loadBKeep(sv, N, polKeep) : loadB(sv, N);
    RecP d1{0, 0, 0, 0, 0, 0}, d2{0, 0, 0, 0, 0, 0};
    if (flagsCur & FF_POINTS) {
        if (HINT & 2) {
            d1 = recDiff(loadPKeep(sv, v.x, polKeep), loadPKeep(sv, v.z, polKeep));
            d2 = recDiff(loadPKeep(sv, v.y, polKeep), loadPKeep(sv, v.w, polKeep));
        } else {
            d1 = recDiff(loadP(sv, v.x), loadP(sv, v.z));
            d2 = recDiff(loadP(sv, v.y), loadP(sv, v.w));
        }
    }
    const RecP dP{aP.rho - aN.rho, aP.Ux - aN.Ux, aP.Uy - aN.Uy, aP.Uz - aN.Uz, aP.e - aN.e, aP.p - aN.p};
    double g1[3], g2[3], gp[3], Sf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Sf[i] = SMEM ? sd[(7 + i) * kTmaTile + li] : __ldg(&fv.Sf[(size_t)i * fv.fs + f]);
    {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g1[i] = SMEM ? sd[(0 + i) * kTmaTile + li] : __ldg(&fv.G[(size_t)(0 + i) * fv.fs + f]);
            g2[i] = SMEM ? sd[(3 + i) * kTmaTile + li] : __ldg(&fv.G[(size_t)(3 + i) * fv.fs + f]);
            gp[i] = (SMEM ? sd[6 * kTmaTile + li] : __ldg(&fv.G[6 * (size_t)fv.fs + f])) * Sf[i];    // GP = gpS * Sf
        }
    }
    FaceGrads g;
    // linearInterpolate: w*(phiP - phiN) + phiN   [OF surfaceInterpolationScheme::interpolate]
    const double w = SMEM ? sd[10 * kTmaTile + li] : __ldg(&fv.w[f]);
    if (LSQ && (flagsCur & FF_LSQ)) lsqGrads(fv, sv, f, aP, aN, w, g);
    else gradsFromDiffs(g1, g2, gp, flagsCur, d1, d2, dP, g);
    FaceState s;
    s.rho = w * (aP.rho - aN.rho) + aN.rho;
    s.U[0] = w * (aP.Ux - aN.Ux) + aN.Ux; s.U[1] = w * (aP.Uy - aN.Uy) + aN.Uy; s.U[2] = w * (aP.Uz - aN.Uz) + aN.Uz;
    s.rhoU[0] = w * (bP.rhoUx - bN.rhoUx) + bN.rhoUx; s.rhoU[1] = w * (bP.rhoUy - bN.rhoUy) + bN.rhoUy;
    s.rhoU[2] = w * (bP.rhoUz - bN.rhoUz) + bN.rhoUz;
    {
        const double uP[3] = {aP.Ux, aP.Uy, aP.Uz}, uN[3] = {aN.Ux, aN.Uy, aN.Uz};
        const double rP[3] = {bP.rhoUx, bP.rhoUy, bP.rhoUz}, rN[3] = {bN.rhoUx, bN.rhoUy, bN.rhoUz};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double tP = uP[i] * rP[j], tN = uN[i] * rN[j];
                s.UrhoU[3 * i + j] = w * (tP - tN) + tN;
            }
    }
    s.p = w * (aP.p - aN.p) + aN.p;
    s.c = w * (bP.c - bN.c) + bN.c;
    s.H = w * (aP.H - aN.H) + aN.H;
    // implicitDiffusion: the mu / alpha terms are solved implicitly, the explicit fluxes carry none (updateFluxes.H:95,131)
    s.alpha = k.implicit ? 0.0 : w * (bP.alphaEff - bN.alphaEff) + bN.alphaEff;
    s.mu = k.implicit ? 0.0 : w * (bP.mu - bN.mu) + bN.mu;
    const double hf = SMEM ? sd[11 * kTmaTile + li] : __ldg(&fv.hf[f]);
    {
        const double tI = w * (bP.aByC - bN.aByC) + bN.aByC;
        s.tau = k.tauMode == 0 ? tI * hf                     // constScPrModel1.C:103
                               : (k.tauMode == 1 ? tI        // constScPrModel1n.C:128
                                                 : tI * hf / s.c);   // constScPrModel1n.C:104-105: I(alphaQGD) hQGDf / I(c)
    }
    double Fm, FU[3], FE, phiw;
    qgdFluxes(k, s, g, Sf, Fm, FU, FE, phiw);
    if (HINT & 1) {                                         // streamed once: evict-first stores
        __stcs(&sv.FI[0][slot], Fm); __stcs(&sv.FI[1][slot], FU[0]); __stcs(&sv.FI[2][slot], FU[1]); __stcs(&sv.FI[3][slot], FU[2]);
        __stcs(&sv.FI[4][slot], FE);
    } else {
        sv.FI[0][slot] = Fm; sv.FI[1][slot] = FU[0]; sv.FI[2][slot] = FU[1]; sv.FI[3][slot] = FU[2]; sv.FI[4][slot] = FE;
    }
    if (ADJUST) {                                           // QGDCourantNo.H:38-50
        const double ms = SMEM ? sd[12 * kTmaTile + li] : __ldg(&fv.magSf[f]);
        const double Unf = s.U[0] * (Sf[0] / ms) + s.U[1] * (Sf[1] / ms) + s.U[2] * (Sf[2] / ms);
        coMax = fmax(coMax, fmax(fabs(Unf + s.c), fabs(Unf - s.c)) / hf);
        tauMin = fmin(tauMin, s.tau);
    }
}

// ---- the fused internal-face kernel (two-kernel form: fluxes go to a full-size array)
template <bool ADJUST>
__global__ void __launch_bounds__(kBlock, 2) k_face_flux(Consts k, FaceView fv, SolverView sv)
{
    double coMax = 0.0, tauMin = DBL_MAX;
    // software-pipelined indices: the addressing of face f+stride is fetched while face f is computed, so each
    // iteration exposes one memory latency (the gathers), not two (indices -> gathers)
    const int stride = gridDim.x * blockDim.x;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    int P = 0, N = 0, flags = 0;
    int4 v = make_int4(0, 0, 0, 0);
    const int nIA = fv.nIActive;
    if (f < nIA) { P = __ldg(&fv.own[f]); N = __ldg(&fv.nei[f]); flags = __ldg(&fv.flags[f]); v = __ldg(&fv.vtx[f]); }
    for (; f < nIA; f += stride) {
        const int Pc = P, Nc = N, flagsCur = flags;
        const int4 vc = v;
        {
            const int fn = f + stride;
            if (fn < nIA) { P = __ldg(&fv.own[fn]); N = __ldg(&fv.nei[fn]); flags = __ldg(&fv.flags[fn]); v = __ldg(&fv.vtx[fn]); }
        }
        faceFluxOne<ADJUST>(k, fv, sv, f, (size_t)f, Pc, Nc, flagsCur, vc, coMax, tauMin);
    }
    if (ADJUST) blockReduceCo<kBlock>(coMax, tauMin, sv.sc);
}

// setDeltaT-QGDQHD.H:41-61 ; QGDCourantNo.H:47-50
__global__ void k_dt(StepScalars* sc, int* pipeQueue)
{
    if (pipeQueue) *pipeQueue = 0;
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
    sc->stepIndex += 1;
    sc->coMaxBits = 0ull;
    sc->tauMinBits = dbits(DBL_MAX);
}

// per-cell closing of the step: new thermo state from (rho, U, rhoU, rhoE, e) and the OLD p, T
__device__ __forceinline__ void cellThermo(const Consts& k, double rho, const double (&U)[3], const double (&rhoU)[3], double rhoE,
                                           double e, double pOld, double TOld, double aQGD, double hQGD, const SolverView& sv, int cell,
                                           bool haveU = true)
{
    // hePsiQGDThermo::calculate  hePsiQGDThermo.C:48-64,123-124
    const double T = thermoTHE(k, e, TOld);
    const double psi = 1.0 / (k.R * T);
    const double c = sqrt(k.gamma / psi);
    // constScPrModel1::correct  constScPrModel1.C:104-114  (p is still the old pressure here: QGDFoam.C:149-154)
    const bool tauByU = (k.model == 1) && haveU;           // constScPrModel1n.C:119-126
    const double tau = tauByU ? aQGD * hQGD / (sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]) + c) : aQGD * hQGD / c;
    const double muQGD = pOld * (sv.scVar ? sv.scVar[cell] : k.ScQGD) * tau;     // varScModel6.C:313-322
    const double alphauQGD = muQGD / k.PrQGD;
    const double muT = muMol(k, T);
    const double mu = muT + muQGD;                       // QGDThermo.C:91-98
    const double alpha = alphahMol(k, muT) + alphauQGD;
    const double p = rho / psi;                          // QGDFoam.C:152-154
    const double H = (rhoE + p) / rho;                   // updateFields.H:71 (of the next step)
    const double a[8] = {rho, U[0], U[1], U[2], e, p, T, H};
    // slot `aByC`: alphaQGD/c (tauMode 0) | tauQGD (tauMode 1) | alphaQGD (tauMode 2: constScPrModel1n until "U" is registered)
    const double slot = tauByU ? tau : ((k.model == 1) ? aQGD : aQGD / c);
    const double b[8] = {rhoU[0], rhoU[1], rhoU[2], rhoE, c, mu, k.alphaEffGamma ? k.gamma * alpha : alpha, slot};
    storeRec(sv, 0, cell, a);
    storeRec(sv, 8, cell, b);
    if (sv.tauOut) sv.tauOut[cell] = (k.model == 2) ? tau + muT / (pOld * k.ScQGD) : tau;      // constScPrModel2.C:112
}

// ---- leastSquares variant (2D / 1D meshes only, fvsc.C:60-63): cell-stencil gradients instead of the G record
template <bool ADJUST>
__global__ void __launch_bounds__(256, 2) k_face_flux_lsq(Consts k, FaceView fv, SolverView sv)
{
    double coMax = 0.0, tauMin = DBL_MAX;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < fv.nIActive; f += gridDim.x * blockDim.x)
        faceFluxOne<ADJUST, false, true>(k, fv, sv, f, (size_t)f, __ldg(&fv.own[f]), __ldg(&fv.nei[f]), __ldg(&fv.flags[f]),
                                                __ldg(&fv.vtx[f]), coMax, tauMin);
    if (ADJUST) blockReduceCo<256>(coMax, tauMin, sv.sc);
}

// ---- TMA-staged variant of the face kernel.  The 140 B of streamed per-face constants (own, nei, flags, vtx, G[9],
// Sf[3], w, hQGDf [, |Sf|]) of a 256-face tile are fetched by 18 (19) bulk asynchronous copies (cp.async.bulk -> UBLKCP)
// into a 2-stage shared-memory ring, armed on an mbarrier with the expected byte count; while tile i is computed, tile
// i+1 is already in flight, so only the cell/point gathers remain on the register-latency path.
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

__device__ __forceinline__ void bulkLoadHint(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar, unsigned long long pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)), "l"(pol)
                 : "memory");
}

template <bool ADJUST> struct TmaStage {
    static constexpr int kDoubles = ADJUST ? 13 : 12;       // G1[3] G2[3] gpS | Sf[3] | w | hQGDf [| |Sf|]
    static constexpr int kBytes = kTmaTile * (3 * 4 + 16 + 8 * kDoubles);      // stage size (flags slot included)
    // layout inside a stage: vtx int4[T] | doubles [kDoubles][T] | own int[T] | nei int[T] | flags int[T]
    static constexpr int oVtx = 0, oD = kTmaTile * 16, oOwn = oD + kTmaTile * 8 * kDoubles, oNei = oOwn + kTmaTile * 4, oFlags = oNei + kTmaTile * 4;
};

template <bool ADJUST, bool STREAM = false>
__device__ __forceinline__ void tmaIssueTile(const FaceView& fv, unsigned char* stage, unsigned long long* bar, int tile, unsigned long long pol = 0ull)
{
    using L = TmaStage<ADJUST>;
    const size_t f0 = (size_t)tile * kTmaTile, fs = fv.fs;
    // uniform flags (hex / single-type meshes): the 4-byte column is not streamed at all
    mbarExpectTx(bar, (unsigned)(L::kBytes - (fv.flagsUniform >= 0 ? kTmaTile * 4 : 0)));
    auto bulkLoad = [pol](void* d, const void* s, unsigned b, unsigned long long* m) {
        if (STREAM) bulkLoadHint(d, s, b, m, pol);
        else qgd::bulkLoad(d, s, b, m);
    };
    bulkLoad(stage + L::oVtx, fv.vtx + f0, kTmaTile * 16, bar);
    double* d = reinterpret_cast<double*>(stage + L::oD);
#pragma unroll
    for (int q = 0; q < 7; ++q) bulkLoad(d + q * kTmaTile, fv.G + q * fs + f0, kTmaTile * 8, bar);
#pragma unroll
    for (int q = 0; q < 3; ++q) bulkLoad(d + (7 + q) * kTmaTile, fv.Sf + q * fs + f0, kTmaTile * 8, bar);
    bulkLoad(d + 10 * kTmaTile, fv.w + f0, kTmaTile * 8, bar);
    bulkLoad(d + 11 * kTmaTile, fv.hf + f0, kTmaTile * 8, bar);
    if (ADJUST) bulkLoad(d + 12 * kTmaTile, fv.magSf + f0, kTmaTile * 8, bar);
    bulkLoad(stage + L::oOwn, fv.own + f0, kTmaTile * 4, bar);
    bulkLoad(stage + L::oNei, fv.nei + f0, kTmaTile * 4, bar);
    if (fv.flagsUniform < 0) bulkLoad(stage + L::oFlags, fv.flags + f0, kTmaTile * 4, bar);
}

template <bool ADJUST, int HINT>
__global__ void __launch_bounds__(kTmaTile, 2) k_face_flux_tma(Consts k, FaceView fv, SolverView sv)
{
    const unsigned long long polKeep = (HINT & 2) ? l2PolicyEvictLast() : 0ull;
    const unsigned long long polStream = (HINT & 1) ? l2PolicyEvictFirst() : 0ull;
    using L = TmaStage<ADJUST>;
    constexpr int NST = 2;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[NST];
    double coMax = 0.0, tauMin = DBL_MAX;
    const int nIA = fv.nIActive;
    const int nTiles = nIA / kTmaTile;                       // full tiles go through TMA, the remainder through plain loads
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NST; ++q) mbarInit(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int tile = blockIdx.x;
    if (threadIdx.x == 0) {                                  // prologue: NST-1 tiles in flight
#pragma unroll
        for (int q = 0; q < NST - 1; ++q)
            if (tile + q * (int)gridDim.x < nTiles)
                tmaIssueTile<ADJUST, (HINT & 1) != 0>(fv, smem + q * L::kBytes, &bars[q], tile + q * gridDim.x, polStream);
    }
    for (int it = 0; tile < nTiles; ++it, tile += gridDim.x) {
        const int st = it % NST;
        const int next = tile + (NST - 1) * (int)gridDim.x;    // refills the stage consumed in the previous iteration
        const int sn = (it + NST - 1) % NST;
        if (threadIdx.x == 0 && next < nTiles) tmaIssueTile<ADJUST, (HINT & 1) != 0>(fv, smem + sn * L::kBytes, &bars[sn], next, polStream);
        mbarWait(&bars[st], (unsigned)((it / NST) & 1));
        const unsigned char* sg = smem + st * L::kBytes;
        const int li = threadIdx.x;
        const int f = tile * kTmaTile + li;
        const int P = reinterpret_cast<const int*>(sg + L::oOwn)[li], N = reinterpret_cast<const int*>(sg + L::oNei)[li];
        const int flags = fv.flagsUniform >= 0 ? fv.flagsUniform : reinterpret_cast<const int*>(sg + L::oFlags)[li];
        const int4 v = reinterpret_cast<const int4*>(sg + L::oVtx)[li];
        faceFluxOne<ADJUST, true, false, HINT>(k, fv, sv, f, (size_t)f, P, N, flags, v, coMax, tauMin,
                                                      reinterpret_cast<const double*>(sg + L::oD), li, polKeep);
        __syncthreads();                                     // every thread is done with this stage before it is refilled
    }
    // remainder (< one tile): plain loads, one CTA
    if (blockIdx.x == gridDim.x - 1) {
        const int f = nTiles * kTmaTile + (int)threadIdx.x;
        if (f < nIA)
            faceFluxOne<ADJUST>(k, fv, sv, f, (size_t)f, __ldg(&fv.own[f]), __ldg(&fv.nei[f]), __ldg(&fv.flags[f]), __ldg(&fv.vtx[f]),
                                              coMax, tauMin);
    }
    if (ADJUST) blockReduceCo<kTmaTile>(coMax, tauMin, sv.sc);
}

// flux of face f (device id) as seen by the cell update: internal faces live in the flux array / ring, boundary faces
// in their own arrays.  Ring reads bypass L1 (slots are rewritten within one launch of the pipelined kernel).
template <bool RING>
__device__ __forceinline__ void loadFlux(const SolverView& sv, int nI, int f, double& fm, double& f0, double& f1, double& f2, double& fe)
{
    if (RING) {      // ring slots are rewritten within one launch: read through L2 only
        const bool internal = f < nI;
        const int idx = internal ? (int)((unsigned)f % (unsigned)sv.ringSize) : f - nI;
        fm = __ldcg((internal ? sv.FI[0] : sv.FB[0]) + idx); f0 = __ldcg((internal ? sv.FI[1] : sv.FB[1]) + idx);
        f1 = __ldcg((internal ? sv.FI[2] : sv.FB[2]) + idx); f2 = __ldcg((internal ? sv.FI[3] : sv.FB[3]) + idx);
        fe = __ldcg((internal ? sv.FI[4] : sv.FB[4]) + idx);
    } else {         // one [5][nF] array: FB[k] = FI[k] + nI, so every face is FI[k][f]
        fm = __ldg(&sv.FI[0][f]); f0 = __ldg(&sv.FI[1][f]); f1 = __ldg(&sv.FI[2][f]); f2 = __ldg(&sv.FI[3][f]); fe = __ldg(&sv.FI[4][f]);
    }
}

// a, b: old state of the cell (only rho,U,e,p,T and rhoU,rhoE are used); enc: its ELL row; V, aQ, hQ: cell constants
template <int W, bool RING>
__device__ __forceinline__ void cellUpdateCore(const Consts& k, const SolverView& sv, int nI, int c, const RecA& a, const RecB& b,
                                               const int (&enc)[W], double V, double aQ, double hQ)
{
    double sm = 0.0, su0 = 0.0, su1 = 0.0, su2 = 0.0, se = 0.0;
    // fvc::surfaceIntegrate: the cell's faces in ascending polyMesh order, no atomics (ELL row, then CSR tail)
    {
        double fm[W], f0[W], f1[W], f2[W], fe[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (enc[j] >= 0) loadFlux<RING>(sv, nI, enc[j] >> 1, fm[j], f0[j], f1[j], f2[j], fe[j]);
            else { fm[j] = f0[j] = f1[j] = f2[j] = fe[j] = 0.0; }
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const double sgn = (enc[j] & 1) ? -1.0 : 1.0;
            sm += sgn * fm[j]; su0 += sgn * f0[j]; su1 += sgn * f1[j]; su2 += sgn * f2[j]; se += sgn * fe[j];
        }
    }
    if (enc[W - 1] >= 0)
        for (int q = __ldg(&sv.cfTailOff[c]); q < __ldg(&sv.cfTailOff[c + 1]); ++q) {
            const int e1 = __ldg(&sv.cfTailEnc[q]);
            const double sgn = (e1 & 1) ? -1.0 : 1.0;
            double fm, f0, f1, f2, fe;
            loadFlux<RING>(sv, nI, e1 >> 1, fm, f0, f1, f2, fe);
            sm += sgn * fm; su0 += sgn * f0; su1 += sgn * f1; su2 += sgn * f2; se += sgn * fe;
        }
    const double rDeltaT = 1.0 / sv.sc->dt;
    const double diag = rDeltaT * V;
    // explicit source matrices rhoSu / rhoUSu / rhoESu (QGDRhoEqn.H:46, QGDUEqn.H:85, QGDEEqn.H:71): zero in QGDFoam
    // (createZeroSources.H:28-44), the cloud's sources in particlesQGDFoam; volume-integrated, added to the matrix source
    const bool src = sv.su != nullptr;
    const size_t nS = sv.nCells;
    // QGDRhoEqn.H:40-47
    const double rho = src ? (rDeltaT * a.rho * V - V * (sm / V) + __ldg(sv.su + c)) / diag : (rDeltaT * a.rho * V - V * (sm / V)) / diag;
    // QGDUEqn.H:36-51, 79-86
    double rhoU[3], U[3];
    rhoU[0] = (rDeltaT * b.rhoUx * V - V * (su0 / V)) / diag;
    rhoU[1] = (rDeltaT * b.rhoUy * V - V * (su1 / V)) / diag;
    rhoU[2] = (rDeltaT * b.rhoUz * V - V * (su2 / V)) / diag;
    const double diagR = rDeltaT * rho * V;
    U[0] = rDeltaT * a.rho * a.Ux * V + V * (rDeltaT * (rhoU[0] - b.rhoUx));
    U[1] = rDeltaT * a.rho * a.Uy * V + V * (rDeltaT * (rhoU[1] - b.rhoUy));
    U[2] = rDeltaT * a.rho * a.Uz * V + V * (rDeltaT * (rhoU[2] - b.rhoUz));
    if (src) { U[0] += __ldg(sv.su + nS + c); U[1] += __ldg(sv.su + 2 * nS + c); U[2] += __ldg(sv.su + 3 * nS + c); }   // U only: rhoU keeps its value (QGDUEqn.H:79-89)
    U[0] /= diagR; U[1] /= diagR; U[2] /= diagR;
    // QGDEEqn.H:37-50, 65-73
    const double rhoE = (rDeltaT * b.rhoE * V - V * (se / V)) / diag;
    double e = rhoE / rho - 0.5 * (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
    const double ddt = k.energyQuirk ? (rDeltaT * (rhoE - b.rhoE)) : (rDeltaT * (rho * e - a.rho * a.e));
    e = rDeltaT * a.rho * a.e * V + V * ddt;
    if (src) e += __ldg(sv.su + 4 * nS + c);
    e /= diagR;
    if (e <= 0.0 || rho <= 0.0) atomicCAS(&sv.sc->guardStep, 0, sv.sc->stepIndex);      // QGDFoam.C:142-147 (rare: no cost otherwise)
    cellThermo(k, rho, U, rhoU, rhoE, e, a.p, a.T, aQ, hQ, sv, c);
}

template <int W, bool RING>
__device__ __forceinline__ void cellUpdateOne(const Consts& k, const SolverView& sv, int nI, int c)
{
    const RecA a = loadA(sv, c);
    const RecB b = loadB(sv, c);
    int enc[W];
#pragma unroll
    for (int j = 0; j < W; ++j) enc[j] = __ldg(&sv.cfEll[(size_t)j * sv.nCells + c]);
    cellUpdateCore<W, RING>(k, sv, nI, c, a, b, enc, __ldg(&sv.V[c]), __ldg(&sv.aQGD[c]), __ldg(&sv.hQGD[c]));
}

template <int W>
__global__ void __launch_bounds__(kBlock) k_cell_update(Consts k, SolverView sv, int nI)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nOwned) return;
    cellUpdateOne<W, false>(k, sv, nI, c);
}

// ---- pipelined face+cell kernel (fixed deltaT): see DESIGN.md "flux ring".
// Faces are owner-sorted and owner < neighbour, so once the faces owned by cells < X are done, cells < X are final and no
// later face reads them.  A persistent grid pulls work items from an in-order queue: F(k) = fluxes of the faces owned by
// cell chunk k, C(k) = update of cell chunk k, C(k) queued `lag` chunks behind F(k).  Fluxes live in an L2-resident ring
// (slot = face % ringSize) instead of a full HBM array; completion flags (epoch-valued) order producers and consumers:
//   C(k) waits for F(depLo[k] .. k)                 (all faces of its cells)
//   F(k) waits for C(ringLo[k] .. ringHi[k])        (consumers of the ring slots it overwrites)
// Every wait targets items that were dequeued earlier (checked on the host when the plan is built), so the persistent
// grid cannot deadlock.
__device__ __forceinline__ int ldAcquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// block-wide wait until every listed item carries this epoch: warp 0 polls up to 32 flags per round
__device__ __forceinline__ void waitList(const int* flags, const int* list, int lo, int hi, int epoch)
{
    if (threadIdx.x < 32) {
        for (int base = lo; base < hi; base += 32) {
            const int i = base + (int)threadIdx.x;
            const int* fp = flags + (i < hi ? __ldg(&list[i]) : 0);
            bool ok;
            do {
                ok = (i >= hi) || (ldAcquire(fp) == epoch);
            } while (!__all_sync(0xffffffffu, ok));
        }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void publish(int* flags, int item, int epoch)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(flags + item, epoch);
}

template <int W, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_face_cell_pipeline(Consts k, FaceView fv, SolverView sv, PipeView pv)
{
    __shared__ int sItem[2];
    const int nCh = pv.nChunks, lag = pv.lag;
    const int nItems = 2 * nCh;
    double coMax = 0.0, tauMin = DBL_MAX;
    if (threadIdx.x == 0) sItem[0] = atomicAdd(pv.queue, 1);
    int par = 0;
    while (true) {
        __syncthreads();
        const int q = sItem[par];
        if (q >= nItems) break;
        if (threadIdx.x == 0) sItem[par ^ 1] = atomicAdd(pv.queue, 1);      // next item, fetched while this one runs
        par ^= 1;
        int kind, ch;       // 0 = faces, 1 = cells
        if (q < lag) { kind = 0; ch = q; }
        else if (q < lag + 2 * (nCh - lag)) { const int t = q - lag; kind = t & 1; ch = kind ? (t >> 1) : lag + (t >> 1); }
        else { kind = 1; ch = nCh - lag + (q - (lag + 2 * (nCh - lag))); }
        if (kind == 0) {
            const int r0 = __ldg(&pv.ringOff[ch]), r1 = __ldg(&pv.ringOff[ch + 1]);
            if (r1 > r0) waitList(pv.doneC, pv.ringList, r0, r1, pv.epoch);
            const int f0 = __ldg(&pv.faceOff[ch]), f1 = __ldg(&pv.faceOff[ch + 1]);
            int f = f0 + (int)threadIdx.x;
            int P = 0, N = 0, flags = 0;
            int4 v = make_int4(0, 0, 0, 0);
            if (f < f1) { P = __ldg(&fv.own[f]); N = __ldg(&fv.nei[f]); flags = __ldg(&fv.flags[f]); v = __ldg(&fv.vtx[f]); }
            for (; f < f1; f += BLOCK) {
                const int Pc = P, Nc = N, flagsCur = flags;
                const int4 vc = v;
                const int fn = f + BLOCK;
                if (fn < f1) { P = __ldg(&fv.own[fn]); N = __ldg(&fv.nei[fn]); flags = __ldg(&fv.flags[fn]); v = __ldg(&fv.vtx[fn]); }
                faceFluxOne<false>(k, fv, sv, f, (size_t)((unsigned)f % (unsigned)sv.ringSize), Pc, Nc, flagsCur, vc, coMax, tauMin);
            }
            publish(pv.doneF, ch, pv.epoch);
        } else {
            waitList(pv.doneF, pv.depList, __ldg(&pv.depOff[ch]), __ldg(&pv.depOff[ch + 1]), pv.epoch);
            const int c0 = __ldg(&pv.cellOff[ch]), c1 = __ldg(&pv.cellOff[ch + 1]);
            for (int c = c0 + (int)threadIdx.x; c < c1; c += BLOCK) cellUpdateOne<W, true>(k, sv, fv.nI, c);
            publish(pv.doneC, ch, pv.epoch);
        }
    }
}

// boundary state after the cell update: the correctBoundaryConditions() sequence of QGDUEqn.H:51,88 ; QGDEEqn.H:50,75 ;
// hePsiQGDThermo.C:84-121 ; constScPrModel1.C:117-130 ; QGDFoam.C:155-156
__device__ __forceinline__ void bndClose(const Consts& k, const FaceView& fv, const SolverView& sv, const BndState& bs, int b,
                                         bool init, const RecA& cA, double aQGD)
{
    const int f = fv.nI + b;
    RecA a = bs.A[b];
    const bool fixU = bs.bcU[b] == QGD_BC_FIXED_VALUE, fixT = bs.bcT[b] == QGD_BC_FIXED_VALUE;
    // p_b as left by the last p.correctBoundaryConditions() (mid-step for qgdFlux patches)
    const double rhoOld = a.rho, pOld = init ? a.p : bs.pNew[b];
    double U[3] = {cA.Ux, cA.Uy, cA.Uz};
    if (fixU) { U[0] = bs.bvU[3 * (size_t)b]; U[1] = bs.bvU[3 * (size_t)b + 1]; U[2] = bs.bvU[3 * (size_t)b + 2]; }
    else if (bs.bcU[b] == QGD_BC_SLIP) {        // [OF-v2312 basicSymmetryFvPatchField::evaluate] U_b = U_P - n (n . U_P)
        const double ms = fv.magSf[f];
        const double n[3] = {fv.Sf[f] / ms, fv.Sf[(size_t)fv.fs + f] / ms, fv.Sf[2 * (size_t)fv.fs + f] / ms};
        const double un = n[0] * U[0] + n[1] * U[1] + n[2] * U[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) U[j] = U[j] - n[j] * un;
    }
    double T, e;
    if (fixT) { T = bs.bvT[b]; e = thermoEs(k, T); }                     // fixedEnergy ; hePsiQGDThermo.C:93-105
    else { e = cA.e; T = thermoTHE(k, e, init ? cA.T : a.T); }           // gradientEnergy (gradient 0) ; :109-119
    const double psi = 1.0 / (k.R * T);
    const double rhoB = init ? psi * pOld : rhoOld;                      // rho_b is refreshed last (QGDFoam.C:156)
    const double rhoU[3] = {rhoB * U[0], rhoB * U[1], rhoB * U[2]};      // QGDUEqn.H:88-89
    const double rhoE = rhoB * (e + 0.5 * (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]));   // QGDEEqn.H:75-76
    const double c = sqrt(k.gamma / psi);
    const double hf = fv.hf[f];
    const bool tauByU = (k.model == 1) && !init;                         // constScPrModel1n.C:119-126 (boundary part)
    const double tauB = tauByU ? aQGD * hf / (sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]) + c)
                               : aQGD * hf / c;                          // hQGD_b = hQGDf_b  QGDCoeffs.C:373
    const double muQGD = pOld * k.ScB * tauB;                            // constScPrModel1.C:121-124
    const double muT = muMol(k, T);
    const double mu = muT + muQGD;
    const double alpha = alphahMol(k, muT) + muQGD / k.PrQGD;
    const double aByC = tauByU ? tauB : ((k.model == 1) ? aQGD : aQGD / c);      // slot: see Consts::tauMode / cellThermo
    const double tauFace = tauByU ? tauB : ((k.model == 1) ? aQGD * hf / c : aByC * hf);   // tauQGDf on this boundary face after correct()
    if (bs.tauOutB) bs.tauOutB[b] = (k.model == 2) ? tauB + muT / (pOld * k.ScB) : tauB;
    // p.correctBoundaryConditions()   QGDFoam.C:155 ; qgdFluxFvPatchScalarField.C:159-208
    double p;
    const int bcP = bs.bcP[b];
    if (bcP == QGD_BC_FIXED_VALUE) p = bs.bvP[b];
    else if (bcP == QGD_BC_ZERO_GRADIENT) p = cA.p;
    else {
        double grad = bs.pGrad[b];
        if (!init) { grad = -(bs.phiw[b] / tauFace / fv.magSf[f]); bs.pGrad[b] = grad; }
        p = cA.p + grad / fv.dC[f];
    }
    const double rhoNew = init ? rhoB : psi * p;                         // QGDFoam.C:156
    a.rho = rhoNew; a.Ux = U[0]; a.Uy = U[1]; a.Uz = U[2]; a.e = e; a.p = p; a.T = T;
    a.H = (rhoE + p) / rhoNew;                                           // updateFields.H:71 boundary part
    bs.A[b] = a;
    bs.B[b] = RecB{rhoU[0], rhoU[1], rhoU[2], rhoE, c, mu, k.alphaEffGamma ? k.gamma * alpha : alpha, aByC};
    bs.psi[b] = psi;
    bs.pNew[b] = p;
}

__global__ void k_bnd_post(Consts k, FaceView fv, SolverView sv, BndState bs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    if (fv.bKind[b] == QGD_PATCH_EMPTY) return;
    const int P = fv.own[fv.nI + b];
    if (P >= sv.nOwned) return;
    bndClose(k, fv, sv, bs, b, false, loadA(sv, P), sv.aQGD[P]);
}

// wedge velocity condition [OF-v2312 wedgeFvPatchField::evaluate]: U_b = transform(faceT, U_P).  bndClose has just closed the face
// like a zeroGradient one (U_b = U_P, rhoU_b = rho_b U_P); the condition is a rotation, so it is applied to the two stored vectors
// afterwards (|U| and with it rhoE_b, H_b are unchanged) - the closing kernels themselves stay as they are.
__global__ void k_wedge_bnd(FaceView fv, SolverView sv, BndState bs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    if (fv.bKind[b] != QGD_PATCH_WEDGE || bs.bcU[b] != QGD_BC_WEDGE) return;
    const int f = fv.nI + b;
    if (fv.own[f] >= sv.nOwned) return;
    const double ms = fv.magSf[f];
    const double n[3] = {fv.Sf[f] / ms, fv.Sf[(size_t)fv.fs + f] / ms, fv.Sf[2 * (size_t)fv.fs + f] / ms};
    double Tw[9];
    wedgeFaceT(n, Tw);
    RecA a = bs.A[b];
    RecB bb = bs.B[b];
    const double u[3] = {a.Ux, a.Uy, a.Uz}, r[3] = {bb.rhoUx, bb.rhoUy, bb.rhoUz};
    a.Ux = Tw[0] * u[0] + Tw[1] * u[1] + Tw[2] * u[2];
    a.Uy = Tw[3] * u[0] + Tw[4] * u[1] + Tw[5] * u[2];
    a.Uz = Tw[6] * u[0] + Tw[7] * u[1] + Tw[8] * u[2];
    bb.rhoUx = Tw[0] * r[0] + Tw[1] * r[1] + Tw[2] * r[2];
    bb.rhoUy = Tw[3] * r[0] + Tw[4] * r[1] + Tw[5] * r[2];
    bb.rhoUz = Tw[6] * r[0] + Tw[7] * r[1] + Tw[8] * r[2];
    bs.A[b] = a;
    bs.B[b] = bb;
}

// ---- initialisation: QGDFoam/createFields.H:3-87 on the device
__global__ void k_init_cells(Consts k, SolverView sv, const double* __restrict__ U0, const double* __restrict__ T0,
                             const double* __restrict__ p0)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nCells) return;
    const double p = p0[c];
    double T = T0[c];
    const double e = thermoEs(k, T);                       // heThermo::init  he = HE(p,T)
    T = thermoTHE(k, e, T);                                // hePsiQGDThermo ctor: calculate()
    T = thermoTHE(k, e, T);                                // createFields.H:8  thermo.correct()
    const double psi = 1.0 / (k.R * T);
    const double rho = psi * p;                            // psiThermo::rho()
    const double U[3] = {U0[3 * (size_t)c], U0[3 * (size_t)c + 1], U0[3 * (size_t)c + 2]};
    const double rhoU[3] = {rho * U[0], rho * U[1], rho * U[2]};
    const double rhoE = rho * e + rho * 0.5 * (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
    // cellThermo recomputes p = rho/psi; at start-up p is the field as read -> store explicitly afterwards
    cellThermo(k, rho, U, rhoU, rhoE, e, p, T, sv.aQGD[c], sv.hQGD[c], sv, c, false);    // "U" not registered yet
    sv.S[5 * (size_t)sv.nCells + c] = p;
    sv.S[7 * (size_t)sv.nCells + c] = (rhoE + p) / rho;
}

__global__ void k_init_bnd(Consts k, FaceView fv, SolverView sv, BndState bs, const double* __restrict__ T0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fv.nB) return;
    bs.pGrad[b] = 0.0; bs.phiw[b] = 0.0; bs.pNew[b] = 0.0; bs.psi[b] = 0.0;
    if (fv.bKind[b] == QGD_PATCH_EMPTY) { bs.A[b] = RecA{1, 0, 0, 0, 0, 0, 1, 0}; bs.B[b] = RecB{0, 0, 0, 0, 1, 0, 0, 0}; return; }
    const int P = fv.own[fv.nI + b];
    RecA cA = loadA(sv, P);
    cA.T = T0[P];
    // boundary values of the fields as read: p_b (value / internal), then the common closing sequence
    RecA a{};
    a.p = (bs.bcP[b] == QGD_BC_FIXED_VALUE) ? bs.bvP[b] : cA.p;
    bs.A[b] = a;
    bndClose(k, fv, sv, bs, b, true, cA, sv.aQGD[P]);
}


// ============================================================================ implicit-diffusion branch
// QGDFoam/updateFluxes.H:107-111, QGDUEqn.H:54-75, QGDEEqn.H:53-64.  Not the headline path: written for correctness
// with plain thread-per-item kernels; the three U solves and the e solve run in the persistent PCG kernel (qgd_pcg.cu).

// iterate over the faces of cell c in ascending polyMesh order: fn(deviceFace, isNeighbourSide)
template <class F>
__device__ __forceinline__ void forCellFacesS(const SolverView& sv, int c, F fn)
{
    int last = -1;
    for (int j = 0; j < sv.cfEllW; ++j) {
        const int e = __ldg(&sv.cfEll[(size_t)j * sv.nCells + c]);
        last = e;
        if (e >= 0) fn(e >> 1, e & 1);
    }
    if (last >= 0)
        for (int t = __ldg(&sv.cfTailOff[c]); t < __ldg(&sv.cfTailOff[c + 1]); ++t) {
            const int e = __ldg(&sv.cfTailEnc[t]);
            fn(e >> 1, e & 1);
        }
}

// ---- varScModel6::correct varScModel6.C:210-269 | varScModel7::correct varScModel7.C:176-254 : ScQGD of a cell from
// the pressure jumps over its faces, cSc1 |sum +-dpf| / (sum pf / n) with pf = linearInterpolate(p) and
// dpf = fvc::snGrad(p)/deltaCoeffs, summed like mesh.cells()[c]: the faces the cell owns, then the faces it is the
// neighbour of, each ascending; empty / wedge faces skipped.  p is the OLD pressure (thermo.correct() runs before
// p = rho/psi, QGDFoam.C:149-154), so the kernel runs before the cell update overwrites it.  INIT: p as read (p0), the
// boundary value from the BC of the read field.
template <bool INIT>
__global__ void __launch_bounds__(kBlock) k_varsc(Consts k, FaceView fv, SolverView sv, BndState bs, const double* __restrict__ p)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nOwned) return;
    const double pc = p[c];
    double sumDpF = 0.0, sumpf = 0.0, n = 0.0;
    for (int pass = 0; pass < 2; ++pass)
        forCellFacesS(sv, c, [&](int f, int side) {
            if (side != pass) return;
            if (f < fv.nI) {
                const int o = side ? __ldg(&fv.own[f]) : __ldg(&fv.nei[f]);
                const double po = p[o], w = __ldg(&fv.w[f]);
                sumpf += side ? (w * (po - pc) + pc) : (w * (pc - po) + po);           // w (p_P - p_N) + p_N
                const double d = (side ? __ldg(&fv.ndC[f]) * (pc - po) : __ldg(&fv.ndC[f]) * (po - pc)) / __ldg(&fv.dC[f]);
                if (side) sumDpF -= d; else sumDpF += d;
                n = n + 1;
            } else {
                const int b = f - fv.nI;
                const int kind = __ldg(&fv.bKind[b]);
                if (kind == QGD_PATCH_EMPTY || kind == QGD_PATCH_WEDGE) return;
                const int bc = bs.bcP[b];
                const double dC = __ldg(&fv.dC[f]);
                const double pb = INIT ? (bc == QGD_BC_FIXED_VALUE ? bs.bvP[b] : pc) : bs.pNew[b];
                const double sn = bc == QGD_BC_FIXED_VALUE ? dC * (pb - pc) : (bc == QGD_BC_ZERO_GRADIENT ? 0.0 : (INIT ? 0.0 : bs.pGrad[b]));
                sumDpF += sn / dC;
                sumpf += pb;
                n = n + 1;
            }
        });
    sumpf /= n;
    double sc = (k.varSc == 7 ? k.cSc1 : 1.0) * (fabs(sumDpF) / sumpf);
    if (k.varSc == 7) {
        if (k.minSc >= 0) sc = fmax(sc, k.minSc);
        if (k.maxSc >= 0) sc = fmin(sc, k.maxSc);
        if (sv.scConst && sv.scConst[c]) sc = k.ScQGD;
    }
    sv.scVar[c] = sc;
}

// boundary value of U.  newU: after the solve (fixedValue keeps its value, zeroGradient follows the cell)
__device__ __forceinline__ void bndU(const SolverView& sv, const BndState& bs, int b, int P, bool newU, double (&u)[3])
{
    if (!newU) { const RecA a = bs.A[b]; u[0] = a.Ux; u[1] = a.Uy; u[2] = a.Uz; return; }
    if (bs.bcU[b] == QGD_BC_FIXED_VALUE) { u[0] = bs.bvU[3 * (size_t)b]; u[1] = bs.bvU[3 * (size_t)b + 1]; u[2] = bs.bvU[3 * (size_t)b + 2]; }
    else { const size_t n = sv.nCells; u[0] = sv.S[n + P]; u[1] = sv.S[2 * n + P]; u[2] = sv.S[3 * n + P]; }
}

// [OF-v2312] fvc::grad(U), Gauss linear, cell values: (1/V) sum_f +-Sf (x) U_f
__global__ void __launch_bounds__(kBlock) k_gauss_gradU(FaceView fv, SolverView sv, BndState bs, double* __restrict__ GU, int newU)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nCells) return;
    const size_t n = sv.nCells;
    const double uc[3] = {sv.S[n + c], sv.S[2 * n + c], sv.S[3 * n + c]};
    double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    forCellFacesS(sv, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        double uf[3];
        if (f < fv.nI) {
            const int o = side ? __ldg(&fv.own[f]) : __ldg(&fv.nei[f]);
            const double w = __ldg(&fv.w[f]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double uo = sv.S[(1 + j) * n + o];
                uf[j] = side ? (w * (uo - uc[j]) + uc[j]) : (w * (uc[j] - uo) + uo);
            }
        } else {
            const int b = f - fv.nI;
            if (__ldg(&fv.bKind[b]) == QGD_PATCH_EMPTY) return;
            bndU(sv, bs, b, c, newU != 0, uf);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double Si = __ldg(&fv.Sf[(size_t)i * fv.fs + f]);
#pragma unroll
            for (int j = 0; j < 3; ++j) G[3 * i + j] += sgn * (Si * uf[j]);
        }
    });
    const double V = __ldg(&sv.V[c]);
#pragma unroll
    for (int t = 0; t < 9; ++t) GU[t * n + c] = G[t] / V;
}

// mu * dev2(T(g)) ; dev2(A) = A - (2/3) tr(A) I
__device__ __forceinline__ void muDev2T(double mu, const double (&g)[9], double (&o)[9])
{
    const double tr = g[0] + g[4] + g[8];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[3 * i + j] = mu * (g[3 * j + i] - (2.0 / 3.0) * tr * (i == j ? 1.0 : 0.0));
}
__device__ __forceinline__ void loadG(const double* GU, size_t n, int c, double (&g)[9])
{
#pragma unroll
    for (int t = 0; t < 9; ++t) g[t] = GU[t * n + c];
}
// boundary value of fvc::grad(U): gaussGrad::correctBoundaryConditions  [OF-v2312]
__device__ __forceinline__ void bndGrad(const double (&gP)[9], const double (&nrm)[3], const double (&sn)[3], double (&gB)[9])
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double nG = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) nG += nrm[i] * gP[3 * i + j];
#pragma unroll
        for (int i = 0; i < 3; ++i) gB[3 * i + j] = gP[3 * i + j] + nrm[i] * (sn[j] - nG);
    }
}

// phase "diff": per face phiTauMC (3) and the Laplacian coefficients of the U and e equations
__global__ void __launch_bounds__(kBlock) k_face_diff(FaceView fv, SolverView sv, BndState bs, ImplicitView iv)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fv.nF) return;
    const size_t n = sv.nCells, nF = fv.nF;
    double Sf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Sf[i] = fv.Sf[(size_t)i * fv.fs + f];
    const double ms = fv.magSf[f], nd = fv.ndC[f];
    double tMC[9], muf, alf;
    if (f < fv.nI) {
        const int P = fv.own[f], N = fv.nei[f];
        const double w = fv.w[f];
        const double muP = sv.S[13 * n + P], muN = sv.S[13 * n + N], aP = sv.S[14 * n + P], aN = sv.S[14 * n + N];
        muf = w * (muP - muN) + muN;
        alf = w * (aP - aN) + aN;
        double gP[9], gN[9], tP[9], tN[9];
        loadG(iv.GU0, n, P, gP); loadG(iv.GU0, n, N, gN);
        muDev2T(muP, gP, tP); muDev2T(muN, gN, tN);
#pragma unroll
        for (int t = 0; t < 9; ++t) tMC[t] = w * (tP[t] - tN[t]) + tN[t];
    } else {
        const int b = f - fv.nI;
        if (fv.bKind[b] == QGD_PATCH_EMPTY) {
            iv.FT[f] = 0.0; iv.FT[nF + f] = 0.0; iv.FT[2 * nF + f] = 0.0; iv.aU[f] = 0.0; iv.aE[f] = 0.0;
            return;
        }
        const int P = fv.own[f];
        const RecA a = bs.A[b];
        const RecB bb = bs.B[b];
        muf = bb.mu; alf = bb.alphaEff;
        const double nrm[3] = {Sf[0] / ms, Sf[1] / ms, Sf[2] / ms};
        const bool fixU = bs.bcU[b] == QGD_BC_FIXED_VALUE;
        const double delta = fv.dC[f];
        const double sn[3] = {fixU ? delta * (a.Ux - sv.S[n + P]) : 0.0, fixU ? delta * (a.Uy - sv.S[2 * n + P]) : 0.0,
                              fixU ? delta * (a.Uz - sv.S[3 * n + P]) : 0.0};
        double gP[9], gB[9];
        loadG(iv.GU0, n, P, gP);
        bndGrad(gP, nrm, sn, gB);
        muDev2T(bb.mu, gB, tMC);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) iv.FT[j * nF + f] = Sf[0] * tMC[j] + Sf[1] * tMC[3 + j] + Sf[2] * tMC[6 + j];
    iv.aU[f] = nd * (muf * ms);
    iv.aE[f] = nd * (alf * ms);
}

// phase A: rho, rhoU (explicit), U* = rhoU/rho ; U-equation diagonal and sources
template <int W>
__global__ void __launch_bounds__(kBlock) k_cell_implA(Consts k, FaceView fv, SolverView sv, BndState bs, ImplicitView iv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nCells) return;
    const size_t n = sv.nCells, nF = fv.nF;
    const RecA a = loadA(sv, c);
    if (c >= sv.nOwned) {        // halo copy (multi-GPU): only the old U / rho that k_face_sigma interpolates; the row belongs to another rank
        iv.old[c] = a.Ux; iv.old[n + c] = a.Uy; iv.old[2 * n + c] = a.Uz; iv.old[3 * n + c] = a.rho;
        return;
    }
    const RecB b = loadB(sv, c);
    double sm = 0.0, su[3] = {0, 0, 0}, st[3] = {0, 0, 0}, dL = 0.0, bI = 0.0, bB[3] = {0, 0, 0};
    forCellFacesS(sv, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        double fm, f0, f1, f2, fe;
        loadFlux<false>(sv, fv.nI, f, fm, f0, f1, f2, fe);
        sm += sgn * fm; su[0] += sgn * f0; su[1] += sgn * f1; su[2] += sgn * f2;
#pragma unroll
        for (int j = 0; j < 3; ++j) st[j] += sgn * iv.FT[j * nF + f];
        if (f < fv.nI) dL += iv.aU[f];
        else {
            const int bf = f - fv.nI;
            if (fv.bKind[bf] != QGD_PATCH_EMPTY && bs.bcU[bf] == QGD_BC_FIXED_VALUE) {      // internalCoeffs / boundaryCoeffs
                const double gS = iv.aU[f];
                bI += gS;
#pragma unroll
                for (int j = 0; j < 3; ++j) bB[j] += gS * bs.bvU[3 * (size_t)bf + j];
            }
        }
    });
    const double V = __ldg(&sv.V[c]);
    const double rDeltaT = 1.0 / sv.sc->dt;
    const double diag = rDeltaT * V;
    const bool src = sv.su != nullptr;
    const double rho = src ? (rDeltaT * a.rho * V - V * (sm / V) + sv.su[c]) / diag : (rDeltaT * a.rho * V - V * (sm / V)) / diag;   // QGDRhoEqn.H:40-47
    const double u0[3] = {a.Ux, a.Uy, a.Uz}, r0[3] = {b.rhoUx, b.rhoUy, b.rhoUz};
    iv.old[c] = a.Ux; iv.old[n + c] = a.Uy; iv.old[2 * n + c] = a.Uz; iv.old[3 * n + c] = a.rho;
    sv.S[c] = rho;
    iv.diagU[c] = rDeltaT * rho * V + dL + bI;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double rhoU = (rDeltaT * r0[j] * V - V * (su[j] / V)) / diag;                 // QGDUEqn.H:36-45
        const double Us = rhoU / rho;                                                       // :48-50
        sv.S[(8 + j) * n + c] = rhoU;
        sv.S[(1 + j) * n + c] = Us;
        // :56-62  fvm::ddt(rho,U) - fvc::ddt(rho,U) - fvm::laplacian(muf,U) - fvc::div(phiTauMC) == 0
        double bj = rDeltaT * a.rho * u0[j] * V + V * (rDeltaT * (rho * Us - a.rho * u0[j])) + V * (st[j] / V);
        if (src) bj += sv.su[(1 + j) * n + c];                                              // == rhoUSu  :62
        iv.bU[j * n + c] = bj + bB[j];
    }
}

// phase "sigma": phiSigmaDotU = Sf & ((muf*linearInterpolate(fvc::grad(U)) + tauMC) & Uf)   QGDUEqn.H:72-74
__global__ void __launch_bounds__(kBlock) k_face_sigma(FaceView fv, SolverView sv, BndState bs, ImplicitView iv)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fv.nF) return;
    const size_t n = sv.nCells;
    double Sf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Sf[i] = fv.Sf[(size_t)i * fv.fs + f];
    const double ms = fv.magSf[f];
    double T[9], Uf[3];
    if (f < fv.nI) {
        const int P = fv.own[f], N = fv.nei[f];
        const double w = fv.w[f];
        const double muP = sv.S[13 * n + P], muN = sv.S[13 * n + N];
        const double muf = w * (muP - muN) + muN;
        double g0P[9], g0N[9], g1P[9], g1N[9], tP[9], tN[9];
        loadG(iv.GU0, n, P, g0P); loadG(iv.GU0, n, N, g0N); loadG(iv.GU1, n, P, g1P); loadG(iv.GU1, n, N, g1N);
        muDev2T(muP, g0P, tP); muDev2T(muN, g0N, tN);
#pragma unroll
        for (int t = 0; t < 9; ++t) T[t] = muf * (w * (g1P[t] - g1N[t]) + g1N[t]) + (w * (tP[t] - tN[t]) + tN[t]);
#pragma unroll
        for (int j = 0; j < 3; ++j) { const double uP = iv.old[j * n + P], uN = iv.old[j * n + N]; Uf[j] = w * (uP - uN) + uN; }
    } else {
        const int b = f - fv.nI;
        if (fv.bKind[b] == QGD_PATCH_EMPTY) { iv.Fs[f] = 0.0; return; }
        const int P = fv.own[f];
        const RecA a = bs.A[b];
        const RecB bb = bs.B[b];
        const double nrm[3] = {Sf[0] / ms, Sf[1] / ms, Sf[2] / ms};
        const bool fixU = bs.bcU[b] == QGD_BC_FIXED_VALUE;
        const double delta = fv.dC[f];
        const double uOld[3] = {iv.old[P], iv.old[n + P], iv.old[2 * n + P]};
        const double ub[3] = {a.Ux, a.Uy, a.Uz};
        double un[3];
        bndU(sv, bs, b, P, true, un);
        double sn0[3], sn1[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            sn0[j] = fixU ? delta * (ub[j] - uOld[j]) : 0.0;
            sn1[j] = fixU ? delta * (un[j] - sv.S[(1 + j) * n + P]) : 0.0;
        }
        double g0[9], g1[9], g0B[9], g1B[9], tB[9];
        loadG(iv.GU0, n, P, g0); loadG(iv.GU1, n, P, g1);
        bndGrad(g0, nrm, sn0, g0B); bndGrad(g1, nrm, sn1, g1B);
        muDev2T(bb.mu, g0B, tB);
#pragma unroll
        for (int t = 0; t < 9; ++t) T[t] = bb.mu * g1B[t] + tB[t];
#pragma unroll
        for (int j = 0; j < 3; ++j) Uf[j] = ub[j];
    }
    double sg[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) sg[i] = T[3 * i] * Uf[0] + T[3 * i + 1] * Uf[1] + T[3 * i + 2] * Uf[2];
    iv.Fs[f] = Sf[0] * sg[0] + Sf[1] * sg[1] + Sf[2] * sg[2];
}

// phase B: rhoU = rho*U ; rhoE (explicit, with phiSigmaDotU) ; e* ; e-equation diagonal and source
__global__ void __launch_bounds__(kBlock) k_cell_implB(Consts k, FaceView fv, SolverView sv, BndState bs, ImplicitView iv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nOwned) return;
    const size_t n = sv.nCells;
    double se = 0.0, ss = 0.0, dL = 0.0, bI = 0.0, bB = 0.0;
    forCellFacesS(sv, c, [&](int f, int side) {
        const double sgn = side ? -1.0 : 1.0;
        double fm, f0, f1, f2, fe;
        loadFlux<false>(sv, fv.nI, f, fm, f0, f1, f2, fe);
        se += sgn * fe; ss += sgn * iv.Fs[f];
        if (f < fv.nI) dL += iv.aE[f];
        else {
            const int bf = f - fv.nI;
            if (fv.bKind[bf] != QGD_PATCH_EMPTY && bs.bcT[bf] == QGD_BC_FIXED_VALUE) {      // fixedEnergy
                bI += iv.aE[f];
                bB += iv.aE[f] * thermoEs(k, bs.bvT[bf]);
            }
        }
    });
    const double V = __ldg(&sv.V[c]);
    const double rDeltaT = 1.0 / sv.sc->dt;
    const double rho = sv.S[c], rho0 = iv.old[3 * n + c], e0 = sv.S[4 * n + c], rhoE0 = sv.S[11 * n + c];
    const double U[3] = {sv.S[n + c], sv.S[2 * n + c], sv.S[3 * n + c]};
#pragma unroll
    for (int j = 0; j < 3; ++j) sv.S[(8 + j) * n + c] = rho * U[j];                         // QGDUEqn.H:70
    const double rhoE = (rDeltaT * rhoE0 * V - V * (se / V) + V * (ss / V)) / (rDeltaT * V);   // QGDEEqn.H:37-46
    const double es = rhoE / rho - 0.5 * (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);         // :49
    sv.S[11 * n + c] = rhoE;
    sv.S[4 * n + c] = es;
    iv.diagE[c] = rDeltaT * rho * V + dL + bI;                                              // :55-60
    double be = rDeltaT * rho0 * e0 * V + V * (rDeltaT * (rho * es - rho0 * e0));
    if (sv.su) be += sv.su[4 * n + c];                                                      // == rhoESu  QGDEEqn.H:60
    iv.bE[c] = be + bB;
}

// phase C: rhoE = rho (e + 0.5 |U|^2) ; thermo.correct() ; p = rho/psi   (QGDEEqn.H:63, QGDFoam.C:149-154)
__global__ void __launch_bounds__(kBlock) k_cell_implC(Consts k, SolverView sv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sv.nOwned) return;
    const size_t n = sv.nCells;
    const double rho = sv.S[c], e = sv.S[4 * n + c], pOld = sv.S[5 * n + c], TOld = sv.S[6 * n + c];
    const double U[3] = {sv.S[n + c], sv.S[2 * n + c], sv.S[3 * n + c]};
    const double rhoU[3] = {rho * U[0], rho * U[1], rho * U[2]};
    const double rhoE = rho * (e + 0.5 * (U[0] * U[0] + U[1] * U[1] + U[2] * U[2]));
    if (e <= 0.0 || rho <= 0.0) atomicCAS(&sv.sc->guardStep, 0, sv.sc->stepIndex);
    cellThermo(k, rho, U, rhoU, rhoE, e, pOld, TOld, __ldg(&sv.aQGD[c]), __ldg(&sv.hQGD[c]), sv, c);
}

} // namespace

// ============================================================================ launchers
static inline int nblk(long n, int b = kBlock) { return (int)((n + b - 1) / b); }

namespace {
int g_faceTma = 1;          // QGD_FACE_TMA=0 selects the register-prefetch kernel instead of the TMA-staged one
int g_faceHint = 3;         // QGD_FACE_L2HINT bit 0: face constants / fluxes evict_first, bit 1: cell / point gathers evict_last
template <bool ADJUST> void launchTma(int hint, int grid, cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv)
{
    const int sm = 2 * TmaStage<ADJUST>::kBytes;
    switch (hint & 3) {
    case 1: k_face_flux_tma<ADJUST, 1><<<grid, kTmaTile, sm, st>>>(c, fv, sv); break;
    case 2: k_face_flux_tma<ADJUST, 2><<<grid, kTmaTile, sm, st>>>(c, fv, sv); break;
    case 3: k_face_flux_tma<ADJUST, 3><<<grid, kTmaTile, sm, st>>>(c, fv, sv); break;
    default: k_face_flux_tma<ADJUST, 0><<<grid, kTmaTile, sm, st>>>(c, fv, sv); break;
    }
}
template <bool ADJUST, int HINT> void tmaSmemAttr()
{
    cudaFuncSetAttribute(k_face_flux_tma<ADJUST, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TmaStage<ADJUST>::kBytes);
}
template <bool ADJUST> int tmaGridOf()
{
    int dev = 0, sms = 148, perSM = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    tmaSmemAttr<ADJUST, 0>(); tmaSmemAttr<ADJUST, 1>(); tmaSmemAttr<ADJUST, 2>(); tmaSmemAttr<ADJUST, 3>();
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_face_flux_tma<ADJUST, 3>, kTmaTile, 2 * TmaStage<ADJUST>::kBytes);
    return sms * (perSM < 1 ? 1 : perSM);
}
int faceTmaGrid(bool adjust)
{
    static int grid[2] = {0, 0};
    int& g = grid[adjust ? 1 : 0];
    if (!g) g = adjust ? tmaGridOf<true>() : tmaGridOf<false>();
    return g;
}
}

void warmFaceKernel(bool adjust) { (void)faceTmaGrid(adjust); }
void setFaceTma(int on) { g_faceTma = on ? 1 : 0; }
void setFaceL2Hint(int bits) { g_faceHint = bits & 3; }

int faceKernelGrid()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int perSM = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_face_flux<false>, kBlock, 0);
    if (perSM < 1) perSM = 1;
    return sms * perSM;
}

void launchPointGather(cudaStream_t st, int K, const qgd_mesh& m, const double* cell, const double* bnd, double* pts)
{
    const int nP = m.h.nPoints, nPP = (int)m.h.patchPoints.size();
#define QGD_PG(KK)                                                                                                          \
    k_point_gather_generic<KK><<<nblk(nP), kBlock, 0, st>>>(nP, m.pcEllW, m.pcEll.p, m.pcEllWt.p, m.pcCount.p, m.pcTailOff.p, m.pcTailCell.p, m.pcTailW.p, cell, pts); \
    if (nPP) k_patch_point_gather_generic<KK><<<nblk(nPP), kBlock, 0, st>>>(nPP, m.patchPoints.p, m.ppOff.p, m.ppFace.p, m.ppW.p, bnd, pts);
    switch (K) {
        case 1: QGD_PG(1) break;
        case 3: QGD_PG(3) break;
        case 9: QGD_PG(9) break;
        default: throw Error(QGD_ERR_INVALID, "fvsc: unsupported number of components");
    }
#undef QGD_PG
    const int nW = (int)m.h.wedgePts.size();        // pointConstraints on the vertices of wedge patches (vectors, tensors)
    if (nW && K == 3) k_wedge_points_generic<3><<<nblk(nW), kBlock, 0, st>>>(nW, m.wedgePts.p, m.wedgeR.p, pts);
    if (nW && K == 9) k_wedge_points_generic<9><<<nblk(nW), kBlock, 0, st>>>(nW, m.wedgePts.p, m.wedgeR.p, pts);
    QGD_CUDA(cudaGetLastError());
}

void launchFvscGrad(cudaStream_t st, int K, const FaceView& fv, const double* cell, const double* pts, const double* bnd,
                    const double* bsg, const double* nbr, double* out)
{
    if (K == 1) k_fvsc_grad<1><<<nblk(fv.nF), kBlock, 0, st>>>(fv, cell, pts, bnd, bsg, nbr, out);
    else if (K == 3) k_fvsc_grad<3><<<nblk(fv.nF), kBlock, 0, st>>>(fv, cell, pts, bnd, bsg, nbr, out);
    else throw Error(QGD_ERR_INVALID, "fvsc::grad: ncmpt must be 1 or 3");
    QGD_CUDA(cudaGetLastError());
}

void launchFvscDiv(cudaStream_t st, int K, const FaceView& fv, const double* cell, const double* pts, const double* bnd,
                   const double* bsg, const double* nbr, double* out)
{
    if (K == 3) k_fvsc_div<3><<<nblk(fv.nF), kBlock, 0, st>>>(fv, cell, pts, bnd, bsg, nbr, out);
    else if (K == 9) k_fvsc_div<9><<<nblk(fv.nF), kBlock, 0, st>>>(fv, cell, pts, bnd, bsg, nbr, out);
    else throw Error(QGD_ERR_INVALID, "fvsc::div: ncmpt must be 3 or 9");
    QGD_CUDA(cudaGetLastError());
}

void launchInit(cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
                const double* U0, const double* T0, const double* p0, bool wedge)
{
    if (c.varSc) k_varsc<true><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, fv, sv, bs, p0);
    k_init_cells<<<nblk(sv.nCells), kBlock, 0, st>>>(c, sv, U0, T0, p0);
    if (fv.nB) k_init_bnd<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs, T0);
    if (fv.nB && wedge) k_wedge_bnd<<<nblk(fv.nB), kBlock, 0, st>>>(fv, sv, bs);
    QGD_CUDA(cudaGetLastError());
}

// the TMA-staged kernel needs one full tile; the SoA columns are 128-B aligned for any face count (FaceView::fs)
bool faceKernelIsTma(const FaceView& fv) { return g_faceTma && fv.lsqW == 0 && fv.nIActive >= kTmaTile; }
const char* faceKernelName(const FaceView& fv)
{
    if (fv.lsqW > 0) return "k_face_flux_lsq";
    return faceKernelIsTma(fv) ? "k_face_flux_tma" : "k_face_flux";
}
int faceKernelL2Hint() { return g_faceHint; }

// internal-face kernel of the two-kernel step form: leastSquares | TMA-staged | register-prefetch variants
static void launchFaceKernel(cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv, bool adjust, int gridFaces)
{
    if (fv.lsqW > 0) {
        const int gl = std::min(2 * 148, nblk(fv.nIActive, 256));
        if (adjust) k_face_flux_lsq<true><<<gl, 256, 0, st>>>(c, fv, sv);
        else k_face_flux_lsq<false><<<gl, 256, 0, st>>>(c, fv, sv);
    } else if (faceKernelIsTma(fv)) {
        const int gridT = std::min(faceTmaGrid(adjust), fv.nIActive / kTmaTile);
        if (adjust) launchTma<true>(g_faceHint, gridT, st, c, fv, sv);
        else launchTma<false>(g_faceHint, gridT, st, c, fv, sv);
    } else {
        const int grid = std::min(gridFaces, nblk(fv.nIActive, kBlock));
        if (adjust) k_face_flux<true><<<grid, kBlock, 0, st>>>(c, fv, sv);
        else k_face_flux<false><<<grid, kBlock, 0, st>>>(c, fv, sv);
    }
}

namespace {
template <int W> void launchPipe(cudaStream_t st, int grid, const Consts& c, const FaceView& fv, const SolverView& sv, const PipeView& pv)
{
    k_face_cell_pipeline<W, 256, 2><<<grid, 256, 0, st>>>(c, fv, sv, pv);
}
template <int W> int pipeGrid()
{
    int dev = 0, sms = 148, perSM = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_face_cell_pipeline<W, 256, 2>, 256, 0);
    return sms * (perSM < 1 ? 1 : perSM);
}
}

// resident CTAs of the pipelined kernel: the persistent grid must be fully co-resident (its items wait on each other)
int pipelineKernelGrid(int cfEllW) { return cfEllW == 4 ? pipeGrid<4>() : (cfEllW == 6 ? pipeGrid<6>() : pipeGrid<8>()); }

int launchStep(cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
               bool anyQgdFlux, int gridFaces, bool adjust, cudaEvent_t* ev, const StepHooks* hooks, const PipeView* pipe, int gridPipe,
               const StepFork* fork, const WedgeView* wedge)
{
    int n = 0;
    const bool pointsNeeded = !c.reducedScheme;
    const bool usePipe = pipe && !adjust && !c.varSc;
    // boundary kernels beside the point gather (see StepFork); the qgdFlux mid-step sequence and the pipelined form stay in line
    const bool forked = fork && fv.nB && !anyQgdFlux && !usePipe;
    cudaStream_t sb = forked ? fork->side : st;
    auto points = [&](const int* list, int cnt) {
        if (cnt <= 0) return;
        if (sv.pcEllW == 4) k_points<4><<<nblk(cnt), kBlock, 0, st>>>(sv, list, cnt);
        else if (sv.pcEllW == 6) k_points<6><<<nblk(cnt), kBlock, 0, st>>>(sv, list, cnt);
        else k_points<8><<<nblk(cnt), kBlock, 0, st>>>(sv, list, cnt);
        ++n;
    };
    if (forked) {                    // the side stream starts from the state the main stream has reached (cell update / exchange done)
        cudaEventRecord(fork->evEntry, st);
        cudaStreamWaitEvent(sb, fork->evEntry, 0);
    }
    if (pointsNeeded) {
        if (ev) cudaEventRecord(ev[0], st);
        if (sv.ptsInterior) {
            points(sv.ptsInterior, sv.nPtsInterior);                 // overlaps the halo exchange of the previous step
            if (hooks && hooks->waitHalo) hooks->waitHalo(st);
            points(sv.ptsHalo, sv.nPtsHalo);
        } else {
            if (hooks && hooks->waitHalo) hooks->waitHalo(st);
            points(nullptr, sv.nPoints);
        }
        if (ev) cudaEventRecord(ev[1], st);
        if (forked && hooks && hooks->waitHalo) hooks->waitHalo(sb);
        if (sv.nPatchPoints) { k_patch_points<<<nblk(sv.nPatchPoints), kBlock, 0, sb>>>(sv, bs, 0); ++n; }
        if (wedge && wedge->n) { k_wedge_points<<<nblk(wedge->n), kBlock, 0, sb>>>(wedge->n, wedge->pts, wedge->nrm, sv.P, (size_t)sv.nPoints); ++n; }
        if (forked) cudaEventRecord(fork->evPatch, sb);
    } else if (hooks && hooks->waitHalo) {
        hooks->waitHalo(st);
        if (forked) hooks->waitHalo(sb);
    }
    if (anyQgdFlux) {                // a global decision (patch table): the exchange in midStep is matched by every neighbour rank,
        if (fv.nB) { k_bnd_pre<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs); ++n; }      // also by one that holds no boundary face
        if (hooks && hooks->midStep) hooks->midStep();
        if (fv.nB && pointsNeeded && sv.nPatchPoints) { k_patch_points<<<nblk(sv.nPatchPoints), kBlock, 0, st>>>(sv, bs, 1); ++n; }
    }
    if (fv.nB) { k_bnd_flux<<<nblk(fv.nB), kBlock, 0, sb>>>(c, fv, sv, bs); ++n; }
    if (forked) {
        cudaEventRecord(fork->evBndFlux, sb);
        if (pointsNeeded) cudaStreamWaitEvent(st, fork->evPatch, 0);      // the face kernel gathers boundary points
    }
    if (usePipe) {
        // fixed deltaT: faces and cells in one persistent kernel, fluxes stay in the L2-resident ring
        k_dt<<<1, 1, 0, st>>>(sv.sc, pipe->queue); ++n;
        if (ev) cudaEventRecord(ev[2], st);
        if (pipe->nChunks > 0) {
            if (sv.cfEllW == 4) launchPipe<4>(st, gridPipe, c, fv, sv, *pipe);
            else if (sv.cfEllW == 6) launchPipe<6>(st, gridPipe, c, fv, sv, *pipe);
            else launchPipe<8>(st, gridPipe, c, fv, sv, *pipe);
            ++n;
        }
        if (ev) { cudaEventRecord(ev[3], st); cudaEventRecord(ev[4], st); cudaEventRecord(ev[5], st); }
    } else {
        if (fv.nIActive) {
            if (ev) cudaEventRecord(ev[2], st);
            launchFaceKernel(st, c, fv, sv, adjust, gridFaces);
            ++n;
            if (ev) cudaEventRecord(ev[3], st);
        }
        if (forked) cudaStreamWaitEvent(st, fork->evBndFlux, 0);         // boundary fluxes + their Courant / tau contributions
        if (hooks && hooks->beforeDt) hooks->beforeDt();
        k_dt<<<1, 1, 0, st>>>(sv.sc, nullptr); ++n;
        if (c.varSc) { k_varsc<false><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, fv, sv, bs, sv.S + 5 * (size_t)sv.nCells); ++n; }
        if (ev) cudaEventRecord(ev[4], st);
        if (sv.cfEllW == 4) k_cell_update<4><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, sv, fv.nI);
        else if (sv.cfEllW == 6) k_cell_update<6><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, sv, fv.nI);
        else k_cell_update<8><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, sv, fv.nI);
        ++n;
        if (ev) cudaEventRecord(ev[5], st);
    }
    if (hooks && hooks->beforeBndPost) hooks->beforeBndPost();
    if (fv.nB) {
        if (forked && fork->postOnSide) {
            cudaEventRecord(fork->evCell, st);
            cudaStreamWaitEvent(sb, fork->evCell, 0);
            k_bnd_post<<<nblk(fv.nB), kBlock, 0, sb>>>(c, fv, sv, bs); ++n;
            if (wedge && wedge->n) { k_wedge_bnd<<<nblk(fv.nB), kBlock, 0, sb>>>(fv, sv, bs); ++n; }
            cudaEventRecord(fork->evBndPost, sb);
        } else {
            k_bnd_post<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs); ++n;
            if (wedge && wedge->n) { k_wedge_bnd<<<nblk(fv.nB), kBlock, 0, st>>>(fv, sv, bs); ++n; }
        }
    }
    return n;
}

int launchImplicitPhase(cudaStream_t st, int phase, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
                        const ImplicitView& iv, bool anyQgdFlux, int gridFaces, bool adjust, const StepHooks* hooks)
{
    int n = 0;
    const bool pointsNeeded = !c.reducedScheme;
    if (phase == 0) {          // updateFields + updateFluxes (reduced explicit fluxes), phiTauMC, rho/rhoU update, U system
        if (pointsNeeded) {
            if (sv.pcEllW == 4) k_points<4><<<nblk(sv.nPoints), kBlock, 0, st>>>(sv, nullptr, sv.nPoints);
            else if (sv.pcEllW == 6) k_points<6><<<nblk(sv.nPoints), kBlock, 0, st>>>(sv, nullptr, sv.nPoints);
            else k_points<8><<<nblk(sv.nPoints), kBlock, 0, st>>>(sv, nullptr, sv.nPoints);
            ++n;
            if (sv.nPatchPoints) { k_patch_points<<<nblk(sv.nPatchPoints), kBlock, 0, st>>>(sv, bs, 0); ++n; }
        }
        if (anyQgdFlux) {            // global decision, see launchStep
            if (fv.nB) { k_bnd_pre<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs); ++n; }
            if (hooks && hooks->midStep) hooks->midStep();
            if (fv.nB && pointsNeeded && sv.nPatchPoints) { k_patch_points<<<nblk(sv.nPatchPoints), kBlock, 0, st>>>(sv, bs, 1); ++n; }
        }
        k_gauss_gradU<<<nblk(sv.nCells), kBlock, 0, st>>>(fv, sv, bs, iv.GU0, 0); ++n;
        if (hooks && hooks->afterGrad) hooks->afterGrad(iv.GU0);
        if (fv.nB) { k_bnd_flux<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs); ++n; }
        if (fv.nIActive) {
            launchFaceKernel(st, c, fv, sv, adjust, gridFaces);
            ++n;
        }
        k_face_diff<<<nblk(fv.nF), kBlock, 0, st>>>(fv, sv, bs, iv); ++n;
        if (hooks && hooks->beforeDt) hooks->beforeDt();
        k_dt<<<1, 1, 0, st>>>(sv.sc, nullptr); ++n;
        if (sv.cfEllW == 4) k_cell_implA<4><<<nblk(sv.nCells), kBlock, 0, st>>>(c, fv, sv, bs, iv);
        else if (sv.cfEllW == 6) k_cell_implA<6><<<nblk(sv.nCells), kBlock, 0, st>>>(c, fv, sv, bs, iv);
        else k_cell_implA<8><<<nblk(sv.nCells), kBlock, 0, st>>>(c, fv, sv, bs, iv);
        ++n;
    } else if (phase == 1) {   // after the U solves: sigmaDotU, rhoE update, e system
        k_gauss_gradU<<<nblk(sv.nCells), kBlock, 0, st>>>(fv, sv, bs, iv.GU1, 1); ++n;
        if (hooks && hooks->afterGrad) hooks->afterGrad(iv.GU1);
        k_face_sigma<<<nblk(fv.nF), kBlock, 0, st>>>(fv, sv, bs, iv); ++n;
        k_cell_implB<<<nblk(sv.nOwned), kBlock, 0, st>>>(c, fv, sv, bs, iv); ++n;
    } else {                   // after the e solve: conserved variables, thermo, boundary state
        if (c.varSc) { k_varsc<false><<<nblk(sv.nOwned), kBlock, 0, st>>>(c, fv, sv, bs, sv.S + 5 * (size_t)sv.nCells); ++n; }
        k_cell_implC<<<nblk(sv.nOwned), kBlock, 0, st>>>(c, sv); ++n;
        if (hooks && hooks->beforeBndPost) hooks->beforeBndPost();
        if (fv.nB) { k_bnd_post<<<nblk(fv.nB), kBlock, 0, st>>>(c, fv, sv, bs); ++n; }
    }
    return n;
}


} // namespace qgd
