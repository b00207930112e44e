// ============================================================================
// CPU ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see qgd_oracle.hpp).
// Field-at-a-time restatement of the reference hot path; never linked or
// loaded by the product.
// ============================================================================
#include "qgd_oracle.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Vec = std::vector<double>;
using IVec = std::vector<int>;
struct V3 { double x, y, z; };
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline double mag(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 ld3(const double* p, long i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
static inline double cmp(V3 a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }

const double SMALL = 1e-15;   // OpenFOAM SMALL (double)

} // namespace

// QHDFoam state (oracle/qhd_oracle.inc)
struct QhdState {
    or_qhd_params_t prm{};
    int scheme = OR_FVSC_GAUSSVOLPOINT;
    IVec bcU, bcT, bcP;
    Vec bvU, bvT, bvP;          // fixedValue: value ; fixedGradient / qhdFlux: gradient
    double deltaT = 0, CoNum = -1;
    Vec U, UB, T, TB, p, pB, rho, rhoB, mu, muB, alpha, alphaB, aQGD, aQGDB, tau, tauB, BdFrc, BdFrcB;
    Vec tauf, rhof, Uf, Tf, muf, alphaf, Hif, BdFrcf, gradUf, gradTf, gradPf, phiu, phiwo, taubyrhof, phi, Wf, phiUf,
        phiTf, phiTauTReg;
    int lastIters = 0;
    double lastRes0 = 0, lastRes = 0;
};

struct or_ctx {
    std::unique_ptr<QhdState> qhd;
    // ---- mesh (borrowed pointers are copied)
    int nCells = 0, nFaces = 0, nInternal = 0, nPoints = 0, nPatches = 0, nBnd = 0;
    Vec points, C, V, Cf, Sf, magSf, w, dC, ndC, nbrCC;
    IVec faceOff, faceVerts, owner, neighbour, patchStart, patchSize, patchKind;
    int gD[3] = {1, 1, 1};
    int nD = 3;
    int nThreads = 1;
    // ---- derived
    IVec bfacePatch;            // nBnd -> patch id
    IVec cfOff, cfFace;         // cell -> faces (ascending)
    Vec nf;                     // nFaces*3   fvscStencil.C:126-131
    // volPointInterpolation [OF-v2312]
    IVec pcOff, pcCell; Vec pcW;       // point -> cells, weights (non patch points)
    std::vector<char> isPatchPoint;
    IVec pbOff, pbFace; Vec pbW;       // patch point -> boundary faces (bface index), weights
    IVec conPts; Vec conR;             // vertices of constraint patches (wedge, symmetryPlane) and their constraint tensor (pointConstraints [OF-v2312])
    // QGDCoeffs
    Vec hQGDf, hQGD;
    // GaussVolPointBase3D
    IVec ftype;                 // nFaces: 3,4 or 0 (other)
    Vec bmvON;                  // nBnd
    Vec gcoef;                  // nFaces*18 : [dir][6]
    Vec gvol;                   // nFaces
    // leastSquares (extendedFaceStencil*.C): per internal face CSR of neighbour cells, wf2 and Gdf, degenerate flag
    IVec lsOff, lsCell; Vec lsWf2, lsGdf; std::vector<char> lsDeg; bool lsBuilt = false;
    // GaussVolPointBase2D
    int ie1 = -1, ie2 = -1, ie3 = -1;
    Vec c1, c2, c3, c4, mv42, mv13;    // nFaces (boundary entries for "ordinary" patches)
    IVec ip1, ip3;
    // ---- QGDFoam state
    or_qgd_params_t prm{};
    int scheme = OR_FVSC_GAUSSVOLPOINT;
    IVec bcU, bcT, bcP;
    Vec bvU, bvT, bvP;
    double deltaT = 0;
    // vol fields: internal (nCells*k) + boundary (nBnd*k)
    Vec rho, rhoB, U, UB, p, pB, e, eB, T, TB, psi, psiB, mu, muB, alpha, alphaB, gamma, gammaB, c, cB;
    Vec rhoU, rhoUB, rhoE, rhoEB, H, HB;
    Vec aQGD, aQGDB, tauQGD, tauQGDB, muQGD, muQGDB, alphauQGD, alphauQGDB, ScQGD, ScQGDB, PrQGD, PrQGDB, hQGDB;
    Vec pGrad;                  // qgdFlux gradient per bface
    IVec constScCells;          // varScModel7 / varScModel5 constScCellSet
    Vec cqSc;                   // varScModel5: mesh-quality floor of ScQGD (varScModel5.C:112-132)
    IVec pcgBlocks;             // processor of each cell for the linear solvers of a decomposed run (empty = serial)
    IVec lsForcedDeg;           // faceSet degenerateStencilFaces (leastSquaresStencil.C:63-132): internal faces forced to nf*snGrad
    Vec suRho, suU, suE;        // explicit source matrices rhoSu / rhoUSu / rhoESu: volume-integrated source per cell (empty = zero)
    // surface fields (nFaces*k)
    Vec tauQGDf, rhof, Uf, rhoUf, UrhoUf, pf, cf, gammaf, Hf, alphauf, muf;
    Vec gradUf, divUf, gradef, gradRhof, gradPf, rhoW, phiw, jm, phiJm, phi, phiJmU, phiP, Pif, phiPi,
        phiJmH, qf, phiQ, phiPiU, phiSigmaDotU, tauMC, phiTauMC;
    int lastDiffIters[4] = {0, 0, 0, 0};     // PCG iterations of the last Ux,Uy,Uz,e solves (implicit branch)
    bool qgdReady = false;
    bool havePhiwStar = false;
    bool haveU = false;         // a volVectorField "U" is registered (false while the thermo is constructed, createFields.H:3-24)
};

namespace {

// ----------------------------------------------------------------------------
// small helpers
static inline bool patchIsEmpty(const or_ctx& m, int bf) { return m.patchKind[m.bfacePatch[bf]] == OR_PATCH_EMPTY; }
static inline bool patchIsProc(const or_ctx& m, int bf) { return m.patchKind[m.bfacePatch[bf]] == OR_PATCH_PROCESSOR; }
static inline bool patchIsWedge(const or_ctx& m, int bf) { return m.patchKind[m.bfacePatch[bf]] == OR_PATCH_WEDGE; }

// [OF-v2312] linearInterpolate / surfaceInterpolation::weights  
//   internal: w(phiP - phiN) + phiN ; coupled patch: w phiP + (1-w) phiNbr ; other patch: boundary value
void linearInterpolate(const or_ctx& m, int k, const double* cell, const double* bnd, const double* nbr, double* out)
{
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
    for (int f = 0; f < m.nInternal; ++f) {
        const int P = m.owner[f], N = m.neighbour[f];
        const double w = m.w[f];
        for (int j = 0; j < k; ++j) out[(long)f * k + j] = w * (cell[(long)P * k + j] - cell[(long)N * k + j]) + cell[(long)N * k + j];
    }
    for (int b = 0; b < m.nBnd; ++b) {
        const int f = m.nInternal + b;
        for (int j = 0; j < k; ++j) {
            double v = 0.0;
            if (patchIsEmpty(m, b)) v = 0.0;
            else if (patchIsProc(m, b) && nbr) {
                const int P = m.owner[f];
                v = m.w[f] * cell[(long)P * k + j] + (1.0 - m.w[f]) * nbr[(long)b * k + j];
            } else v = bnd[(long)b * k + j];
            out[(long)f * k + j] = v;
        }
    }
}

// [OF-v2312] fvc::snGrad, "uncorrected" scheme on internal faces (nonOrthDeltaCoeffs), patch snGrad() on
// boundary faces (supplied by the caller, it depends on the BC type)  
void snGrad(const or_ctx& m, int k, const double* cell, const double* bndSnGrad, double* out)
{
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
    for (int f = 0; f < m.nInternal; ++f) {
        const int P = m.owner[f], N = m.neighbour[f];
        for (int j = 0; j < k; ++j) out[(long)f * k + j] = m.ndC[f] * (cell[(long)N * k + j] - cell[(long)P * k + j]);
    }
    for (int b = 0; b < m.nBnd; ++b)
        for (int j = 0; j < k; ++j) out[(long)(m.nInternal + b) * k + j] = patchIsEmpty(m, b) ? 0.0 : bndSnGrad[(long)b * k + j];
}

// [OF-v2312] volPointInterpolation::interpolate  (called at
// GaussVolPointBase3D.C:936-942, GaussVolPointBase2D.C:307-313)
void volPointInterpolate(const or_ctx& m, int k, const double* cell, const double* bnd, double* pf)
{
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
    for (int p = 0; p < m.nPoints; ++p) {
        for (int j = 0; j < k; ++j) pf[(long)p * k + j] = 0.0;
        if (!m.isPatchPoint[p]) {
            for (int q = m.pcOff[p]; q < m.pcOff[p + 1]; ++q)
                for (int j = 0; j < k; ++j) pf[(long)p * k + j] += m.pcW[q] * cell[(long)m.pcCell[q] * k + j];
        } else {
            for (int q = m.pbOff[p]; q < m.pbOff[p + 1]; ++q)
                for (int j = 0; j < k; ++j) pf[(long)p * k + j] += m.pbW[q] * bnd[(long)m.pbFace[q] * k + j];
        }
    }
    // [OF-v2312] pointConstraints::constrain: wedgePointPatchField / symmetryPlanePointPatchField::evaluate take the patch-normal
    // component out of a vector at the patch's vertices (transform(I - nHat nHat, v)), constrainCorners applies the combined
    // constraint where such patches meet; with the constraint tensor R of the vertex: v -> R.v, a tensor T -> R.T.R^T; scalars stay
    if (k == 3 || k == 9)
        for (size_t i = 0; i < m.conPts.size(); ++i) {
            const double* R = &m.conR[9 * i];
            double* v = &pf[(long)m.conPts[i] * k];
            if (k == 3) {
                const double u[3] = {v[0], v[1], v[2]};
                for (int a = 0; a < 3; ++a) v[a] = R[3 * a] * u[0] + R[3 * a + 1] * u[1] + R[3 * a + 2] * u[2];
            } else {
                double t[9];
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { t[3 * a + b] = 0.0; for (int c = 0; c < 3; ++c) t[3 * a + b] += R[3 * a + c] * v[3 * c + b]; }
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { v[3 * a + b] = 0.0; for (int c = 0; c < 3; ++c) v[3 * a + b] += t[3 * a + c] * R[3 * b + c]; }
            }
        }
}

void buildDerived(or_ctx& m)
{
    m.nBnd = m.nFaces - m.nInternal;
    m.nD = (m.gD[0] > 0) + (m.gD[1] > 0) + (m.gD[2] > 0);
    m.bfacePatch.assign(m.nBnd, -1);
    for (int pi = 0; pi < m.nPatches; ++pi)
        for (int i = 0; i < m.patchSize[pi]; ++i) m.bfacePatch[m.patchStart[pi] - m.nInternal + i] = pi;
    // cell -> faces, ascending face index (same visiting order as the sequential owner/neighbour loops of
    // fvc::surfaceIntegrate [OF-v2312] and of mesh.cells() for our purposes)
    {
        IVec cnt(m.nCells + 1, 0);
        for (int f = 0; f < m.nFaces; ++f) cnt[m.owner[f] + 1]++;
        for (int f = 0; f < m.nInternal; ++f) cnt[m.neighbour[f] + 1]++;
        for (int c = 0; c < m.nCells; ++c) cnt[c + 1] += cnt[c];
        m.cfOff = cnt;
        m.cfFace.assign(cnt[m.nCells], 0);
        IVec pos(cnt.begin(), cnt.end() - 1);
        for (int f = 0; f < m.nFaces; ++f) {
            if (f < m.nInternal) {
                // a face is visited for owner and neighbour at the same face index: keep ascending order per cell
            }
            m.cfFace[pos[m.owner[f]]++] = f;
            if (f < m.nInternal) m.cfFace[pos[m.neighbour[f]]++] = f;
        }
    }
    // nf = Sf/magSf   fvscStencil.C:126-131
    m.nf.assign((size_t)m.nFaces * 3, 0.0);
    for (int f = 0; f < m.nFaces; ++f)
        for (int d = 0; d < 3; ++d) m.nf[3 * (size_t)f + d] = m.Sf[3 * (size_t)f + d] / m.magSf[f];

    // ---- volPointInterpolation weights [OF-v2312 makeInternalWeights / makeBoundaryWeights]
    {
        std::vector<std::vector<int>> pc(m.nPoints);
        for (int f = 0; f < m.nFaces; ++f)
            for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q) {
                pc[m.faceVerts[q]].push_back(m.owner[f]);
                if (f < m.nInternal) pc[m.faceVerts[q]].push_back(m.neighbour[f]);
            }
        m.pcOff.assign(m.nPoints + 1, 0);
        for (int p = 0; p < m.nPoints; ++p) {
            auto& v = pc[p];
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
            m.pcOff[p + 1] = m.pcOff[p] + (int)v.size();
        }
        m.pcCell.resize(m.pcOff[m.nPoints]);
        m.pcW.resize(m.pcOff[m.nPoints]);
        m.isPatchPoint.assign(m.nPoints, 0);
        std::vector<std::vector<int>> pb(m.nPoints);
        for (int b = 0; b < m.nBnd; ++b) {
            const int kind = m.patchKind[m.bfacePatch[b]];
            const bool isPatchFace = (kind != OR_PATCH_EMPTY && kind != OR_PATCH_PROCESSOR);
            const int f = m.nInternal + b;
            for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q) {
                if (isPatchFace) { m.isPatchPoint[m.faceVerts[q]] = 1; pb[m.faceVerts[q]].push_back(b); }
            }
        }
        for (int p = 0; p < m.nPoints; ++p) {
            double sum = 0.0;
            const V3 x = ld3(m.points.data(), p);
            for (size_t i = 0; i < pc[p].size(); ++i) {
                const int q = m.pcOff[p] + (int)i;
                m.pcCell[q] = pc[p][i];
                m.pcW[q] = 1.0 / mag(x - ld3(m.C.data(), pc[p][i]));
                sum += m.pcW[q];
            }
            for (int q = m.pcOff[p]; q < m.pcOff[p + 1]; ++q) m.pcW[q] /= sum;
        }
        {   // vertices of the constraint patches (wedge, symmetryPlane) and their pointConstraint [OF-v2312 pointConstraintI.H]: each
            // patch the vertex lies on adds its (planar) normal - the patch's first point normal = its face normal - in patch order:
            //   none yet -> one plane (n) ; one -> a line along n x n_old if the two normals differ (|n x n_old| > 1e-3) ;
            //   a line -> fixed if n is not perpendicular to it (|n . d| > 1e-3)
            // constraintTransformation: I - n n | d d | 0
            std::vector<int> cnt(m.nPoints, 0);
            std::vector<V3> dir(m.nPoints, V3{0, 0, 0});
            for (int pi = 0; pi < m.nPatches; ++pi) {
                if ((m.patchKind[pi] != OR_PATCH_WEDGE && m.patchKind[pi] != OR_PATCH_SYMMETRY_PLANE) || m.patchSize[pi] == 0) continue;
                const int f0 = m.patchStart[pi];
                const V3 n = (1.0 / m.magSf[f0]) * ld3(m.Sf.data(), f0);
                std::vector<char> done(m.nPoints, 0);
                for (int f = f0; f < f0 + m.patchSize[pi]; ++f)
                    for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q) {
                        const int p = m.faceVerts[q];
                        if (done[p]) continue;
                        done[p] = 1;
                        if (cnt[p] == 0) { cnt[p] = 1; dir[p] = n; }
                        else if (cnt[p] == 1) { const V3 pl = cross(n, dir[p]); const double mp = mag(pl); if (mp > 1e-3) { cnt[p] = 2; dir[p] = (1.0 / mp) * pl; } }
                        else if (cnt[p] == 2) { if (std::fabs(dot(n, dir[p])) > 1e-3) { cnt[p] = 3; dir[p] = V3{0, 0, 0}; } }
                    }
            }
            for (int p = 0; p < m.nPoints; ++p) {
                if (!cnt[p]) continue;
                m.conPts.push_back(p);
                const double d[3] = {dir[p].x, dir[p].y, dir[p].z};
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        m.conR.push_back(cnt[p] == 1 ? (a == b ? 1.0 : 0.0) - d[a] * d[b] : (cnt[p] == 2 ? d[a] * d[b] : 0.0));
            }
        }
        m.pbOff.assign(m.nPoints + 1, 0);
        for (int p = 0; p < m.nPoints; ++p) m.pbOff[p + 1] = m.pbOff[p] + (int)pb[p].size();
        m.pbFace.resize(m.pbOff[m.nPoints]);
        m.pbW.resize(m.pbOff[m.nPoints]);
        for (int p = 0; p < m.nPoints; ++p) {
            double sum = 0.0;
            const V3 x = ld3(m.points.data(), p);
            for (size_t i = 0; i < pb[p].size(); ++i) {
                const int q = m.pbOff[p] + (int)i;
                m.pbFace[q] = pb[p][i];
                m.pbW[q] = 1.0 / mag(x - ld3(m.Cf.data(), m.nInternal + pb[p][i]));
                sum += m.pbW[q];
            }
            for (int q = m.pbOff[p]; q < m.pbOff[p + 1]; ++q) m.pbW[q] /= sum;
        }
    }

    // ---- QGDCoeffs ctor + updateQGDLength   QGDCoeffs.C:195-199, 298-376
    {
        m.hQGDf.assign(m.nFaces, 0.0);
        for (int f = 0; f < m.nFaces; ++f) m.hQGDf[f] = 1.0 / std::fabs(m.dC[f]);          // :198
        for (int f = 0; f < m.nInternal; ++f) {                                            // :303-308
            const double hown = mag(ld3(m.C.data(), m.owner[f]) - ld3(m.Cf.data(), f));
            const double hnei = mag(ld3(m.C.data(), m.neighbour[f]) - ld3(m.Cf.data(), f));
            m.hQGDf[f] = 2.0 * std::min(hown, hnei);
        }
        for (int b = 0; b < m.nBnd; ++b)                                                   // :310-317
            if (!patchIsProc(m, b)) m.hQGDf[m.nInternal + b] *= 2.0;
        m.hQGD.assign(m.nCells, 0.0);
        for (int c = 0; c < m.nCells; ++c) {                                               // :323-362
            double hint = 0.0, surf = 0.0;
            for (int q = m.cfOff[c]; q < m.cfOff[c + 1]; ++q) {
                const int fid = m.cfFace[q];
                if (fid < m.nInternal) { hint += m.hQGDf[fid] * m.magSf[fid]; surf += m.magSf[fid]; }
                else {
                    const int b = fid - m.nInternal;
                    if (!patchIsEmpty(m, b) && !patchIsWedge(m, b)) { hint += m.hQGDf[fid] * m.magSf[fid]; surf += m.magSf[fid]; }
                }
            }
            m.hQGD[c] = hint / surf;
        }
    }

    // ---- GaussVolPointBase3D ctor   GaussVolPointBase3D.C:74-158
    {
        m.ftype.assign(m.nFaces, 0);
        for (int f = 0; f < m.nFaces; ++f) {
            const int n = m.faceOff[f + 1] - m.faceOff[f];
            m.ftype[f] = (n == 3) ? 3 : ((n == 4) ? 4 : 0);
        }
        m.bmvON.assign(m.nBnd, 0.0);
        m.gcoef.assign((size_t)m.nFaces * 18, 0.0);
        m.gvol.assign(m.nFaces, 1.0);
        const double OneBySix = (1.0 / 6.0);
        for (int f = 0; f < m.nFaces; ++f) {
            V3 vO, vN;   // owner-side and neighbour-side "cell centres"
            if (f < m.nInternal) { vO = ld3(m.C.data(), m.owner[f]); vN = ld3(m.C.data(), m.neighbour[f]); }
            else {
                const int b = f - m.nInternal;
                vO = ld3(m.C.data(), m.owner[f]);                                          // :134  Cn()
                if (patchIsProc(m, b)) vN = ld3(m.nbrCC.data(), b);                        // :137-138
                else vN = vO + 2.0 * (ld3(m.Cf.data(), f) - vO);                           // :142-147
                m.bmvON[b] = mag(vO - vN);                                                 // :150-153
            }
            const int* fv = &m.faceVerts[m.faceOff[f]];
            double* a = &m.gcoef[(size_t)f * 18];
            if (m.ftype[f] == 3) {                                                         // triCalcWeights :161-318
                const V3 p1 = ld3(m.points.data(), fv[0]), p2 = ld3(m.points.data(), fv[1]), p3 = ld3(m.points.data(), fv[2]);
                m.gvol[f] = dot(cross(p2 - p1, p3 - p1), vO - vN) * OneBySix;              // :186-190
                // X  :193-203
                a[0] = OneBySix * ((vO.z - vN.z) * (p2.y - p3.y) + (vN.y - vO.y) * (p2.z - p3.z));
                a[1] = OneBySix * ((vN.y - vO.y) * (p3.z - p1.z) + (vO.z - vN.z) * (p3.y - p1.y));
                a[2] = OneBySix * ((vN.y - vO.y) * (p1.z - p2.z) + (vO.z - vN.z) * (p1.y - p2.y));
                a[3] = OneBySix * (p1.z * (p2.y - p3.y) + p2.z * (p3.y - p1.y) + p3.z * (p1.y - p2.y));
                a[4] = -a[3];
                // Y  :206-216
                a[6 + 0] = OneBySix * ((vO.x - vN.x) * (p2.z - p3.z) + (vN.z - vO.z) * (p2.x - p3.x));
                a[6 + 1] = OneBySix * ((vN.z - vO.z) * (p3.x - p1.x) + (vO.x - vN.x) * (p3.z - p1.z));
                a[6 + 2] = OneBySix * ((vN.z - vO.z) * (p1.x - p2.x) + (vO.x - vN.x) * (p1.z - p2.z));
                a[6 + 3] = OneBySix * (p1.x * (p2.z - p3.z) + p2.x * (p3.z - p1.z) + p3.x * (p1.z - p2.z));
                a[6 + 4] = -a[6 + 3];
                // Z  :219-229
                a[12 + 0] = OneBySix * ((vO.y - vN.y) * (p2.x - p3.x) + (vN.x - vO.x) * (p2.y - p3.y));
                a[12 + 1] = OneBySix * ((vN.x - vO.x) * (p3.y - p1.y) + (vO.y - vN.y) * (p3.x - p1.x));
                a[12 + 2] = OneBySix * ((vN.x - vO.x) * (p1.y - p2.y) + (vO.y - vN.y) * (p1.x - p2.x));
                a[12 + 3] = OneBySix * (p1.y * (p2.x - p3.x) + p2.y * (p3.x - p1.x) + p3.y * (p1.x - p2.x));
                a[12 + 4] = -a[12 + 3];
            } else if (m.ftype[f] == 4) {                                                  // quaCalcWeights :320-476
                const V3 p1 = ld3(m.points.data(), fv[0]), p2 = ld3(m.points.data(), fv[1]),
                         p3 = ld3(m.points.data(), fv[2]), p4 = ld3(m.points.data(), fv[3]);
                m.gvol[f] = dot(p3 - p1, cross(p4 - p2, vO - vN)) * OneBySix;              // :346-350
                // X  :353-363
                a[0] = OneBySix * ((vN.y - vO.y) * (p2.z - p4.z) - (vN.z - vO.z) * (p2.y - p4.y));
                a[1] = OneBySix * ((vN.y - vO.y) * (p3.z - p1.z) - (vN.z - vO.z) * (p3.y - p1.y));
                a[5] = OneBySix * ((p1.y - p3.y) * (p2.z - p4.z) - (p1.z - p3.z) * (p2.y - p4.y));
                a[2] = -a[0]; a[3] = -a[1]; a[4] = -a[5];
                // Y  :366-376
                a[6 + 0] = OneBySix * ((vN.z - vO.z) * (p2.x - p4.x) - (vN.x - vO.x) * (p2.z - p4.z));
                a[6 + 1] = OneBySix * ((vN.z - vO.z) * (p3.x - p1.x) - (vN.x - vO.x) * (p3.z - p1.z));
                a[6 + 5] = OneBySix * ((p1.z - p3.z) * (p2.x - p4.x) - (p1.x - p3.x) * (p2.z - p4.z));
                a[6 + 2] = -a[6 + 0]; a[6 + 3] = -a[6 + 1]; a[6 + 4] = -a[6 + 5];
                // Z  :379-389
                a[12 + 0] = OneBySix * ((vN.x - vO.x) * (p2.y - p4.y) - (vN.y - vO.y) * (p2.x - p4.x));
                a[12 + 1] = OneBySix * ((vN.x - vO.x) * (p3.y - p1.y) - (vN.y - vO.y) * (p3.x - p1.x));
                a[12 + 5] = OneBySix * ((p1.x - p3.x) * (p2.y - p4.y) - (p1.y - p3.y) * (p2.x - p4.x));
                a[12 + 2] = -a[12 + 0]; a[12 + 3] = -a[12 + 1]; a[12 + 4] = -a[12 + 5];
            }
        }
    }

    // ---- GaussVolPointBase2D ctor   GaussVolPointBase2D.C:72-293
    if (m.nD == 2) {
        for (int d = 0; d < 3; ++d) if (m.gD[d] < 1) m.ie3 = d;                            // :90-96
        V3 e1{1, 0, 0}, e2{0, 1, 0};
        if (m.ie3 == 0) { e1 = {0, 1, 0}; e2 = {0, 0, 1}; m.ie1 = 1; m.ie2 = 2; }          // :97-120
        if (m.ie3 == 1) { e1 = {1, 0, 0}; e2 = {0, 0, 1}; m.ie1 = 0; m.ie2 = 2; }
        if (m.ie3 == 2) { e1 = {1, 0, 0}; e2 = {0, 1, 0}; m.ie1 = 0; m.ie2 = 1; }
        m.c1.assign(m.nFaces, 0.0); m.c2 = m.c1; m.c3 = m.c1; m.c4 = m.c1;
        m.mv42.assign(m.nFaces, 1.0); m.mv13 = m.mv42;
        m.ip1.assign(m.nFaces, -1); m.ip3 = m.ip1;
        int ip1 = -1, ip3 = -1;
        for (int f = 0; f < m.nFaces; ++f) {
            V3 cRef, v42;
            if (f < m.nInternal) {
                const int ic2 = m.neighbour[f], ic4 = m.owner[f];
                cRef = ld3(m.C.data(), ic2);                                               // :131  compares with C[ic2]
                v42 = ld3(m.C.data(), ic2) - ld3(m.C.data(), ic4);                         // :154
            } else {
                const int b = f - m.nInternal;
                const int kind = m.patchKind[m.bfacePatch[b]];
                if (kind == OR_PATCH_EMPTY || kind == OR_PATCH_WEDGE) continue;            // :175-179
                const int ic4 = m.owner[f];
                cRef = ld3(m.C.data(), ic4);                                               // :250
                if (kind == OR_PATCH_PROCESSOR) v42 = ld3(m.nbrCC.data(), b) - cRef;       // :229-230
                else v42 = 2.0 * (ld3(m.Cf.data(), f) - cRef);                             // :237-238
                ip1 = -1; ip3 = -1;                                                        // :244
            }
            for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q)                          // :129-136
                if (cmp(ld3(m.points.data(), m.faceVerts[q]), m.ie3) >= cmp(cRef, m.ie3)) { ip1 = m.faceVerts[q]; break; }
            for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q)                          // :137-147
                if (cmp(ld3(m.points.data(), m.faceVerts[q]), m.ie3) >= cmp(cRef, m.ie3))
                    if (ip1 != m.faceVerts[q]) { ip3 = m.faceVerts[q]; break; }
            m.ip1[f] = ip1; m.ip3[f] = ip3;
            const V3 v13 = ld3(m.points.data(), ip3) - ld3(m.points.data(), ip1);          // :155
            m.mv42[f] = mag(v42); m.mv13[f] = mag(v13);
            const double cosa1 = dot((1.0 / m.mv42[f]) * v42, e1), cosa2 = dot((1.0 / m.mv13[f]) * v13, e1);
            const double sina1 = dot((1.0 / m.mv42[f]) * v42, e2), sina2 = dot((1.0 / m.mv13[f]) * v13, e2);
            const double den = sina2 * cosa1 - sina1 * cosa2;                              // :164
            m.c1[f] = sina2 / den; m.c2[f] = sina1 / den; m.c3[f] = cosa1 / den; m.c4[f] = cosa2 / den;
        }
    }
}

// ghost / neighbour value used on boundary faces  GaussVolPointBase3D.C:780-794, GaussVolPointBase2D.C:333-347
static inline double psiN(const or_ctx& m, int b, int k, int j, const double* bnd, const double* bsg, const double* nbr, double halfDist)
{
    if (patchIsProc(m, b)) return nbr[(long)b * k + j];
    return bnd[(long)b * k + j] + bsg[(long)b * k + j] * halfDist;
}

// One application of the dfdxif / dfdxbf macros (GaussVolPointBase3D.C:488-539) for face f:
// returns sum_k a[k] phi_k / v  for input component icmpt, direction dir.
static inline double dfdx(const or_ctx& m, int f, int dir, int k, int icmpt, const double* cell, const double* pF,
                          const double* bnd, const double* bsg, const double* nbr)
{
    const int nv = m.ftype[f];
    const int iown = nv + 1, inei = nv;                        // :492-493  (row size nv+2)
    const double* a = &m.gcoef[(size_t)f * 18 + 6 * dir];
    double phiN, phiP;
    if (f < m.nInternal) { phiN = cell[(long)m.neighbour[f] * k + icmpt]; phiP = cell[(long)m.owner[f] * k + icmpt]; }
    else {
        const int b = f - m.nInternal;
        phiN = psiN(m, b, k, icmpt, bnd, bsg, nbr, m.bmvON[b] * 0.5);
        phiP = cell[(long)m.owner[f] * k + icmpt];             // patchInternalField :796
    }
    double s = phiN * a[inei];
    s += phiP * a[iown];
    for (int q = 0; q < nv; ++q) s += pF[(long)m.faceVerts[m.faceOff[f] + q] * k + icmpt] * a[q];
    return s / m.gvol[f];
}

// GaussVolPoint::Grad / reduced::Grad for scalar (k=1) and vector (k=3) fields.
// out: nFaces*3k, tensor index 3*i+j = d_i phi_j.   GaussVolPointStencil.C:71-99, GaussVolPointBase.C:54-121,
// GaussVolPointBase1D.C:49-63, GaussVolPointBase2D.C:301-367, GaussVolPointBase3D.C:740-993,
// reducedFaceNormalStencil.C:69-88.  The caller has already applied correctBoundaryConditions().
// leastSquares scheme: leastSquaresBase::findNeighbours (extendedFaceStencilFindNeighbours.C:41-86, serial part) and
// calculateWeights (extendedFaceStencilCalculateWeights.C:43-155)
void buildLeastSquares(or_ctx& m)
{
    if (m.lsBuilt) return;
    // [OF-v2312] primitiveMesh::pointCells(): cells of each point in ascending order
    std::vector<std::vector<int>> pc(m.nPoints);
    for (int f = 0; f < m.nFaces; ++f)
        for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q) {
            pc[m.faceVerts[q]].push_back(m.owner[f]);
            if (f < m.nInternal) pc[m.faceVerts[q]].push_back(m.neighbour[f]);
        }
    for (auto& v : pc) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    m.lsOff.assign(m.nInternal + 1, 0);
    m.lsCell.clear(); m.lsWf2.clear(); m.lsGdf.clear();
    m.lsDeg.assign(std::max(m.nInternal, 1), 0);
    for (int f = 0; f < m.nInternal; ++f) {
        std::vector<int> nb;                                                               // :52-82, first-seen order
        for (int q = m.faceOff[f]; q < m.faceOff[f + 1]; ++q)
            for (int c : pc[m.faceVerts[q]]) if (std::find(nb.begin(), nb.end(), c) == nb.end()) nb.push_back(c);
        const V3 Cf = ld3(m.Cf.data(), f);
        std::vector<V3> df(nb.size());
        std::vector<double> wf2(nb.size());
        double G[6] = {0, 0, 0, 0, 0, 0};                                                  // symmTensor xx,xy,xz,yy,yz,zz
        for (size_t i = 0; i < nb.size(); ++i) {                                           // :70-78
            df[i] = ld3(m.C.data(), nb[i]) - Cf;
            wf2[i] = 1.0 / dot(df[i], df[i]);
            const V3 d = df[i];
            const double a[6] = {d.x * d.x, d.x * d.y, d.x * d.z, d.y * d.y, d.y * d.z, d.z * d.z};
            for (int t = 0; t < 6; ++t) G[t] += a[t] * wf2[i];
        }
        double G0[6] = {0, 0, 0, 0, 0, 0};
        if (std::fabs(G[0]) < SMALL) G0[0] = 1;                                            // :85-99
        if (std::fabs(G[3]) < SMALL) G0[3] = 1;
        if (std::fabs(G[5]) < SMALL) G0[5] = 1;
        for (int t = 0; t < 6; ++t) G[t] += G0[t];                                         // :129
        const double xx = G[0], xy = G[1], xz = G[2], yy = G[3], yz = G[4], zz = G[5];
        const double detG = xx * yy * zz + xy * yz * xz + xz * xy * yz - xx * yz * yz - xy * xy * zz - xz * yy * xz;   // [OF det(symmTensor)]
        double Gi[6] = {G[0], G[1], G[2], G[3], G[4], G[5]};
        if (detG < 1) m.lsDeg[f] = 1;                                                      // :136-140
        else {                                                                             // :143-144  inv(G) - G0  [OF inv(symmTensor)]
            Gi[0] = (yy * zz - yz * yz) / detG; Gi[1] = (xz * yz - xy * zz) / detG; Gi[2] = (xy * yz - xz * yy) / detG;
            Gi[3] = (xx * zz - xz * xz) / detG; Gi[4] = (xy * xz - xx * yz) / detG; Gi[5] = (xx * yy - xy * xy) / detG;
            for (int t = 0; t < 6; ++t) Gi[t] -= G0[t];
        }
        for (size_t i = 0; i < nb.size(); ++i) {                                           // :147-150  G & df
            const V3 d = df[i];
            m.lsCell.push_back(nb[i]);
            m.lsWf2.push_back(wf2[i]);
            m.lsGdf.push_back(Gi[0] * d.x + Gi[1] * d.y + Gi[2] * d.z);
            m.lsGdf.push_back(Gi[1] * d.x + Gi[3] * d.y + Gi[4] * d.z);
            m.lsGdf.push_back(Gi[2] * d.x + Gi[4] * d.y + Gi[5] * d.z);
        }
        m.lsOff[f + 1] = (int)m.lsCell.size();
    }
    for (int f : m.lsForcedDeg) if (f >= 0 && f < m.nInternal) m.lsDeg[f] = 1;            // leastSquaresStencil.C:84-91,117
    m.lsBuilt = true;
}

// leastSquares::Grad(volScalarField) applied per component (extendedFaceStencilScalarGrad.C:50-109 ;
// leastSquaresStencil.C:145-196 for vectors): out[f][k*i + j] = d_i phi_j
// opt: leastSquaresOpt (leastSquaresStencilOpt.C:75-181, extendedFaceStencilScalarDer.C:40-52): the same sums, but the
// degenerate faces are not replaced by nf*snGrad - their Gdf is the un-inverted (G+G0)&df
void leastSquaresGrad(const or_ctx& mc, int k, const double* cell, const double* bnd, const double* bsg, double* out, bool opt = false)
{
    or_ctx& m = const_cast<or_ctx&>(mc);
    buildLeastSquares(m);
    const int ok = 3 * k;
    Vec sF((size_t)m.nFaces * k), sn((size_t)m.nFaces * k);
    linearInterpolate(m, k, cell, bnd, nullptr, sF.data());                                // :52
    snGrad(m, k, cell, bsg, sn.data());                                                    // :53
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
    for (int f = 0; f < m.nInternal; ++f)
        for (int j = 0; j < k; ++j) {
            double g[3] = {0, 0, 0};
            if (m.lsDeg[f] && !opt) for (int i = 0; i < 3; ++i) g[i] = sn[(size_t)f * k + j] * m.nf[3 * (size_t)f + i];        // :78-83
            else
                for (int q = m.lsOff[f]; q < m.lsOff[f + 1]; ++q) {                        // :67-70
                    const double d = cell[(size_t)m.lsCell[q] * k + j] - sF[(size_t)f * k + j];
                    for (int i = 0; i < 3; ++i) g[i] = g[i] + m.lsWf2[q] * m.lsGdf[3 * (size_t)q + i] * d;
                }
            for (int i = 0; i < 3; ++i) out[(size_t)f * ok + k * i + j] = g[i];
        }
    for (int b = 0; b < m.nBnd; ++b) {                                                     // :86-109
        const int f = m.nInternal + b;
        const int kind = m.patchKind[m.bfacePatch[b]];
        const bool constrained = kind == OR_PATCH_EMPTY || kind == OR_PATCH_WEDGE || kind == OR_PATCH_PROCESSOR || kind == OR_PATCH_SYMMETRY_PLANE;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < k; ++j) out[(size_t)f * ok + k * i + j] = constrained ? 0.0 : m.nf[3 * (size_t)f + i] * bsg[(size_t)b * k + j];
    }
}

void fvscGrad(const or_ctx& m, int scheme, int k, const double* cell, const double* bnd, const double* bsg,
              const double* nbr, double* out)
{
    if (scheme == OR_FVSC_LEASTSQUARES || scheme == OR_FVSC_LEASTSQUARESOPT) {
        leastSquaresGrad(m, k, cell, bnd, bsg, out, scheme == OR_FVSC_LEASTSQUARESOPT);
        return;
    }
    const int ok = 3 * k;
    std::fill(out, out + (size_t)m.nFaces * ok, 0.0);          // vector::zero * fvc::snGrad(vF)
    Vec sn((size_t)m.nFaces * k);
    snGrad(m, k, cell, bsg, sn.data());
    if (scheme == OR_FVSC_REDUCED || m.nD == 1) {              // nf * snGrad
        for (int f = 0; f < m.nFaces; ++f)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < k; ++j) out[(size_t)f * ok + k * i + j] = m.nf[3 * (size_t)f + i] * sn[(size_t)f * k + j];
        return;
    }
    Vec pF((size_t)m.nPoints * k);
    volPointInterpolate(m, k, cell, bnd, pF.data());
    if (m.nD == 2) {
        // scalar 2D version, applied per component for vectors (GaussVolPointBase.C:79-116)
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
        for (int f = 0; f < m.nFaces; ++f) {
            if (f >= m.nInternal) {
                const int b = f - m.nInternal;
                const int kind = m.patchKind[m.bfacePatch[b]];
                if (kind == OR_PATCH_EMPTY || kind == OR_PATCH_WEDGE) continue;
            }
            for (int j = 0; j < k; ++j) {
                double phiN, phiP;
                if (f < m.nInternal) { phiN = cell[(long)m.neighbour[f] * k + j]; phiP = cell[(long)m.owner[f] * k + j]; }
                else {
                    const int b = f - m.nInternal;
                    phiN = psiN(m, b, k, j, bnd, bsg, nbr, m.mv42[f] * 0.5);
                    phiP = cell[(long)m.owner[f] * k + j];
                }
                const double dfdn = (phiN - phiP) / m.mv42[f];                                       // :317
                const double dfdt = (pF[(long)m.ip3[f] * k + j] - pF[(long)m.ip1[f] * k + j]) / m.mv13[f]; // :321
                out[(size_t)f * ok + k * m.ie1 + j] = (dfdn * m.c1[f] - dfdt * m.c2[f]);             // :324
                out[(size_t)f * ok + k * m.ie2 + j] = (dfdt * m.c3[f] - dfdn * m.c4[f]);             // :325
                out[(size_t)f * ok + k * m.ie3 + j] = 0.0;
            }
        }
        return;
    }
    // 3D
#pragma omp parallel for num_threads(m.nThreads) schedule(static)
    for (int f = 0; f < m.nFaces; ++f) {
        if (f >= m.nInternal && patchIsEmpty(m, f - m.nInternal)) continue;
        if (m.ftype[f] == 0) {                                  // other faces: nf*snGrad  :760-768
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < k; ++j) out[(size_t)f * ok + k * i + j] = m.nf[3 * (size_t)f + i] * sn[(size_t)f * k + j];
            continue;
        }
        if (k == 1) {
            for (int dir = 0; dir < 3; ++dir) out[(size_t)f * 3 + dir] += dfdx(m, f, dir, 1, 0, cell, pF.data(), bnd, bsg, nbr);
        } else if (m.ftype[f] == 3 && f < m.nInternal) {
            // internal triangular faces: the reference's index pattern  :844-854
            for (int row = 0; row < 3; ++row)
                for (int d = 0; d < 3; ++d) out[(size_t)f * 9 + 3 * row + d] += dfdx(m, f, d, 3, d, cell, pF.data(), bnd, bsg, nbr);
        } else {
            for (int dir = 0; dir < 3; ++dir)                   // :831-841 / :896-919
                for (int j = 0; j < 3; ++j) out[(size_t)f * 9 + 3 * dir + j] += dfdx(m, f, dir, 3, j, cell, pF.data(), bnd, bsg, nbr);
        }
    }
}

// GaussVolPoint::Div / reduced::Div for vector (k=3 -> scalar) and tensor (k=9 -> vector) fields.
// GaussVolPointStencil.C:101-129, GaussVolPointBase1D.C:65-79, GaussVolPointBase2D.C:369-539,
// GaussVolPointBase3D.C:543-738, 996-1046
void fvscDiv(const or_ctx& m, int scheme, int k, const double* cell, const double* bnd, const double* bsg,
             const double* nbr, double* out)
{
    if (scheme == OR_FVSC_LEASTSQUARES || scheme == OR_FVSC_LEASTSQUARESOPT) {
        // leastSquaresStencil.C:204-275, leastSquaresStencilOpt.C:189-260: Div(vector) = sum_i d_i U_i ; Div(tensor)_j = sum_i d_i T_ij, from the component gradients
        const int okd = k / 3;
        Vec g((size_t)m.nFaces * 3 * k);
        leastSquaresGrad(m, k, cell, bnd, bsg, g.data(), scheme == OR_FVSC_LEASTSQUARESOPT);
        for (int f = 0; f < m.nFaces; ++f)
            for (int jj = 0; jj < okd; ++jj) {
                double sacc = 0.0;
                for (int i = 0; i < 3; ++i) sacc += g[(size_t)f * 3 * k + k * i + (okd * i + jj)];
                out[(size_t)f * okd + jj] = sacc;
            }
        return;
    }
    const int ok = k / 3;
    std::fill(out, out + (size_t)m.nFaces * ok, 0.0);
    Vec sn((size_t)m.nFaces * k);
    snGrad(m, k, cell, bsg, sn.data());
    auto nfDotSn = [&](int f) {
        for (int j = 0; j < ok; ++j) {
            double s = 0.0;
            for (int i = 0; i < 3; ++i) s += m.nf[3 * (size_t)f + i] * sn[(size_t)f * k + ok * i + j];   // nf & snGrad
            out[(size_t)f * ok + j] = s;
        }
    };
    if (scheme == OR_FVSC_REDUCED || m.nD == 1) { for (int f = 0; f < m.nFaces; ++f) nfDotSn(f); return; }
    Vec pF((size_t)m.nPoints * k);
    volPointInterpolate(m, k, cell, bnd, pF.data());
    if (m.nD == 2) {
        for (int f = 0; f < m.nFaces; ++f) {
            if (f >= m.nInternal) {
                const int b = f - m.nInternal;
                const int kind = m.patchKind[m.bfacePatch[b]];
                if (kind == OR_PATCH_EMPTY || kind == OR_PATCH_WEDGE) continue;
            }
            auto dn = [&](int j) {
                double phiN, phiP = cell[(long)m.owner[f] * k + j];
                if (f < m.nInternal) phiN = cell[(long)m.neighbour[f] * k + j];
                else phiN = psiN(m, f - m.nInternal, k, j, bnd, bsg, nbr, m.mv42[f] * 0.5);
                return (phiN - phiP) / m.mv42[f];
            };
            auto dt = [&](int j) { return (pF[(long)m.ip3[f] * k + j] - pF[(long)m.ip1[f] * k + j]) / m.mv13[f]; };
            if (k == 3) {                                      // :384-397
                out[f] = (dn(m.ie1) * m.c1[f] - dt(m.ie1) * m.c2[f]) + (dt(m.ie2) * m.c3[f] - dn(m.ie2) * m.c4[f]);
            } else {                                           // :447-485
                const int i11 = m.ie1 * 3 + m.ie1, i21 = m.ie2 * 3 + m.ie1, i22 = m.ie2 * 3 + m.ie2, i12 = m.ie1 * 3 + m.ie2;
                out[(size_t)f * 3 + m.ie1] = (dn(i11) * m.c1[f] - dt(i11) * m.c2[f]) + (dt(i21) * m.c3[f] - dn(i21) * m.c4[f]);
                out[(size_t)f * 3 + m.ie2] = (dn(i12) * m.c1[f] - dt(i12) * m.c2[f]) + (dt(i22) * m.c3[f] - dn(i22) * m.c4[f]);
            }
        }
        return;
    }
    for (int f = 0; f < m.nFaces; ++f) {
        if (f >= m.nInternal && patchIsEmpty(m, f - m.nInternal)) continue;
        if (m.ftype[f] == 0) { nfDotSn(f); continue; }        // :562-571
        if (k == 3) {
            for (int dir = 0; dir < 3; ++dir) out[f] += dfdx(m, f, dir, 3, dir, cell, pF.data(), bnd, bsg, nbr);   // :553-560
        } else {
            for (int j = 0; j < 3; ++j)                        // :635-659
                for (int dir = 0; dir < 3; ++dir) out[(size_t)f * 3 + j] += dfdx(m, f, dir, 9, 3 * dir + j, cell, pF.data(), bnd, bsg, nbr);
        }
    }
}

// ----------------------------------------------------------------------------
// thermo: perfectGas + hConst + sensibleInternalEnergy + constTransport [OF-v2312]  
struct Thermo {
    double R, Cp, Hf, Tref, Hsref, mu, Pr;
    int transport = 0; double mu0 = 0, T0 = 1, kExp = 0;                       // powerLawTransport.C:53-60
    double As = 0, Ts = 0;                                                     // sutherlandTransport [OF-v2312]
    int eConst = 0; double CvE = 0, Esref = 0;                                 // eConstThermo [OF-v2312]: Cv, Esref; Cp = Cv + CpMCv
    double Cv() const { return eConst ? CvE : Cp - R; }                        // hConst: Cv = Cp - CpMCv, CpMCv = R
    double CpT() const { return eConst ? CvE + R : Cp; }
    double Es(double /*p*/, double T) const {                                  // hConst: Hs - p/rho ; eConst: eConstThermoI.H Es
        return eConst ? CvE * (T - Tref) + Esref : Cp * (T - Tref) + Hsref - R * T;
    }
    double HE(double p, double T) const { return Es(p, T); }
    double THE(double e, double p, double T0) const {                          // thermo::T Newton loop, tol 1e-4
        double Test = T0, Tnew = T0;
        const double Ttol = T0 * 1e-4;
        int iter = 0;
        do {
            Test = Tnew;
            Tnew = Test - (Es(p, Test) - e) / Cv();
            if (iter++ > 100) break;
        } while (std::fabs(Tnew - Test) > Ttol);
        return Tnew;
    }
    double psi(double /*p*/, double T) const { return 1.0 / (R * T); }
    double muF(double, double T) const {
        if (transport == 1) return mu0 * std::pow(T / T0, kExp);               // powerLawTransportI.H:121-128
        if (transport == 2) return As * std::sqrt(T) / (1.0 + Ts / T);         // sutherlandTransportI.H mu()
        return mu;
    }
    double alphah(double p, double T) const {
        if (transport == 1) return muF(p, T) * (1.0 / Pr);                     // powerLawTransportI.H:143-150 (rPr_)
        if (transport == 2) { const double cv = Cv(); return muF(p, T) * cv * (1.32 + 1.77 * R / cv) / CpT(); }   // kappa/Cp, modified Eucken
        return mu / Pr;                                                        // constTransport
    }
    double gamma() const { return CpT() / Cv(); }
};

Thermo thermoOf(const or_ctx& s)
{
    Thermo t{s.prm.R, s.prm.Cp, s.prm.Hf, s.prm.Tref, s.prm.Hsref, s.prm.mu, s.prm.Pr};
    t.transport = s.prm.transportModel; t.mu0 = s.prm.mu0; t.T0 = s.prm.T0; t.kExp = s.prm.kExp;
    t.As = s.prm.As; t.Ts = s.prm.Ts;
    t.eConst = s.prm.thermoModel == 1; t.CvE = s.prm.Cv; t.Esref = s.prm.Esref;
    return t;
}

// [OF-v2312] wedgePolyPatch::calcGeometry + rotationTensor (restated from the OpenFOAM source as remembered): the patch normal n,
// the centre-plane normal = the coordinate axis n is closest to (components sign(n_i) (max(|n_i|, 0.5) - 0.5), normalised),
// faceT = rotationTensor(centreNormal, n) = s I + (1 - s) n3 n3 / |n3|^2 + (n n1 - n1 n) with n1 = centreNormal, s = n1 . n,
// n3 = n1 x n; cellT = faceT . faceT.  OpenFOAM averages n over the (planar) patch; here it is the face's own normal.
static void wedgeFaceT(const double* n, double (&T)[9])
{
    auto sgn = [](double v) { return v >= 0 ? 1.0 : -1.0; };
    double n1[3];
    for (int i = 0; i < 3; ++i) n1[i] = sgn(n[i]) * (std::max(std::fabs(n[i]), 0.5) - 0.5);
    const double m1 = std::sqrt(n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2]);
    for (int i = 0; i < 3; ++i) n1[i] /= m1;
    const double s = n1[0] * n[0] + n1[1] * n[1] + n1[2] * n[2];
    const double n3[3] = {n1[1] * n[2] - n1[2] * n[1], n1[2] * n[0] - n1[0] * n[2], n1[0] * n[1] - n1[1] * n[0]};
    const double m3 = n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = (i == j) ? s : 0.0;
            if (m3 > 1.0e-15) v += (1.0 - s) * n3[i] * n3[j] / m3 + (n[i] * n1[j] - n1[i] * n[j]);      // SMALL
            else v = (i == j) ? 1.0 : 0.0;                                                             // n == centreNormal
            T[3 * i + j] = v;
        }
}
static inline void matVec(const double (&T)[9], const double* u, double (&o)[3])
{
    for (int i = 0; i < 3; ++i) o[i] = T[3 * i] * u[0] + T[3 * i + 1] * u[1] + T[3 * i + 2] * u[2];
}

// patch snGrad() of a field with the given BC [OF-v2312 fvPatchField::snGrad, zeroGradient, fixedGradient]
void patchSnGrad(const or_ctx& s, int k, const IVec& bc, const double* cell, const double* bnd, const double* grad, double* out)
{
    for (int b = 0; b < s.nBnd; ++b) {
        const int pi = s.bfacePatch[b];
        const int f = s.nInternal + b;
        for (int j = 0; j < k; ++j) {
            double v;
            if (s.patchKind[pi] == OR_PATCH_EMPTY) v = 0.0;
            else if (bc[pi] == OR_BC_ZERO_GRADIENT) v = 0.0;
            else if (bc[pi] == OR_BC_FIXED_GRADIENT || bc[pi] == OR_BC_QGD_FLUX) v = grad ? grad[(long)b * k + j] : 0.0;
            else if (bc[pi] == OR_BC_WEDGE && k == 3) {      // [OF-v2312 wedgeFvPatchField::snGrad] (cellT . U_P - U_P) deltaCoeffs / 2
                double T[9], t1[3], t2[3];
                wedgeFaceT(&s.nf[3 * (size_t)f], T);
                matVec(T, &cell[(long)s.owner[f] * 3], t1);
                matVec(T, t1, t2);
                v = (t2[j] - cell[(long)s.owner[f] * 3 + j]) * (0.5 * s.dC[f]);
            }
            else v = s.dC[f] * (bnd[(long)b * k + j] - cell[(long)s.owner[f] * k + j]);
            out[(long)b * k + j] = v;
        }
    }
}

// correctBoundaryConditions() for U (fixedValue | zeroGradient | slip | wedge)
// wedge [OF-v2312 wedgeFvPatchField::evaluate]: U_b = transform(faceT, U_P)
// slip / symmetryPlane [OF-v2312 basicSymmetryFvPatchField::evaluate]: (U_P + transform(I - 2 nn, U_P))/2 = U_P - n (n . U_P);
// its snGrad (transform(I - 2nn, U_P) - U_P) deltaCoeffs/2 equals deltaCoeffs (U_b - U_P), the generic branch of patchSnGrad.
// Oracle only so far (explicit branch): the device library has no slip condition yet.
void correctU(or_ctx& s)
{
    for (int b = 0; b < s.nBnd; ++b) {
        const int pi = s.bfacePatch[b];
        if (s.patchKind[pi] == OR_PATCH_EMPTY) continue;
        const int f = s.nInternal + b;
        const int P = s.owner[f];
        if (s.bcU[pi] == OR_BC_SLIP) {
            const double* n = &s.nf[3 * (size_t)f];
            const double* u = &s.U[3 * (size_t)P];
            const double un = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
            for (int j = 0; j < 3; ++j) s.UB[3 * (size_t)b + j] = u[j] - n[j] * un;
            continue;
        }
        if (s.bcU[pi] == OR_BC_WEDGE) {
            double T[9], ub[3];
            wedgeFaceT(&s.nf[3 * (size_t)f], T);
            matVec(T, &s.U[3 * (size_t)P], ub);
            for (int j = 0; j < 3; ++j) s.UB[3 * (size_t)b + j] = ub[j];
            continue;
        }
        for (int j = 0; j < 3; ++j)
            s.UB[3 * (size_t)b + j] = (s.bcU[pi] == OR_BC_FIXED_VALUE) ? s.bvU[3 * (size_t)b + j] : s.U[3 * (size_t)P + j];
    }
}
// e: fixedEnergy (T fixesValue) | gradientEnergy (T zeroGradient -> gradient 0 for constant Cv) [OF-v2312]
void correctE(or_ctx& s)
{
    const Thermo th = thermoOf(s);
    for (int b = 0; b < s.nBnd; ++b) {
        const int pi = s.bfacePatch[b];
        if (s.patchKind[pi] == OR_PATCH_EMPTY) continue;
        const int P = s.owner[s.nInternal + b];
        if (s.bcT[pi] == OR_BC_FIXED_VALUE) s.eB[b] = th.HE(s.pB[b], s.bvT[b]);
        else s.eB[b] = s.e[P] + 0.0 / s.dC[s.nInternal + b];
    }
}
// p: fixedValue | zeroGradient | qgdFlux (qgdFluxFvPatchScalarField.C:159-208 + fixedGradient evaluate)
void correctP(or_ctx& s)
{
    for (int b = 0; b < s.nBnd; ++b) {
        const int pi = s.bfacePatch[b];
        if (s.patchKind[pi] == OR_PATCH_EMPTY) continue;
        const int f = s.nInternal + b;
        const int P = s.owner[f];
        if (s.bcP[pi] == OR_BC_FIXED_VALUE) s.pB[b] = s.bvP[b];
        else if (s.bcP[pi] == OR_BC_ZERO_GRADIENT) s.pB[b] = s.p[P];
        else if (s.bcP[pi] == OR_BC_QGD_FLUX) {
            if (s.havePhiwStar) {
                const double fluxSnGrad = s.phiw[f] / s.tauQGDf[f] / s.magSf[f];           // :184-191
                s.pGrad[b] = -fluxSnGrad;                                                  // :192
            }
            s.pB[b] = s.p[P] + s.pGrad[b] / s.dC[f];                                       // fixedGradient::evaluate
        }
    }
}

void patchSnGrad(const or_ctx& s, int k, const IVec& bc, const double* cell, const double* bnd, const double* grad, double* out);

// varScModel6::correct varScModel6.C:210-269 | varScModel7::correct varScModel7.C:176-254 : the pressure-jump sensor
//   ScQGD_c = cSc1 * |sum_f (+-) dpf| / (sum_f pf / n),  pf = linearInterpolate(p), dpf = fvc::snGrad(p)/deltaCoeffs,
// summed over mesh.cells()[c] = the faces the cell owns, then the faces it is the neighbour of, each ascending
// [OF-v2312 primitiveMesh::calcCells]; empty and wedge patch faces are skipped; p is the field as it stands when
// thermo.correct() runs (QGDFoam.C:149: the old-step pressure).  The boundary ScQGD stays the dictionary value
// (calculated patches, QGDCoeffs.C:249-261), clamped with the internal field by model 7's min/max.
void varScCorrect(or_ctx& s)
{
    const int model = s.prm.qgdModel;
    Vec pf(s.nFaces), sn(s.nFaces), bsg(s.nBnd);
    linearInterpolate(s, 1, s.p.data(), s.pB.data(), nullptr, pf.data());                   // :210-211 | :176-177
    patchSnGrad(s, 1, s.bcP, s.p.data(), s.pB.data(), s.pGrad.data(), bsg.data());
    snGrad(s, 1, s.p.data(), bsg.data(), sn.data());
    for (int f = 0; f < s.nFaces; ++f) sn[f] = sn[f] / s.dC[f];                             // :212-213 | :178-179
    const double cSc1 = (model == 7) ? s.prm.varScCSc1 : 1.0;
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) {                                                   // :215-269 | :181-235
        double sumDpF = 0.0, sumpf = 0.0, n = 0.0;
        for (int pass = 0; pass < 2; ++pass)
            for (int q = s.cfOff[c]; q < s.cfOff[c + 1]; ++q) {
                const int f = s.cfFace[q];
                const bool own = (s.owner[f] == c);
                if (own != (pass == 0)) continue;
                if (f < s.nInternal) {
                    sumpf += pf[f];
                    if (own) sumDpF += sn[f]; else sumDpF -= sn[f];
                    n = n + 1;
                } else {
                    const int b = f - s.nInternal;
                    if (patchIsEmpty(s, b) || patchIsWedge(s, b)) continue;
                    sumDpF += sn[f];                    // processor patches never reach the serial oracle (extended sub-meshes)
                    sumpf += pf[f];
                    n = n + 1;
                }
            }
        sumpf /= n;
        s.ScQGD[c] = cSc1 * (std::fabs(sumDpF) / sumpf);
    }
    if (model == 7) {
        if (s.prm.varScMinSc >= 0) {                                                       // :237-240
            for (double& v : s.ScQGD) v = std::max(v, s.prm.varScMinSc);
            for (double& v : s.ScQGDB) v = std::max(v, s.prm.varScMinSc);
        }
        if (s.prm.varScMaxSc >= 0) {                                                       // :241-244
            for (double& v : s.ScQGD) v = std::min(v, s.prm.varScMaxSc);
            for (double& v : s.ScQGDB) v = std::min(v, s.prm.varScMaxSc);
        }
        for (int c : s.constScCells) s.ScQGD[c] = s.prm.ScQGD;                             // :246-254, constSc_ = ScQGD :146
    }
}

// ---- [OF-v2312] primitiveMeshTools::cellClosedness, the aspect-ratio part (restated from the OpenFOAM source as remembered;
// unverified like every [OF] item): per cell the sums of |Sf| components over all its faces; aspect ratio = max/min of the
// sums over the solved directions, in 3D also at least (1/6) sum(|Sf| components) / V^(2/3) (1 for a cube)
void cellAspectRatio(const or_ctx& m, Vec& aratio)
{
    Vec sumMag(3 * (size_t)m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f)                       // forAll(own, facei): every face adds to its owner
        for (int d = 0; d < 3; ++d) sumMag[3 * (size_t)m.owner[f] + d] += std::fabs(m.Sf[3 * (size_t)f + d]);
    for (int f = 0; f < m.nInternal; ++f)                    // forAll(nei, facei)
        for (int d = 0; d < 3; ++d) sumMag[3 * (size_t)m.neighbour[f] + d] += std::fabs(m.Sf[3 * (size_t)f + d]);
    const double ROOTVSMALL = 1.0e-150, VGREAT = 1.0e+300;
    aratio.assign(m.nCells, 1.0);
    for (int c = 0; c < m.nCells; ++c) {
        double minC = VGREAT, maxC = -VGREAT;
        for (int d = 0; d < 3; ++d)
            if (m.gD[d] == 1) { minC = std::min(minC, sumMag[3 * (size_t)c + d]); maxC = std::max(maxC, sumMag[3 * (size_t)c + d]); }
        double ar = maxC / (minC + ROOTVSMALL);
        if (m.nD == 3) {
            const double v = std::max(ROOTVSMALL, m.V[c]);
            ar = std::max(ar, 1.0 / 6.0 * (sumMag[3 * (size_t)c] + sumMag[3 * (size_t)c + 1] + sumMag[3 * (size_t)c + 2]) / std::pow(v, 2.0 / 3.0));
        }
        aratio[c] = ar;
    }
}

// varScModel5.C:112-132 : cqSc = badQualitySc * aspectRatio / maxAspectRatio where aspectRatio > maxAspectRatio, else 0
void varSc5CellQuality(const or_ctx& m, double badQualitySc, double thr, Vec& cqSc, Vec* aratioOut = nullptr)
{
    Vec ar;
    cellAspectRatio(m, ar);
    cqSc.assign(m.nCells, 0.0);
    for (int c = 0; c < m.nCells; ++c)
        if (ar[c] > thr) cqSc[c] = badQualitySc * ar[c] / thr;
    if (aratioOut) *aratioOut = ar;
}

// ---- [OF-v2312] fvc::smooth(field, coeff)  (finiteVolume/fvc/fvcSmooth/smooth.C + smoothData + FaceCellWave), restated
// sequentially in the reference's visiting order, because the result depends on it inside the 1 % propagation tolerance:
//   * initial changed faces: internal faces, ascending, whose two cell values differ by more than maxRatio = 1 + coeff; the
//     face carries the larger value (no coupled patches on a serial mesh);
//   * faceToCell: the changed faces in list order, owner then neighbour; a cell that differs from the face value takes
//     smoothData::updateCell = update(scale = maxRatio): an unset / ~zero cell copies the face value, otherwise the cell is
//     raised to faceValue/maxRatio when faceValue > (1 + tol) maxRatio cellValue (tol = FaceCellWave::propagationTol_ = 0.01);
//     a cell enters the changed-cell list the first time it changes;
//   * cellToFace: the changed cells in list order, their faces in mesh.cells() order (owned faces ascending, then the faces
//     the cell is the neighbour of); smoothData::updateFace = update(scale = 1); boundary faces take part (they receive the
//     owner's value and later offer it back, which can never raise the owner, so they only lengthen the lists);
//   * until a sweep changes nothing.  Returns the number of completed iterations.
int fvcSmooth(const or_ctx& m, double* field, double coeff)
{
    const double maxRatio = 1.0 + coeff, tol = 0.01, SMALL_ = 1.0e-15, VSMALL_ = 1.0e-300, GREAT_ = 1.0e+15;
    const int nC = m.nCells, nF = m.nFaces, nI = m.nInternal;
    std::vector<double> cellV(field, field + nC), faceV(nF, -GREAT_);          // smoothData() : value_(-GREAT)
    auto valid = [&](double v) { return v > -SMALL_; };
    auto update = [&](double& mine, double other, double scale) {            // smoothDataI.H update()
        if (!valid(mine) || mine < VSMALL_) { mine = other; return true; }
        if (other > (1 + tol) * scale * mine) { mine = other / scale; return true; }
        return false;
    };
    std::vector<int> chF, chC;
    std::vector<char> inF(nF, 0), inC(nC, 0);
    for (int f = 0; f < nI; ++f) {                                             // smooth.C: initial field on faces
        const int own = m.owner[f], nbr = m.neighbour[f];
        double v;
        if (field[own] > maxRatio * field[nbr]) v = field[own];
        else if (field[nbr] > maxRatio * field[own]) v = field[nbr];
        else continue;
        faceV[f] = v; inF[f] = 1; chF.push_back(f);                            // FaceCellWave::setFaceInfo
    }
    // mesh.cells()[c]: owned faces ascending, then neighbour-side faces ascending [OF-v2312 primitiveMesh::calcCells]
    auto forCellFaces = [&](int c, auto fn) {
        for (int q = m.cfOff[c]; q < m.cfOff[c + 1]; ++q) if (m.owner[m.cfFace[q]] == c) fn(m.cfFace[q]);
        for (int q = m.cfOff[c]; q < m.cfOff[c + 1]; ++q) if (m.owner[m.cfFace[q]] != c) fn(m.cfFace[q]);
    };
    int iter = 0;
    while (true) {
        // faceToCell
        for (int f : chF) {
            const double fv = faceV[f];
            const int cells[2] = {m.owner[f], f < nI ? m.neighbour[f] : -1};
            for (int c : cells) {
                if (c < 0) continue;
                if (cellV[c] == fv) continue;                                  // !currInfo.equal(newInfo)
                if (update(cellV[c], fv, maxRatio) && !inC[c]) { inC[c] = 1; chC.push_back(c); }
            }
            inF[f] = 0;
        }
        chF.clear();
        if (chC.empty()) break;
        // cellToFace
        for (int c : chC) {
            const double cv = cellV[c];
            forCellFaces(c, [&](int f) {
                if (faceV[f] == cv) return;
                if (update(faceV[f], cv, 1.0) && !inF[f]) { inF[f] = 1; chF.push_back(f); }
            });
            inC[c] = 0;
        }
        chC.clear();
        if (chF.empty()) break;
        ++iter;
    }
    std::copy(cellV.begin(), cellV.end(), field);
    return iter;
}

void gaussGradCells(const or_ctx& s, int k, const double* phif, double* out);
// varScModel5::correct varScModel5.C:197-232 : the density-gradient sensor, relaxed against the previous ScQGD, clamped, floored
// by the mesh-quality value, then smoothed with fvc::smooth.  rho = qgdThermo.rho() = psi*p [OF-v2312 psiThermo::rho] with the
// psi just computed and the pressure of the old step (thermo.correct() runs before p = rho/psi, QGDFoam.C:149-154), boundary
// values psi_b*p_b.  fvc::grad(rho) [OF-v2312 Gauss linear]: cell values (1/V) sum_f Sf rho_f; boundary values (they enter the
// boundary ScQGD) gaussGrad::correctBoundaryConditions: g_P + n (snGrad_b - n.g_P) with the snGrad of a calculated patch,
// deltaCoeffs (rho_b - rho_P).  ScQGD is a calculated-patch field (QGDCoeffs.C:249-261): its boundary values follow the same
// algebra and the clamps; cqSc, the cell set and fvc::smooth act on the cells only (varScModel5.C:219-232).
void varSc5Correct(or_ctx& s)
{
    const int nC = s.nCells, nB = s.nBnd;
    const double rC = s.prm.varSc5RC;
    Vec rho(nC), rhoB(nB, 0.0), rhof(s.nFaces), g(3 * (size_t)nC), gB(3 * (size_t)nB, 0.0);
    for (int c = 0; c < nC; ++c) rho[c] = s.p[c] * s.psi[c];                               // :202
    for (int b = 0; b < nB; ++b) if (!patchIsEmpty(s, b)) rhoB[b] = s.pB[b] * s.psiB[b];
    linearInterpolate(s, 1, rho.data(), rhoB.data(), nullptr, rhof.data());
    gaussGradCells(s, 1, rhof.data(), g.data());
    for (int b = 0; b < nB; ++b) {
        if (patchIsEmpty(s, b)) continue;
        const int f = s.nInternal + b, P = s.owner[f];
        const double* n = &s.nf[3 * (size_t)f];
        const double sn = s.dC[f] * (rhoB[b] - rho[P]);
        const double nG = n[0] * g[3 * (size_t)P] + n[1] * g[3 * (size_t)P + 1] + n[2] * g[3 * (size_t)P + 2];
        for (int i = 0; i < 3; ++i) gB[3 * (size_t)b + i] = g[3 * (size_t)P + i] + n[i] * (sn - nG);
    }
    for (int c = 0; c < nC; ++c) {                                                         // :209-212
        const double mg = std::sqrt(g[3 * (size_t)c] * g[3 * (size_t)c] + g[3 * (size_t)c + 1] * g[3 * (size_t)c + 1] + g[3 * (size_t)c + 2] * g[3 * (size_t)c + 2]);
        s.ScQGD[c] = rC * (mg * s.hQGD[c] / rho[c]) + (1.0 - rC) * s.ScQGD[c];
    }
    for (int b = 0; b < nB; ++b) {
        if (patchIsEmpty(s, b)) continue;
        const double mg = std::sqrt(gB[3 * (size_t)b] * gB[3 * (size_t)b] + gB[3 * (size_t)b + 1] * gB[3 * (size_t)b + 1] + gB[3 * (size_t)b + 2] * gB[3 * (size_t)b + 2]);
        s.ScQGDB[b] = rC * (mg * s.hQGDB[b] / rhoB[b]) + (1.0 - rC) * s.ScQGDB[b];
    }
    for (double& v : s.ScQGD) v = std::min(std::max(v, s.prm.varScMinSc), s.prm.varScMaxSc);          // :214-217
    for (int b = 0; b < nB; ++b) if (!patchIsEmpty(s, b)) s.ScQGDB[b] = std::min(std::max(s.ScQGDB[b], s.prm.varScMinSc), s.prm.varScMaxSc);
    for (int c = 0; c < nC; ++c) s.ScQGD[c] = std::max(s.ScQGD[c], s.cqSc[c]);             // :219-220
    for (int c : s.constScCells) s.ScQGD[c] = s.prm.ScQGD;                                 // :222-230, constSc_ = ScQGD :137
    fvcSmooth(s, s.ScQGD.data(), s.prm.varSc5SmoothCoeff);                                 // :232 (calculated patches: boundary unchanged)
}

// hePsiQGDThermo::calculate  hePsiQGDThermo.C:37-126 ; constScPrModel1::correct constScPrModel1.C:97-131 ;
// QGDThermo::correctQGD QGDThermo.C:84-111
void thermoCorrect(or_ctx& s)
{
    const Thermo th = thermoOf(s);
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) {                                                   // :48-64
        s.T[c] = th.THE(s.e[c], s.p[c], s.T[c]);
        s.psi[c] = th.psi(s.p[c], s.T[c]);
        s.mu[c] = th.muF(s.p[c], s.T[c]);
        s.alpha[c] = th.alphah(s.p[c], s.T[c]);
    }
    for (int b = 0; b < s.nBnd; ++b) {                                                     // :84-121
        const int pi = s.bfacePatch[b];
        if (s.patchKind[pi] == OR_PATCH_EMPTY) continue;
        if (s.bcT[pi] == OR_BC_FIXED_VALUE) { s.TB[b] = s.bvT[b]; s.eB[b] = th.HE(s.pB[b], s.TB[b]); }
        else s.TB[b] = th.THE(s.eB[b], s.pB[b], s.TB[b]);
        s.psiB[b] = th.psi(s.pB[b], s.TB[b]);
        s.muB[b] = th.muF(s.pB[b], s.TB[b]);
        s.alphaB[b] = th.alphah(s.pB[b], s.TB[b]);
    }
    const double g = th.gamma();                                                           // :123  Cp/Cv
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) { s.gamma[c] = g; s.c[c] = std::sqrt(s.gamma[c] / s.psi[c]); }   // :124
    for (int b = 0; b < s.nBnd; ++b) { s.gammaB[b] = g; s.cB[b] = patchIsEmpty(s, b) ? 1.0 : std::sqrt(s.gammaB[b] / s.psiB[b]); }
    // ---- QGDCoeffs::correct : constScPrModel1.C:97-131 | constScPrModel1n.C:98-156 | constScPrModel2.C:75-113
    {
        Vec aByC(s.nCells), aByCB(s.nBnd);
        const int model = s.prm.qgdModel;
        if (model == 1 && s.haveU) {                                                       // constScPrModel1n.C:107-129
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
            for (int c = 0; c < s.nCells; ++c) {
                const double* u = &s.U[3 * (size_t)c];
                s.tauQGD[c] = s.aQGD[c] * s.hQGD[c] / (std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) + s.c[c]);
            }
            for (int b = 0; b < s.nBnd; ++b) {
                if (patchIsEmpty(s, b)) continue;
                const double* u = &s.UB[3 * (size_t)b];
                s.tauQGDB[b] = s.aQGDB[b] * s.hQGDB[b] / (std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) + s.cB[b]);
            }
            linearInterpolate(s, 1, s.tauQGD.data(), s.tauQGDB.data(), nullptr, s.tauQGDf.data());
        } else {
            if (model == 5) {                                                              // varScModel5.C:204-205
                Vec af(s.nFaces), cf(s.nFaces);
                linearInterpolate(s, 1, s.aQGD.data(), s.aQGDB.data(), nullptr, af.data());
                linearInterpolate(s, 1, s.c.data(), s.cB.data(), nullptr, cf.data());
                for (int f = 0; f < s.nFaces; ++f) {
                    if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) { s.tauQGDf[f] = 0.0; continue; }
                    s.tauQGDf[f] = af[f] / cf[f] * s.hQGDf[f];
                }
            } else if (model == 1) {                                                       // constScPrModel1n.C:104-105
                Vec af(s.nFaces), cf(s.nFaces);
                linearInterpolate(s, 1, s.aQGD.data(), s.aQGDB.data(), nullptr, af.data());
                linearInterpolate(s, 1, s.c.data(), s.cB.data(), nullptr, cf.data());
                for (int f = 0; f < s.nFaces; ++f) {
                    if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) { s.tauQGDf[f] = 0.0; continue; }
                    s.tauQGDf[f] = af[f] * s.hQGDf[f] / cf[f];
                }
            } else {
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
                for (int c = 0; c < s.nCells; ++c) aByC[c] = s.aQGD[c] / s.c[c];
                for (int b = 0; b < s.nBnd; ++b) aByCB[b] = s.aQGDB[b] / s.cB[b];
                linearInterpolate(s, 1, aByC.data(), aByCB.data(), nullptr, s.tauQGDf.data());
                for (int f = 0; f < s.nFaces; ++f) s.tauQGDf[f] *= s.hQGDf[f];             // constScPrModel1.C:103
            }
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
            for (int c = 0; c < s.nCells; ++c) s.tauQGD[c] = s.aQGD[c] * s.hQGD[c] / s.c[c];   // :104
            for (int b = 0; b < s.nBnd; ++b) if (!patchIsEmpty(s, b)) s.tauQGDB[b] = s.aQGDB[b] * s.hQGDB[b] / s.cB[b];
        }
        if (model == 6 || model == 7) varScCorrect(s);
        if (model == 5) varSc5Correct(s);
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < s.nCells; ++c) {
            s.muQGD[c] = s.p[c] * s.ScQGD[c] * s.tauQGD[c];                                // :108-111
            s.alphauQGD[c] = s.muQGD[c] / s.PrQGD[c];                                      // :113-114
        }
        for (int b = 0; b < s.nBnd; ++b) {
            if (patchIsEmpty(s, b)) continue;
            s.muQGDB[b] = s.pB[b] * s.ScQGDB[b] * s.tauQGDB[b];                            // :121-124
            s.alphauQGDB[b] = s.muQGDB[b] / s.PrQGDB[b];
        }
        if (model == 2) {                                                                  // constScPrModel2.C:112 (mu: molecular, not yet + muQGD)
            for (int c = 0; c < s.nCells; ++c) s.tauQGD[c] += s.mu[c] / (s.p[c] * s.ScQGD[c]);
            for (int b = 0; b < s.nBnd; ++b) if (!patchIsEmpty(s, b)) s.tauQGDB[b] += s.muB[b] / (s.pB[b] * s.ScQGDB[b]);
        }
    }
    // ---- correctQGD
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) { s.mu[c] += s.muQGD[c]; s.alpha[c] += s.alphauQGD[c]; }
    for (int b = 0; b < s.nBnd; ++b) { s.muB[b] += s.muQGDB[b]; s.alphaB[b] += s.alphauQGDB[b]; }
}

// fvc::div(ssf) = fvc::surfaceIntegrate [OF-v2312]: owner +=, neighbour -=, boundary += faceCells, /V.
// Evaluated cell-wise over the ascending cell->face list: bit-identical to the sequential face loop.
void fvcDiv(const or_ctx& s, int k, const double* ssf, double* out)
{
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) {
        double acc[3] = {0, 0, 0};
        for (int q = s.cfOff[c]; q < s.cfOff[c + 1]; ++q) {
            const int f = s.cfFace[q];
            if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) continue;
            const bool own = (s.owner[f] == c);
            for (int j = 0; j < k; ++j) { if (own) acc[j] += ssf[(size_t)f * k + j]; else acc[j] -= ssf[(size_t)f * k + j]; }
        }
        for (int j = 0; j < k; ++j) out[(size_t)c * k + j] = acc[j] / s.V[c];
    }
}

template <class F> void forFaces(const or_ctx& s, F fn)
{
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int f = 0; f < s.nFaces; ++f) {
        if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) continue;
        fn(f);
    }
}

// [OF-v2312] fvc::grad, Gauss linear, cell values only: (1/V) sum_f Sf (x) phi_f  (owner +, neighbour -)
void gaussGradCells(const or_ctx& s, int k, const double* phif, double* out /*nCells*3k: index 3... [i*k+j] = d_i phi_j*/)
{
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < s.nCells; ++c) {
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int q = s.cfOff[c]; q < s.cfOff[c + 1]; ++q) {
            const int f = s.cfFace[q];
            if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) continue;
            const double sgn = (s.owner[f] == c) ? 1.0 : -1.0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < k; ++j) acc[i * k + j] += sgn * (s.Sf[3 * (size_t)f + i] * phif[(size_t)f * k + j]);
        }
        for (int q = 0; q < 3 * k; ++q) out[(size_t)c * 3 * k + q] = acc[q] / s.V[c];
    }
}

// [OF-v2312] gaussGrad::correctBoundaryConditions: boundary value of fvc::grad(vsf) on non-coupled patches,
//   gradB = grad_P + n (x) (snGrad_b - n . grad_P)
void gaussGradBoundary(const or_ctx& s, const double* gradCells, const double* sn3 /*nBnd*3 patch snGrad of the vector*/, double* gradB)
{
    for (int b = 0; b < s.nBnd; ++b) {
        if (patchIsEmpty(s, b)) { for (int q = 0; q < 9; ++q) gradB[9 * (size_t)b + q] = 0.0; continue; }
        const int f = s.nInternal + b; const int P = s.owner[f];
        const double* n = &s.nf[3 * (size_t)f];
        for (int j = 0; j < 3; ++j) {
            double nG = 0.0;
            for (int i = 0; i < 3; ++i) nG += n[i] * gradCells[9 * (size_t)P + 3 * i + j];
            for (int i = 0; i < 3; ++i) gradB[9 * (size_t)b + 3 * i + j] = gradCells[9 * (size_t)P + 3 * i + j] + n[i] * (sn3[3 * (size_t)b + j] - nG);
        }
    }
}

// QGDFoam/updateFields.H:45-80
void updateFields(or_ctx& s)
{
    const int nC = s.nCells, nB = s.nBnd;
    linearInterpolate(s, 1, s.rho.data(), s.rhoB.data(), nullptr, s.rhof.data());          // :45
    linearInterpolate(s, 3, s.U.data(), s.UB.data(), nullptr, s.Uf.data());                // :48
    linearInterpolate(s, 3, s.rhoU.data(), s.rhoUB.data(), nullptr, s.rhoUf.data());       // :51
    {   // :55  U*rhoU (outer product), then interpolate
        Vec t((size_t)nC * 9), tB((size_t)nB * 9);
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c)
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[(size_t)c * 9 + 3 * i + j] = s.U[3 * (size_t)c + i] * s.rhoU[3 * (size_t)c + j];
        for (int b = 0; b < nB; ++b)
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tB[(size_t)b * 9 + 3 * i + j] = s.UB[3 * (size_t)b + i] * s.rhoUB[3 * (size_t)b + j];
        linearInterpolate(s, 9, t.data(), tB.data(), nullptr, s.UrhoUf.data());
    }
    linearInterpolate(s, 1, s.p.data(), s.pB.data(), nullptr, s.pf.data());                // :58
    linearInterpolate(s, 1, s.c.data(), s.cB.data(), nullptr, s.cf.data());                // :61
    linearInterpolate(s, 1, s.gamma.data(), s.gammaB.data(), nullptr, s.gammaf.data());    // :64
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
    for (int c = 0; c < nC; ++c) s.H[c] = (s.rhoE[c] + s.p[c]) / s.rho[c];                 // :71
    for (int b = 0; b < nB; ++b) s.HB[b] = patchIsEmpty(s, b) ? 0.0 : (s.rhoEB[b] + s.pB[b]) / s.rhoB[b];
    linearInterpolate(s, 1, s.H.data(), s.HB.data(), nullptr, s.Hf.data());                // :72
    {   // :79-80  turbulence->alphaEff(), muEff() for the laminar model [OF-v2312, SURVEY 8c item 10]
        Vec ae(nC), aeB(nB);
        for (int c = 0; c < nC; ++c) ae[c] = s.prm.alphaEffGammaFactor ? s.gamma[c] * s.alpha[c] : s.alpha[c];
        for (int b = 0; b < nB; ++b) aeB[b] = s.prm.alphaEffGammaFactor ? s.gammaB[b] * s.alphaB[b] : s.alphaB[b];
        linearInterpolate(s, 1, ae.data(), aeB.data(), nullptr, s.alphauf.data());
        linearInterpolate(s, 1, s.mu.data(), s.muB.data(), nullptr, s.muf.data());
    }
}

// QGDFoam/updateFluxes.H:41-139 (explicit branch)
void updateFluxes(or_ctx& s)
{
    const int nB = s.nBnd;
    Vec sg3((size_t)nB * 3), sg1(nB);
    const bool gvp = (s.scheme == OR_FVSC_GAUSSVOLPOINT);
    // :41  gradUf = fvsc::grad(U)   (GaussVolPoint::Grad first calls U.correctBoundaryConditions())
    if (gvp) correctU(s);
    patchSnGrad(s, 3, s.bcU, s.U.data(), s.UB.data(), nullptr, sg3.data());
    fvscGrad(s, s.scheme, 3, s.U.data(), s.UB.data(), sg3.data(), nullptr, s.gradUf.data());
    forFaces(s, [&](int f) { const double* g = &s.gradUf[(size_t)f * 9]; s.divUf[f] = g[0] + g[4] + g[8]; });   // :43
    // :45  gradef
    if (gvp) correctE(s);
    {   // e patch snGrad: fixedEnergy -> generic; gradientEnergy -> gradient() = 0
        IVec bcE(s.nPatches);
        for (int pi = 0; pi < s.nPatches; ++pi) bcE[pi] = (s.bcT[pi] == OR_BC_FIXED_VALUE) ? OR_BC_FIXED_VALUE : OR_BC_ZERO_GRADIENT;
        patchSnGrad(s, 1, bcE, s.e.data(), s.eB.data(), nullptr, sg1.data());
    }
    fvscGrad(s, s.scheme, 1, s.e.data(), s.eB.data(), sg1.data(), nullptr, s.gradef.data());
    // :47  gradRhof  (rho has calculated patches: correctBoundaryConditions is a no-op)
    {
        IVec bcR(s.nPatches, OR_BC_CALCULATED);
        patchSnGrad(s, 1, bcR, s.rho.data(), s.rhoB.data(), nullptr, sg1.data());
    }
    fvscGrad(s, s.scheme, 1, s.rho.data(), s.rhoB.data(), sg1.data(), nullptr, s.gradRhof.data());
    // :54-61  rhoW = tau*( ((Uf*gradRhof) & Uf) + rhoUf*divUf + (rhoUf & gradUf) )
    forFaces(s, [&](int f) {
        const double* Uf = &s.Uf[3 * (size_t)f]; const double* gR = &s.gradRhof[3 * (size_t)f];
        const double* rU = &s.rhoUf[3 * (size_t)f]; const double* G = &s.gradUf[9 * (size_t)f];
        const double gRU = gR[0] * Uf[0] + gR[1] * Uf[1] + gR[2] * Uf[2];    // (Uf*gradRhof)&Uf = Uf (gradRhof . Uf)
        for (int j = 0; j < 3; ++j) {
            const double rUG = rU[0] * G[j] + rU[1] * G[3 + j] + rU[2] * G[6 + j];   // (rhoUf & gradUf)_j
            s.rhoW[3 * (size_t)f + j] = s.tauQGDf[f] * (Uf[j] * gRU + rU[j] * s.divUf[f] + rUG);
        }
        s.phiw[f] = s.Sf[3 * (size_t)f] * s.rhoW[3 * (size_t)f] + s.Sf[3 * (size_t)f + 1] * s.rhoW[3 * (size_t)f + 1]
                  + s.Sf[3 * (size_t)f + 2] * s.rhoW[3 * (size_t)f + 2];                         // :63
    });
    s.havePhiwStar = true;
    // :65  gradPf = fvsc::grad(p)   (p.correctBoundaryConditions -> qgdFlux reads phiwStar, tauQGDf)
    if (gvp) correctP(s);
    patchSnGrad(s, 1, s.bcP, s.p.data(), s.pB.data(), s.pGrad.data(), sg1.data());
    fvscGrad(s, s.scheme, 1, s.p.data(), s.pB.data(), sg1.data(), nullptr, s.gradPf.data());
    if (s.prm.implicitDiffusion) {
        // :107-111  tauMC = qgdInterpolate(muEff*dev2(T(fvc::grad(U)))) ; phiTauMC = Sf & tauMC
        const int nC = s.nCells;
        Vec gU(9 * (size_t)nC), gUB(9 * (size_t)nB), t(9 * (size_t)nC), tB(9 * (size_t)nB), snU(3 * (size_t)nB);
        patchSnGrad(s, 3, s.bcU, s.U.data(), s.UB.data(), nullptr, snU.data());
        gaussGradCells(s, 3, s.Uf.data(), gU.data());
        gaussGradBoundary(s, gU.data(), snU.data(), gUB.data());
        auto dev2T = [](double mu, const double* g, double* o) {       // mu * dev2(T(g)) ; dev2(A) = A - (2/3) tr(A) I
            const double tr = g[0] + g[4] + g[8];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[3 * i + j] = mu * (g[3 * j + i] - (2.0 / 3.0) * tr * (i == j ? 1.0 : 0.0));
        };
        for (int c = 0; c < nC; ++c) dev2T(s.mu[c], &gU[9 * (size_t)c], &t[9 * (size_t)c]);
        for (int b = 0; b < nB; ++b) dev2T(s.muB[b], &gUB[9 * (size_t)b], &tB[9 * (size_t)b]);
        linearInterpolate(s, 9, t.data(), tB.data(), nullptr, s.tauMC.data());
        forFaces(s, [&](int f) {
            for (int j = 0; j < 3; ++j) {
                double v = 0.0;
                for (int i = 0; i < 3; ++i) v += s.Sf[3 * (size_t)f + i] * s.tauMC[9 * (size_t)f + 3 * i + j];
                s.phiTauMC[3 * (size_t)f + j] = v;
            }
        });
    }
    forFaces(s, [&](int f) {
        const double* Sf = &s.Sf[3 * (size_t)f]; const double* Uf = &s.Uf[3 * (size_t)f];
        const double* gP = &s.gradPf[3 * (size_t)f]; const double* G = &s.gradUf[9 * (size_t)f];
        const double* UrU = &s.UrhoUf[9 * (size_t)f]; const double tau = s.tauQGDf[f];
        double* rhoW = &s.rhoW[3 * (size_t)f]; double* jm = &s.jm[3 * (size_t)f];
        for (int j = 0; j < 3; ++j) rhoW[j] += tau * gP[j];                               // :67
        for (int j = 0; j < 3; ++j) jm[j] = s.rhoUf[3 * (size_t)f + j] - rhoW[j];         // :69
        s.phiJm[f] = Sf[0] * jm[0] + Sf[1] * jm[1] + Sf[2] * jm[2];                       // :71
        s.phi[f] = Sf[0] * s.rhoUf[3 * (size_t)f] + Sf[1] * s.rhoUf[3 * (size_t)f + 1] + Sf[2] * s.rhoUf[3 * (size_t)f + 2];   // :72
        for (int j = 0; j < 3; ++j) s.phiJmU[3 * (size_t)f + j] = s.phiJm[f] * Uf[j];     // :78  qgdFlux -> flux*psif
        for (int j = 0; j < 3; ++j) s.phiP[3 * (size_t)f + j] = Sf[j] * s.pf[f];          // :79
        // :81-93  Pif
        double* Pi = &s.Pif[9 * (size_t)f];
        const double UgP = Uf[0] * gP[0] + Uf[1] * gP[1] + Uf[2] * gP[2];
        const double iso = UgP + (s.gammaf[f] * s.pf[f] * s.divUf[f]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double UrUG = UrU[3 * i] * G[j] + UrU[3 * i + 1] * G[3 + j] + UrU[3 * i + 2] * G[6 + j];   // (UrhoUf & gradUf)_ij
                Pi[3 * i + j] = tau * (UrUG + Uf[i] * gP[j]) + tau * ((i == j ? 1.0 : 0.0) * iso);
            }
        // :95-106  explicit branch: Navier-Stokes stress
        if (!s.prm.implicitDiffusion)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    Pi[3 * i + j] += s.muf[f] * (G[3 * i + j] + G[3 * j + i] - (2.0 / 3.0) * (i == j ? 1.0 : 0.0) * s.divUf[f]);
        for (int j = 0; j < 3; ++j) s.phiPi[3 * (size_t)f + j] = Sf[0] * Pi[j] + Sf[1] * Pi[3 + j] + Sf[2] * Pi[6 + j];   // :113
        s.phiJmH[f] = s.phiJm[f] * s.Hf[f];                                               // :119
        // :121-135  qf
        const double* ge = &s.gradef[3 * (size_t)f]; const double* gR = &s.gradRhof[3 * (size_t)f];
        const double pr2 = s.pf[f] / s.rhof[f] / s.rhof[f];
        double v[3], q[3];
        for (int j = 0; j < 3; ++j) v[j] = ge[j] - pr2 * gR[j];
        for (int i = 0; i < 3; ++i) q[i] = -tau * (UrU[3 * i] * v[0] + UrU[3 * i + 1] * v[1] + UrU[3 * i + 2] * v[2]);
        if (!s.prm.implicitDiffusion) for (int i = 0; i < 3; ++i) q[i] -= s.alphauf[f] * ge[i];              // :131-135
        for (int i = 0; i < 3; ++i) s.qf[3 * (size_t)f + i] = q[i];
        s.phiQ[f] = Sf[0] * q[0] + Sf[1] * q[1] + Sf[2] * q[2];                           // :137
        double PiU[3];
        for (int i = 0; i < 3; ++i) PiU[i] = Pi[3 * i] * Uf[0] + Pi[3 * i + 1] * Uf[1] + Pi[3 * i + 2] * Uf[2];   // Pif & Uf
        s.phiPiU[f] = Sf[0] * PiU[0] + Sf[1] * PiU[1] + Sf[2] * PiU[2];                   // :139
    });
}

double faceMax(const or_ctx& s, const Vec& v) { double r = -std::numeric_limits<double>::max(); for (int f = 0; f < s.nFaces; ++f) { if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) continue; r = std::max(r, v[f]); } return r; }
double faceMin(const or_ctx& s, const Vec& v) { double r = std::numeric_limits<double>::max(); for (int f = 0; f < s.nFaces; ++f) { if (f >= s.nInternal && patchIsEmpty(s, f - s.nInternal)) continue; r = std::min(r, v[f]); } return r; }

} // namespace

// ============================================================================
extern "C" {

or_ctx* or_create(const or_mesh_t* m, int nThreads)
{
    or_ctx* s = new or_ctx();
    s->nCells = m->nCells; s->nFaces = m->nFaces; s->nInternal = m->nInternal; s->nPoints = m->nPoints; s->nPatches = m->nPatches;
    const int nB = m->nFaces - m->nInternal;
    s->points.assign(m->points, m->points + 3 * (size_t)m->nPoints);
    s->faceOff.assign(m->faceOff, m->faceOff + m->nFaces + 1);
    s->faceVerts.assign(m->faceVerts, m->faceVerts + m->faceOff[m->nFaces]);
    s->owner.assign(m->owner, m->owner + m->nFaces);
    s->neighbour.assign(m->neighbour, m->neighbour + m->nInternal);
    s->patchStart.assign(m->patchStart, m->patchStart + m->nPatches);
    s->patchSize.assign(m->patchSize, m->patchSize + m->nPatches);
    s->patchKind.assign(m->patchKind, m->patchKind + m->nPatches);
    s->C.assign(m->C, m->C + 3 * (size_t)m->nCells);
    s->V.assign(m->V, m->V + m->nCells);
    s->Cf.assign(m->Cf, m->Cf + 3 * (size_t)m->nFaces);
    s->Sf.assign(m->Sf, m->Sf + 3 * (size_t)m->nFaces);
    s->magSf.assign(m->magSf, m->magSf + m->nFaces);
    s->w.assign(m->weights, m->weights + m->nFaces);
    s->dC.assign(m->deltaCoeffs, m->deltaCoeffs + m->nFaces);
    s->ndC.assign(m->nonOrthDeltaCoeffs, m->nonOrthDeltaCoeffs + m->nFaces);
    if (m->neighbCellCentres) s->nbrCC.assign(m->neighbCellCentres, m->neighbCellCentres + 3 * (size_t)nB);
    else s->nbrCC.assign(3 * (size_t)nB, 0.0);
    for (int d = 0; d < 3; ++d) s->gD[d] = m->geometricD[d];
    s->nThreads = nThreads > 0 ? nThreads : 1;
    buildDerived(*s);
    return s;
}

void or_destroy(or_ctx* s) { delete s; }

void or_get_hQGDf(or_ctx* s, double* out) { std::copy(s->hQGDf.begin(), s->hQGDf.end(), out); }
void or_get_hQGD(or_ctx* s, double* out) { std::copy(s->hQGD.begin(), s->hQGD.end(), out); }

void or_fvsc_grad(or_ctx* s, int scheme, int k, const double* cell, const double* bnd, const double* bsg, const double* nbr, double* out)
{ fvscGrad(*s, scheme, k, cell, bnd, bsg, nbr, out); }
void or_fvsc_div(or_ctx* s, int scheme, int k, const double* cell, const double* bnd, const double* bsg, const double* nbr, double* out)
{ fvscDiv(*s, scheme, k, cell, bnd, bsg, nbr, out); }
void or_vol_point_interpolate(or_ctx* s, int k, const double* cell, const double* bnd, double* out)
{ volPointInterpolate(*s, k, cell, bnd, out); }
void or_linear_interpolate(or_ctx* s, int k, const double* cell, const double* bnd, double* out)
{ linearInterpolate(*s, k, cell, bnd, nullptr, out); }

// QGDFoam/createFields.H:3-109, createFaceFields.H, createFaceFluxes.H ; psiQGDThermo.C:57-71 ; QGDCoeffs.C:163-264
void or_qgd_init(or_ctx* sp, const or_qgd_params_t* prm, int fvscScheme, const int* bcU, const int* bcT, const int* bcP,
                 const double* bvU, const double* bvT, const double* bvP, const double* U0, const double* T0,
                 const double* p0, const double* alphaQGD, double deltaT0)
{
    or_ctx& s = *sp;
    s.prm = *prm; s.scheme = fvscScheme; s.deltaT = deltaT0;
    const int nC = s.nCells, nB = s.nBnd, nF = s.nFaces;
    s.bcU.assign(bcU, bcU + s.nPatches); s.bcT.assign(bcT, bcT + s.nPatches); s.bcP.assign(bcP, bcP + s.nPatches);
    s.bvU.assign(bvU, bvU + 3 * (size_t)nB); s.bvT.assign(bvT, bvT + nB); s.bvP.assign(bvP, bvP + nB);
    auto z = [](Vec& v, size_t n) { v.assign(n, 0.0); };
    z(s.rho, nC); z(s.rhoB, nB); z(s.UB, 3 * (size_t)nB); z(s.pB, nB); z(s.e, nC); z(s.eB, nB); z(s.TB, nB);
    z(s.psi, nC); z(s.psiB, nB); z(s.mu, nC); z(s.muB, nB); z(s.alpha, nC); z(s.alphaB, nB);
    z(s.gamma, nC); z(s.gammaB, nB); z(s.c, nC); z(s.cB, nB);
    z(s.rhoU, 3 * (size_t)nC); z(s.rhoUB, 3 * (size_t)nB); z(s.rhoE, nC); z(s.rhoEB, nB); z(s.H, nC); z(s.HB, nB);
    z(s.tauQGD, nC); z(s.tauQGDB, nB); z(s.muQGD, nC); z(s.muQGDB, nB); z(s.alphauQGD, nC); z(s.alphauQGDB, nB);
    z(s.pGrad, nB);
    s.U.assign(U0, U0 + 3 * (size_t)nC); s.T.assign(T0, T0 + nC); s.p.assign(p0, p0 + nC);
    // QGDCoeffs ctor: alphaQGD (read or 0.5, zeroGradient), Sc/Pr from dict   QGDCoeffs.C:119-160, constScPrModel1.C:58-89
    if (alphaQGD) s.aQGD.assign(alphaQGD, alphaQGD + nC); else s.aQGD.assign(nC, 0.5);
    s.aQGDB.assign(nB, 0.0); s.hQGDB.assign(nB, 0.0);
    for (int b = 0; b < nB; ++b) { s.aQGDB[b] = s.aQGD[s.owner[s.nInternal + b]]; s.hQGDB[b] = s.hQGDf[s.nInternal + b]; }   // QGDCoeffs.C:373
    s.ScQGD.assign(nC, prm->ScQGD); s.ScQGDB.assign(nB, prm->ScQGD); s.PrQGD.assign(nC, prm->PrQGD); s.PrQGDB.assign(nB, prm->PrQGD);
    if (prm->qgdModel == 5) varSc5CellQuality(s, prm->varSc5BadQualitySc, prm->varSc5MaxAspectRatio, s.cqSc);   // varScModel5.C:112-132
    for (Vec* v : {&s.tauQGDf, &s.rhof, &s.pf, &s.cf, &s.gammaf, &s.Hf, &s.alphauf, &s.muf, &s.divUf, &s.phiw, &s.phiJm, &s.phi,
                   &s.phiJmH, &s.phiQ, &s.phiPiU, &s.phiSigmaDotU}) z(*v, nF);
    for (Vec* v : {&s.Uf, &s.rhoUf, &s.gradef, &s.gradRhof, &s.gradPf, &s.rhoW, &s.jm, &s.phiJmU, &s.phiP, &s.phiPi, &s.qf}) z(*v, 3 * (size_t)nF);
    for (Vec* v : {&s.UrhoUf, &s.gradUf, &s.Pif, &s.tauMC}) z(*v, 9 * (size_t)nF);
    z(s.phiTauMC, 3 * (size_t)nF);
    const Thermo th = thermoOf(s);
    // boundary values of the read fields T, p, U as given by their BCs
    for (int b = 0; b < nB; ++b) {
        const int pi = s.bfacePatch[b];
        if (s.patchKind[pi] == OR_PATCH_EMPTY) continue;
        const int P = s.owner[s.nInternal + b];
        s.TB[b] = (s.bcT[pi] == OR_BC_FIXED_VALUE) ? s.bvT[b] : s.T[P];
        s.pB[b] = (s.bcP[pi] == OR_BC_FIXED_VALUE) ? s.bvP[b] : s.p[P];   // zeroGradient / qgdFlux (gradient 0, value = internal)
    }
    correctU(s);
    // heThermo ctor: he = HE(p,T) in cells and on patches [OF-v2312 heThermo::init]
    for (int c = 0; c < nC; ++c) s.e[c] = th.HE(s.p[c], s.T[c]);
    for (int b = 0; b < nB; ++b) if (!patchIsEmpty(s, b)) s.eB[b] = th.HE(s.pB[b], s.TB[b]);
    // hePsiQGDThermo ctor: calculate() ; createFields.H:8 thermo.correct()
    s.havePhiwStar = false;
    s.haveU = false;                  // psiQGDThermo::New runs before U is read (createFields.H:3-24)
    thermoCorrect(s);
    thermoCorrect(s);
    s.haveU = true;
    // createFields.H:37-87
    for (int c = 0; c < nC; ++c) {
        s.rho[c] = s.psi[c] * s.p[c];                                                     // psiThermo::rho()
        for (int j = 0; j < 3; ++j) s.rhoU[3 * (size_t)c + j] = s.rho[c] * s.U[3 * (size_t)c + j];
        const double m2 = s.U[3 * (size_t)c] * s.U[3 * (size_t)c] + s.U[3 * (size_t)c + 1] * s.U[3 * (size_t)c + 1] + s.U[3 * (size_t)c + 2] * s.U[3 * (size_t)c + 2];
        s.rhoE[c] = s.rho[c] * s.e[c] + s.rho[c] * 0.5 * m2;
    }
    for (int b = 0; b < nB; ++b) {
        if (patchIsEmpty(s, b)) continue;
        s.rhoB[b] = s.psiB[b] * s.pB[b];
        for (int j = 0; j < 3; ++j) s.rhoUB[3 * (size_t)b + j] = s.rhoB[b] * s.UB[3 * (size_t)b + j];
        const double m2 = s.UB[3 * (size_t)b] * s.UB[3 * (size_t)b] + s.UB[3 * (size_t)b + 1] * s.UB[3 * (size_t)b + 1] + s.UB[3 * (size_t)b + 2] * s.UB[3 * (size_t)b + 2];
        s.rhoEB[b] = s.rhoB[b] * s.eB[b] + s.rhoB[b] * 0.5 * m2;
    }
    // createFaceFluxes.H:40-43  first fvsc::grad(p): p.correctBoundaryConditions() without phiwStar
    correctP(s);
    s.qgdReady = true;
}

double or_qgd_deltaT(or_ctx* s) { return s->deltaT; }
int or_fvc_smooth(or_ctx* s, double* field, double coeff) { return fvcSmooth(*s, field, coeff); }
void or_varsc5_cell_quality(or_ctx* s, double badQualitySc, double maxAspectRatio, double* cqSc, double* aspectRatio)
{
    Vec q, ar;
    varSc5CellQuality(*s, badQualitySc, maxAspectRatio, q, &ar);
    std::copy(q.begin(), q.end(), cqSc);
    if (aspectRatio) std::copy(ar.begin(), ar.end(), aspectRatio);
}
void or_qgd_set_const_sc_cells(or_ctx* s, const int* cells, int n) { s->constScCells.assign(cells, cells + n); }
void or_set_degenerate_faces(or_ctx* s, const int* faces, int n) { s->lsForcedDeg.assign(faces, faces + n); s->lsBuilt = false; }
void or_set_pcg_blocks(or_ctx* s, const int* cellBlock) { if (cellBlock) s->pcgBlocks.assign(cellBlock, cellBlock + s->nCells); else s->pcgBlocks.clear(); }
int or_pcg_solve_blocks(or_ctx* sp, const double* diag, const double* upper, const double* b, double* x, double tol,
                        double relTol, int maxIter, int precond, double* initRes, double* finalRes, const int* cellBlock);
void or_qgd_set_sources(or_ctx* s, const double* suRho, const double* suU, const double* suE)
{
    const size_t n = s->nCells;
    if (suRho) s->suRho.assign(suRho, suRho + n); else s->suRho.clear();
    if (suU) s->suU.assign(suU, suU + 3 * n); else s->suU.clear();
    if (suE) s->suE.assign(suE, suE + n); else s->suE.clear();
}

// QGDFoam.C:90-163
double or_qgd_step(or_ctx* sp, int nSteps, int adjustTimeStep, double maxCo, double maxDeltaT, double cTau)
{
    or_ctx& s = *sp;
    const int nC = s.nCells, nB = s.nBnd;
    double CoNum = -1.0;
    Vec rho0(nC), rhoU0(3 * (size_t)nC), U0(3 * (size_t)nC), rhoE0(nC), e0(nC);
    Vec d1(nC), d3(3 * (size_t)nC), d3b(3 * (size_t)nC), d3c(3 * (size_t)nC), d1b(nC), d1c(nC), d1d(nC);
    for (int step = 0; step < nSteps; ++step) {
        updateFields(s);                                                                  // :104
        updateFluxes(s);                                                                  // :111
        if (adjustTimeStep) {                                                             // QGDCourantNo.H:36-53
            Vec Cof(s.nFaces, 0.0);
            forFaces(s, [&](int f) {
                const double Unf = s.Uf[3 * (size_t)f] * (s.Sf[3 * (size_t)f] / s.magSf[f]) + s.Uf[3 * (size_t)f + 1] * (s.Sf[3 * (size_t)f + 1] / s.magSf[f])
                                 + s.Uf[3 * (size_t)f + 2] * (s.Sf[3 * (size_t)f + 2] / s.magSf[f]);
                Cof[f] = std::max(std::fabs(Unf + s.cf[f]), std::fabs(Unf - s.cf[f])) * s.deltaT / s.hQGDf[f];
            });
            CoNum = faceMax(s, Cof);
            // setDeltaT-QGDQHD.H:41-61
            const double maxDeltaTFact = maxCo / (CoNum + SMALL);
            const double deltaTFact = std::min(std::min(maxDeltaTFact, 1.0 + 0.1 * maxDeltaTFact), 1.2);
            double maxDeltaT1 = cTau * faceMin(s, s.tauQGDf);
            maxDeltaT1 = std::min(maxDeltaT, maxDeltaT1);
            s.deltaT = std::min(deltaTFact * s.deltaT, maxDeltaT1);
        }
        const double rDeltaT = 1.0 / s.deltaT;
        rho0 = s.rho; rhoU0 = s.rhoU; U0 = s.U; rhoE0 = s.rhoE; e0 = s.e;                 // :127-131
        // ---- QGDRhoEqn.H:40-47   fvm::ddt(rho) + fvc::div(phiJm) == 0  [OF-v2312 Euler ddt, diagonal solve]
        fvcDiv(s, 1, s.phiJm.data(), d1.data());
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) {
            const double diag = rDeltaT * s.V[c];
            double source = rDeltaT * rho0[c] * s.V[c] - s.V[c] * d1[c];
            if (!s.suRho.empty()) source += s.suRho[c];                                   // == rhoSu  QGDRhoEqn.H:46
            s.rho[c] = source / diag;
        }
        // ---- QGDUEqn.H:36-45
        fvcDiv(s, 3, s.phiJmU.data(), d3.data());
        fvcDiv(s, 3, s.phiP.data(), d3b.data());
        fvcDiv(s, 3, s.phiPi.data(), d3c.data());
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c)
            for (int j = 0; j < 3; ++j) {
                const size_t i = 3 * (size_t)c + j;
                const double diag = rDeltaT * s.V[c];
                double source = rDeltaT * rhoU0[i] * s.V[c];
                source -= s.V[c] * d3[i]; source -= s.V[c] * d3b[i]; source += s.V[c] * d3c[i];
                s.rhoU[i] = source / diag;
            }
        // :48-51
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) for (int j = 0; j < 3; ++j) s.U[3 * (size_t)c + j] = s.rhoU[3 * (size_t)c + j] / s.rho[c];
        correctU(s);
        if (s.prm.implicitDiffusion) {
            // :54-74  fvm::ddt(rho,U) - fvc::ddt(rho,U) - fvm::laplacian(muf,U) - fvc::div(phiTauMC) == 0, component-wise
            fvcDiv(s, 3, s.phiTauMC.data(), d3.data());
            Vec upper(std::max(s.nInternal, 1)), diagL(nC, 0.0), diag(nC), src(nC), x(nC);
            for (int f = 0; f < s.nInternal; ++f) {                          // -fvm::laplacian(muf,U) [OF gaussLaplacianScheme]
                const double up = s.ndC[f] * (s.muf[f] * s.magSf[f]);
                upper[f] = -up; diagL[s.owner[f]] += up; diagL[s.neighbour[f]] += up;
            }
            for (int j = 0; j < 3; ++j) {
                for (int c = 0; c < nC; ++c) {
                    const size_t i = 3 * (size_t)c + j;
                    diag[c] = rDeltaT * s.rho[c] * s.V[c] + diagL[c];
                    src[c] = rDeltaT * rho0[c] * U0[i] * s.V[c] + s.V[c] * (rDeltaT * (s.rho[c] * s.U[i] - rho0[c] * U0[i])) + s.V[c] * d3[i];
                    if (!s.suU.empty()) src[c] += s.suU[i];                                // == rhoUSu  QGDUEqn.H:62
                    x[c] = s.U[i];
                }
                for (int b = 0; b < nB; ++b) {                               // fixedValue: internalCoeffs / boundaryCoeffs
                    const int pi = s.bfacePatch[b];
                    if (s.patchKind[pi] == OR_PATCH_EMPTY || s.bcU[pi] != OR_BC_FIXED_VALUE) continue;
                    const int f = s.nInternal + b; const int P = s.owner[f];
                    const double gS = s.muf[f] * s.magSf[f];
                    diag[P] += gS * s.ndC[f];
                    src[P] += gS * (s.ndC[f] * s.UB[3 * (size_t)b + j]);
                }
                s.lastDiffIters[j] = or_pcg_solve_blocks(sp, diag.data(), upper.data(), src.data(), x.data(), s.prm.diffTol, s.prm.diffRelTol,
                                                  s.prm.diffMaxIter, s.prm.diffPrecond, nullptr, nullptr, s.pcgBlocks.empty() ? nullptr : s.pcgBlocks.data());
                for (int c = 0; c < nC; ++c) s.U[3 * (size_t)c + j] = x[c];
            }
            correctU(s);
            for (int c = 0; c < nC; ++c) for (int j = 0; j < 3; ++j) s.rhoU[3 * (size_t)c + j] = s.rho[c] * s.U[3 * (size_t)c + j];   // :70
            // :72-74  sigmaDotU = (muf*linearInterpolate(fvc::grad(U)) + tauMC) & Uf ; phiSigmaDotU = Sf & sigmaDotU
            {
                Vec UfN(3 * (size_t)s.nFaces), gU(9 * (size_t)nC), gUB(9 * (size_t)nB), gUf(9 * (size_t)s.nFaces), snU(3 * (size_t)nB);
                linearInterpolate(s, 3, s.U.data(), s.UB.data(), nullptr, UfN.data());
                patchSnGrad(s, 3, s.bcU, s.U.data(), s.UB.data(), nullptr, snU.data());
                gaussGradCells(s, 3, UfN.data(), gU.data());
                gaussGradBoundary(s, gU.data(), snU.data(), gUB.data());
                linearInterpolate(s, 9, gU.data(), gUB.data(), nullptr, gUf.data());
                forFaces(s, [&](int f) {
                    const double* Uf = &s.Uf[3 * (size_t)f];                  // Uf of updateFields.H (old U)
                    double sg[3];
                    for (int i = 0; i < 3; ++i) {
                        sg[i] = 0.0;
                        for (int j = 0; j < 3; ++j) sg[i] += (s.muf[f] * gUf[9 * (size_t)f + 3 * i + j] + s.tauMC[9 * (size_t)f + 3 * i + j]) * Uf[j];
                    }
                    s.phiSigmaDotU[f] = s.Sf[3 * (size_t)f] * sg[0] + s.Sf[3 * (size_t)f + 1] * sg[1] + s.Sf[3 * (size_t)f + 2] * sg[2];
                });
            }
        } else {
        // :79-86  solve(fvm::ddt(rho,U) - fvc::ddt(rhoU) == 0)
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c)
            for (int j = 0; j < 3; ++j) {
                const size_t i = 3 * (size_t)c + j;
                const double diag = rDeltaT * s.rho[c] * s.V[c];
                double source = rDeltaT * rho0[c] * U0[i] * s.V[c] + s.V[c] * (rDeltaT * (s.rhoU[i] - rhoU0[i]));
                if (!s.suU.empty()) source += s.suU[i];                                   // == rhoUSu  QGDUEqn.H:85 (rhoU itself is not touched)
                s.U[i] = source / diag;
            }
        correctU(s);
        }
        for (int b = 0; b < nB; ++b) for (int j = 0; j < 3; ++j) s.rhoUB[3 * (size_t)b + j] = s.rhoB[b] * s.UB[3 * (size_t)b + j];   // :88-89
        // ---- QGDEEqn.H:37-46
        fvcDiv(s, 1, s.phiJmH.data(), d1.data());
        fvcDiv(s, 1, s.phiQ.data(), d1b.data());
        fvcDiv(s, 1, s.phiPiU.data(), d1c.data());
        fvcDiv(s, 1, s.phiSigmaDotU.data(), d1d.data());
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) {
            const double diag = rDeltaT * s.V[c];
            double source = rDeltaT * rhoE0[c] * s.V[c];
            source -= s.V[c] * d1[c]; source -= s.V[c] * d1b[c]; source += s.V[c] * d1c[c]; source += s.V[c] * d1d[c];
            s.rhoE[c] = source / diag;
        }
        // :49-50
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) {
            const double* u = &s.U[3 * (size_t)c];
            s.e[c] = s.rhoE[c] / s.rho[c] - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        }
        correctE(s);
        if (s.prm.implicitDiffusion) {
            // :53-63  fvm::ddt(rho,e) - fvc::ddt(rho,e) - fvm::laplacian(alphauf,e) == 0 ; rhoE = rho*(e + 0.5*magSqr(U))
            Vec upper(std::max(s.nInternal, 1)), diag(nC), src(nC);
            for (int c = 0; c < nC; ++c) {
                diag[c] = rDeltaT * s.rho[c] * s.V[c];
                src[c] = rDeltaT * rho0[c] * e0[c] * s.V[c] + s.V[c] * (rDeltaT * (s.rho[c] * s.e[c] - rho0[c] * e0[c]));
                if (!s.suE.empty()) src[c] += s.suE[c];                                   // == rhoESu  QGDEEqn.H:60
            }
            for (int f = 0; f < s.nInternal; ++f) {
                const double up = s.ndC[f] * (s.alphauf[f] * s.magSf[f]);
                upper[f] = -up; diag[s.owner[f]] += up; diag[s.neighbour[f]] += up;
            }
            for (int b = 0; b < nB; ++b) {                                   // fixedEnergy: value BC ; gradientEnergy: gradient 0
                const int pi = s.bfacePatch[b];
                if (s.patchKind[pi] == OR_PATCH_EMPTY || s.bcT[pi] != OR_BC_FIXED_VALUE) continue;
                const int f = s.nInternal + b; const int P = s.owner[f];
                const double gS = s.alphauf[f] * s.magSf[f];
                diag[P] += gS * s.ndC[f];
                src[P] += gS * (s.ndC[f] * s.eB[b]);
            }
            s.lastDiffIters[3] = or_pcg_solve_blocks(sp, diag.data(), upper.data(), src.data(), s.e.data(), s.prm.diffTol, s.prm.diffRelTol,
                                              s.prm.diffMaxIter, s.prm.diffPrecond, nullptr, nullptr, s.pcgBlocks.empty() ? nullptr : s.pcgBlocks.data());
            correctE(s);
            for (int c = 0; c < nC; ++c) {
                const double* u = &s.U[3 * (size_t)c];
                s.rhoE[c] = s.rho[c] * (s.e[c] + 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]));
            }
        } else {
        // :65-73  solve(fvm::ddt(rho,e) - fvc::ddt(rhoE) == 0)
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) {
            const double diag = rDeltaT * s.rho[c] * s.V[c];
            const double ddt = s.prm.energyDdtRhoEQuirk ? (rDeltaT * (s.rhoE[c] - rhoE0[c]))
                                                        : (rDeltaT * (s.rho[c] * s.e[c] - rho0[c] * e0[c]));
            double source = rDeltaT * rho0[c] * e0[c] * s.V[c] + s.V[c] * ddt;
            if (!s.suE.empty()) source += s.suE[c];                                       // == rhoESu  QGDEEqn.H:71
            s.e[c] = source / diag;
        }
        correctE(s);
        }
        for (int b = 0; b < nB; ++b) {                                                    // :75-76
            const double* u = &s.UB[3 * (size_t)b];
            s.rhoEB[b] = s.rhoB[b] * (s.eB[b] + 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]));
        }
        // ---- QGDFoam.C:149-156
        thermoCorrect(s);
#pragma omp parallel for num_threads(s.nThreads) schedule(static)
        for (int c = 0; c < nC; ++c) s.p[c] = s.rho[c] / s.psi[c];
        correctP(s);
        for (int b = 0; b < nB; ++b) if (!patchIsEmpty(s, b)) s.rhoB[b] = s.psiB[b] * s.pB[b];
    }
    return CoNum;
}

void or_qgd_get(or_ctx* s, int field, double* cells, double* bnd)
{
    const Vec* ci = nullptr; const Vec* bi = nullptr;
    switch (field) {
        case 0: ci = &s->rho; bi = &s->rhoB; break;
        case 1: ci = &s->rhoU; bi = &s->rhoUB; break;
        case 2: ci = &s->rhoE; bi = &s->rhoEB; break;
        case 3: ci = &s->U; bi = &s->UB; break;
        case 4: ci = &s->e; bi = &s->eB; break;
        case 5: ci = &s->p; bi = &s->pB; break;
        case 6: ci = &s->T; bi = &s->TB; break;
        case 7: ci = &s->c; bi = &s->cB; break;
        case 8: ci = &s->mu; bi = &s->muB; break;
        case 9: ci = &s->alpha; bi = &s->alphaB; break;
        case 10: ci = &s->tauQGD; bi = &s->tauQGDB; break;
        case 12: ci = &s->ScQGD; bi = &s->ScQGDB; break;
        default: return;
    }
    if (cells) std::copy(ci->begin(), ci->end(), cells);
    if (bnd) std::copy(bi->begin(), bi->end(), bnd);
}

void or_qgd_get_face(or_ctx* s, int field, double* out)
{
    const Vec* v = nullptr;
    switch (field) {
        case 0: v = &s->phiJm; break;  case 1: v = &s->phiJmU; break; case 2: v = &s->phiP; break;
        case 3: v = &s->phiPi; break;  case 4: v = &s->phiJmH; break; case 5: v = &s->phiQ; break;
        case 6: v = &s->phiPiU; break; case 7: v = &s->tauQGDf; break; case 8: v = &s->gradUf; break;
        case 9: v = &s->gradef; break; case 10: v = &s->gradRhof; break; case 11: v = &s->gradPf; break;
        case 12: v = &s->phiw; break;
        case 13: v = &s->phiTauMC; break; case 14: v = &s->phiSigmaDotU; break;
        default: return;
    }
    std::copy(v->begin(), v->end(), out);
}

// [OF-v2312] lduMatrix PCG with DIC / diagonal / no preconditioner; symmetric matrix
int or_pcg_solve_blocks(or_ctx* sp, const double* diag, const double* upper, const double* b, double* x, double tol,
                        double relTol, int maxIter, int precond, double* initRes, double* finalRes, const int* cellBlock);

int or_pcg_solve(or_ctx* sp, const double* diag, const double* upper, const double* b, double* x, double tol,
                 double relTol, int maxIter, int precond, double* initRes, double* finalRes)
{
    return or_pcg_solve_blocks(sp, diag, upper, b, x, tol, relTol, maxIter, precond, initRes, finalRes, nullptr);
}

// The same solver as `mpirun -np N` runs it [OF-v2312 PCG on a decomposed lduMatrix]: the matrix, the dot products
// (gSumProd) and the residual norms (gSumMag, normFactor with the global average) are those of the whole mesh; only the
// preconditioner is block-local - DIC is factorised and swept on each processor's own block, faces between cells of
// different blocks (processor interfaces) are left out of it (DICPreconditioner works on the local lduMatrix).
// cellBlock: processor of each cell, or NULL = one block (the serial solver).
int or_pcg_solve_blocks(or_ctx* sp, const double* diag, const double* upper, const double* b, double* x, double tol,
                        double relTol, int maxIter, int precond, double* initRes, double* finalRes, const int* cellBlock)
{
    const or_ctx& s = *sp;
    const int n = s.nCells, nf = s.nInternal;
    const int* l = s.owner.data(); const int* u = s.neighbour.data();
    auto Amul = [&](Vec& y, const double* v) {
        for (int c = 0; c < n; ++c) y[c] = diag[c] * v[c];
        for (int f = 0; f < nf; ++f) { y[u[f]] += upper[f] * v[l[f]]; y[l[f]] += upper[f] * v[u[f]]; }
    };
    auto inBlock = [&](int f) { return !cellBlock || cellBlock[l[f]] == cellBlock[u[f]]; };
    Vec wA(n), pA(n, 0.0), rA(n), rD;
    Amul(wA, x);
    for (int c = 0; c < n; ++c) rA[c] = b[c] - wA[c];
    // normFactor
    double xRef = 0.0; for (int c = 0; c < n; ++c) xRef += x[c]; xRef /= n;
    Vec sumA(diag, diag + n);
    for (int f = 0; f < nf; ++f) { sumA[u[f]] += upper[f]; sumA[l[f]] += upper[f]; }
    double normFactor = 0.0;
    for (int c = 0; c < n; ++c) { const double t = sumA[c] * xRef; normFactor += std::fabs(wA[c] - t) + std::fabs(b[c] - t); }
    normFactor += 1e-20;
    auto sumMag = [&](const Vec& v) { double r = 0; for (int c = 0; c < n; ++c) r += std::fabs(v[c]); return r; };
    double res0 = sumMag(rA) / normFactor, res = res0;
    if (initRes) *initRes = res0;
    int it = 0;
    auto converged = [&]() { return res < tol || (relTol > 1e-20 && res < relTol * res0); };
    if (!converged()) {
        if (precond == 2) {
            rD.assign(diag, diag + n);
            for (int f = 0; f < nf; ++f) if (inBlock(f)) rD[u[f]] -= upper[f] * upper[f] / rD[l[f]];
            for (int c = 0; c < n; ++c) rD[c] = 1.0 / rD[c];
        } else if (precond == 1) { rD.resize(n); for (int c = 0; c < n; ++c) rD[c] = 1.0 / diag[c]; }
        double wArA = 1e20, wArAold = wArA;
        do {
            wArAold = wArA;
            if (precond == 0) wA = rA;
            else {
                for (int c = 0; c < n; ++c) wA[c] = rD[c] * rA[c];
                if (precond == 2) {
                    for (int f = 0; f < nf; ++f) if (inBlock(f)) wA[u[f]] -= rD[u[f]] * upper[f] * wA[l[f]];
                    for (int f = nf - 1; f >= 0; --f) if (inBlock(f)) wA[l[f]] -= rD[l[f]] * upper[f] * wA[u[f]];
                }
            }
            wArA = 0; for (int c = 0; c < n; ++c) wArA += wA[c] * rA[c];
            if (it == 0) pA = wA;
            else { const double beta = wArA / wArAold; for (int c = 0; c < n; ++c) pA[c] = wA[c] + beta * pA[c]; }
            Amul(wA, pA.data());
            double wApA = 0; for (int c = 0; c < n; ++c) wApA += wA[c] * pA[c];
            if (std::fabs(wApA) / normFactor < 1e-300) break;
            const double alpha = wArA / wApA;
            for (int c = 0; c < n; ++c) { x[c] += alpha * pA[c]; rA[c] -= alpha * wA[c]; }
            res = sumMag(rA) / normFactor;
        } while (++it < maxIter && !converged());
    }
    if (finalRes) *finalRes = res;
    return it;
}

} // extern "C"

#include "qhd_oracle.inc"
