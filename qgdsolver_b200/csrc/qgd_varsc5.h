// varScModel5 (varScModel5.C:52-269) on the device: the density-gradient ScQGD sensor, relaxed against the previous ScQGD, clamped,
// floored by the mesh-quality value, smoothed with fvc::smooth, then muQGD / alphauQGD.
//
// Why this model has its own pass.  The sensor reads rho = psi*p with the psi of the NEW temperature and the OLD pressure, in the
// cell and in its face neighbours (fvc::grad), so it cannot be fused into the cell update like varScModel6/7: the step runs the
// ordinary kernels (their Consts carry model 0 / tauMode 2, i.e. tauQGDf = I(alphaQGD) hQGDf / I(c), varScModel5.C:204-205), and
// this pass then rewrites the three state fields that depend on ScQGD: mu, alphaEff and the tau slot (cells and boundary faces).
// None of the existing kernels changes.
//
// fvc::smooth is OpenFOAM's FaceCellWave<smoothData>: a sequential, list-driven sweep whose result depends on the visiting order
// inside its 1 % propagation tolerance.  Order-exact parallel form used here (the sequential restatement
// it is tested against lives with the CPU checker, not in this package):
//   * faceToCell reads face values only and writes cell values only, so cells are independent; a cell applies the offers of its
//     changed faces in the order those faces hold in the changed-face list (their list position is kept per face);
//   * cellToFace likewise: a face applies the offers of its (at most two) changed cells in changed-cell-list order;
//   * an item joins the next list at its first successful update; its list position is the rank of the key
//     (position of the offering item, owner-before-neighbour | index of the face in mesh.cells()[cell]) - an ordered compaction
//     (flag array + prefix sum) per half-sweep.  Boundary faces are left out: a boundary face only ever carries a value its
//     owner has held, which can never raise the owner again, so it influences neither values nor the order of the others.
//
// Every per-item body below is a plain functor (QGD_HD) and every sequence a template over an executor, so that the same code
// runs under the CUDA executor (qgd_varsc5.cu: one kernel launch per forEach) and, compiled with g++, under a serial host
// executor in the CPU test suite (tests/test_varsc5_host_cpu.py) against the oracle.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

#include "qgd_kernels.cuh"

#ifndef QGD_HD
#ifdef __CUDACC__
#define QGD_HD __host__ __device__ __forceinline__
#else
#define QGD_HD inline
#endif
#endif

namespace qgd {

// host-side, time-constant data of the model, polyMesh numbering (HostMesh -> buildVarSc5Host in qgd_host_setup.cpp)
struct VarSc5Host {
    int nC = 0, nI = 0, nB = 0, maxCellFaces = 0;
    std::vector<int> own, nei;           // nF, nI
    std::vector<double> w;               // nI   linear weights
    std::vector<double> Sf;              // 3*nF (AoS)
    std::vector<double> bMagSf, bDC, bHf;   // nB
    std::vector<int> bKind;              // nB   patch kind
    std::vector<int> ccOff, ccFace;      // mesh.cells() [OF-v2312 primitiveMesh::calcCells]: owned faces ascending, then neighbour-side faces
    std::vector<int> lidxOwn, lidxNei;   // nI   index of the face inside mesh.cells()[owner] / [neighbour]
    std::vector<double> cqSc;            // nC   varScModel5.C:112-132
    std::vector<double> aspectRatio;     // nC   primitiveMeshTools::cellClosedness [OF-v2312]
};
void buildVarSc5Host(const HostMesh& h, double badQualitySc, double maxAspectRatio, VarSc5Host& out);

struct VarSc5View {
    int nC, nI, nB, maxCellFaces;
    // mesh, polyMesh numbering
    const int* own; const int* nei; const double* w; const double* Sf;
    const double* bMagSf; const double* bDC; const double* bHf; const int* bKind;
    const int* ccOff; const int* ccFace; const int* lidxOwn; const int* lidxNei;
    const double* V; const double* hQGD; const double* cqSc;
    const unsigned char* scConst;        // constScCellSet mask or null
    // dictionary
    double rC, minSc, maxSc, ScDict, maxRatio;
    Consts k;                            // thermo / transport constants of the solver
    // solver state
    double* S;                           // [16][nC] cell state (SoA), see SolverView
    RecA* bA; RecB* bB; const double* psiB; const double* aQGD;
    const double* pOld;                  // nC  pressure before the cell update
    const double* pOldB;                 // nB  p_b as left by the last p.correctBoundaryConditions() before the closing one
    double* Sc; double* ScB;             // ScQGD cells / boundary faces (calculated patches, QGDCoeffs.C:249-261)
    // scratch
    double* rho; double* rhoB; double* g;   // nC, nB, [3][nC]
    double* faceV;                       // nI  smoothData of the internal faces
    int* posF; int* posC;                // nI, nC : position in the current changed list or -1
    int* listF; int* listC;              // changed lists (nI, nC)
    int* mark;                           // flag array of the ordered compaction: max(2 nI, nC maxCellFaces, nI) ints
    int* cnt;                            // per-thread counts / offsets of the compaction (kCompactThreads + 1)
    int* nOut;                           // 1 int: size of the list just built
};
constexpr int kV5CompactThreads = 8192;

// ---- thermo pieces the closing pass needs (same expressions as muMol / alphahMol of qgd_kernels.cu)
QGD_HD double v5MuMol(const Consts& k, double T)
{
    if (k.transport == 1) return k.mu0 * pow(T / k.T0, k.kExp);          // powerLawTransportI.H:121-128
    if (k.transport == 2) return k.As * sqrt(T) / (1.0 + k.Ts / T);      // sutherlandTransport::mu [OF-v2312]
    return k.mu;
}
QGD_HD double v5AlphahMol(const Consts& k, double muT)
{
    if (k.transport == 1) return muT * k.rPr;
    if (k.transport == 2) return muT * k.Cv * (1.32 + 1.77 * k.R / k.Cv) / k.Cp;
    return k.mu / k.Pr;
}

// ---- varScModel5.C:202 : rho = qgdThermo.rho() = psi p [OF-v2312 psiThermo::rho], new psi, old p
struct V5RhoCells {
    VarSc5View v;
    QGD_HD void operator()(int c) const { v.rho[c] = v.pOld[c] * (1.0 / (v.k.R * v.S[6 * (size_t)v.nC + c])); }
};
struct V5RhoBnd {
    VarSc5View v;
    QGD_HD void operator()(int b) const { v.rhoB[b] = v.bKind[b] == QGD_PATCH_EMPTY ? 0.0 : v.pOldB[b] * v.psiB[b]; }
};
// [OF-v2312] fvc::grad(rho), Gauss linear: (1/V) sum_f +-Sf rho_f over mesh.cells()[c]
struct V5GradCells {
    VarSc5View v;
    QGD_HD void operator()(int c) const
    {
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        const double rc = v.rho[c];
        for (int q = v.ccOff[c]; q < v.ccOff[c + 1]; ++q) {
            const int f = v.ccFace[q];
            double rf, sgn = 1.0;
            if (f < v.nI) {
                const int P = v.own[f], N = v.nei[f];
                const double rP = (P == c) ? rc : v.rho[P], rN = (N == c) ? rc : v.rho[N];
                rf = v.w[f] * (rP - rN) + rN;
                if (N == c) sgn = -1.0;
            } else {
                const int b = f - v.nI;
                if (v.bKind[b] == QGD_PATCH_EMPTY) continue;
                rf = v.rhoB[b];
            }
            g0 += sgn * (v.Sf[3 * (size_t)f] * rf); g1 += sgn * (v.Sf[3 * (size_t)f + 1] * rf); g2 += sgn * (v.Sf[3 * (size_t)f + 2] * rf);
        }
        const double V = v.V[c];
        v.g[c] = g0 / V; v.g[(size_t)v.nC + c] = g1 / V; v.g[2 * (size_t)v.nC + c] = g2 / V;
    }
};
// varScModel5.C:209-230 on the cells: relaxation, clamps, quality floor, constScCellSet
struct V5ScCells {
    VarSc5View v;
    QGD_HD void operator()(int c) const
    {
        const double gx = v.g[c], gy = v.g[(size_t)v.nC + c], gz = v.g[2 * (size_t)v.nC + c];
        const double mg = sqrt(gx * gx + gy * gy + gz * gz);
        double sc = v.rC * (mg * v.hQGD[c] / v.rho[c]) + (1.0 - v.rC) * v.Sc[c];
        sc = fmin(fmax(sc, v.minSc), v.maxSc);
        sc = fmax(sc, v.cqSc[c]);
        if (v.scConst && v.scConst[c]) sc = v.ScDict;
        v.Sc[c] = sc;
    }
};
// ... and on the boundary faces: fvc::grad boundary values [OF-v2312 gaussGrad::correctBoundaryConditions] g_P + n (snGrad_b - n.g_P),
// snGrad_b = deltaCoeffs (rho_b - rho_P) (calculated patch), hQGD_b = hQGDf_b (QGDCoeffs.C:373)
struct V5ScBnd {
    VarSc5View v;
    QGD_HD void operator()(int b) const
    {
        if (v.bKind[b] == QGD_PATCH_EMPTY) return;
        const int f = v.nI + b, P = v.own[f];
        const double ms = v.bMagSf[b];
        const double n[3] = {v.Sf[3 * (size_t)f] / ms, v.Sf[3 * (size_t)f + 1] / ms, v.Sf[3 * (size_t)f + 2] / ms};
        const double gP[3] = {v.g[P], v.g[(size_t)v.nC + P], v.g[2 * (size_t)v.nC + P]};
        const double sn = v.bDC[b] * (v.rhoB[b] - v.rho[P]);
        const double nG = n[0] * gP[0] + n[1] * gP[1] + n[2] * gP[2];
        const double gB[3] = {gP[0] + n[0] * (sn - nG), gP[1] + n[1] * (sn - nG), gP[2] + n[2] * (sn - nG)};
        const double mg = sqrt(gB[0] * gB[0] + gB[1] * gB[1] + gB[2] * gB[2]);
        const double sc = v.rC * (mg * v.bHf[b] / v.rhoB[b]) + (1.0 - v.rC) * v.ScB[b];
        v.ScB[b] = fmin(fmax(sc, v.minSc), v.maxSc);
    }
};
// varScModel5.C:207,244-268 + QGDThermo.C:91-98: the state fields that depend on ScQGD
struct V5CloseCells {
    VarSc5View v;
    QGD_HD void operator()(int c) const
    {
        const size_t n = v.nC;
        const double T = v.S[6 * n + c], cs = v.S[12 * n + c], aQ = v.aQGD[c];
        const double tau = aQ * v.hQGD[c] / cs;
        const double muQGD = v.pOld[c] * v.Sc[c] * tau;
        const double muT = v5MuMol(v.k, T);
        const double alpha = v5AlphahMol(v.k, muT) + muQGD / v.k.PrQGD;
        v.S[13 * n + c] = muT + muQGD;
        v.S[14 * n + c] = v.k.alphaEffGamma ? v.k.gamma * alpha : alpha;
        v.S[15 * n + c] = aQ;                          // tau slot of tauMode 2: I(alphaQGD) hQGDf / I(c)
    }
};
struct V5CloseBnd {
    VarSc5View v;
    QGD_HD void operator()(int b) const
    {
        if (v.bKind[b] == QGD_PATCH_EMPTY) return;
        const int P = v.own[v.nI + b];
        const double T = v.bA[b].T, aQ = v.aQGD[P];
        RecB bb = v.bB[b];
        const double tauB = aQ * v.bHf[b] / bb.c;
        const double muQGD = v.pOldB[b] * v.ScB[b] * tauB;
        const double muT = v5MuMol(v.k, T);
        const double alpha = v5AlphahMol(v.k, muT) + muQGD / v.k.PrQGD;
        bb.mu = muT + muQGD;
        bb.alphaEff = v.k.alphaEffGamma ? v.k.gamma * alpha : alpha;
        bb.aByC = aQ;
        v.bB[b] = bb;
    }
};

// ---- [OF-v2312] smoothData::update (smoothDataI.H): valid = value > -SMALL, VSMALL = 1e-300
QGD_HD bool v5Update(double& mine, double other, double scale)
{
    const double tol = 0.01;                           // FaceCellWave::propagationTol_
    if (!(mine > -1.0e-15) || mine < 1.0e-300) { mine = other; return true; }
    if (other > (1 + tol) * scale * mine) { mine = other / scale; return true; }
    return false;
}
// smooth.C: initial changed faces (internal faces, ascending polyMesh id)
struct V5InitFaces {
    VarSc5View v;
    QGD_HD void operator()(int f) const
    {
        const double a = v.Sc[v.own[f]], b = v.Sc[v.nei[f]];
        double val = -1.0e15;                          // smoothData(): value_(-GREAT)
        int m = 0;
        if (a > v.maxRatio * b) { val = a; m = f + 1; }
        else if (b > v.maxRatio * a) { val = b; m = f + 1; }
        v.faceV[f] = val;
        v.mark[f] = m;
        v.posF[f] = -1;
    }
};
// FaceCellWave::faceToCell for one cell: offers of its changed faces in list order
struct V5FaceToCell {
    VarSc5View v;
    QGD_HD void operator()(int c) const
    {
        double val = v.Sc[c];
        int last = -1, firstKey = -1;
        while (true) {
            int best = 0x7fffffff, bf = -1;
            for (int q = v.ccOff[c]; q < v.ccOff[c + 1]; ++q) {
                const int f = v.ccFace[q];
                if (f >= v.nI) continue;
                const int p = v.posF[f];
                if (p > last && p < best) { best = p; bf = f; }
            }
            if (bf < 0) break;
            last = best;
            const double fv = v.faceV[bf];
            if (val == fv) continue;                   // currInfo.equal(newInfo)
            if (v5Update(val, fv, v.maxRatio) && firstKey < 0) firstKey = 2 * best + (v.own[bf] == c ? 0 : 1);
        }
        if (firstKey >= 0) { v.Sc[c] = val; v.mark[firstKey] = c + 1; }
    }
};
// FaceCellWave::cellToFace for one internal face: offers of its changed cells in list order
struct V5CellToFace {
    VarSc5View v;
    QGD_HD void operator()(int f) const
    {
        const int P = v.own[f], N = v.nei[f];
        const int pP = v.posC[P], pN = v.posC[N];
        if (pP < 0 && pN < 0) return;
        double val = v.faceV[f];
        long long firstKey = -1;
        for (int pass = 0; pass < 2; ++pass) {
            // first the cell that comes earlier in the changed-cell list
            const bool ownerFirst = (pN < 0) || (pP >= 0 && pP < pN);
            const bool useOwner = (pass == 0) == ownerFirst;
            const int c = useOwner ? P : N, pc = useOwner ? pP : pN;
            if (pc < 0) continue;
            const double cv = v.Sc[c];
            if (val == cv) continue;
            if (v5Update(val, cv, 1.0) && firstKey < 0)
                firstKey = (long long)pc * v.maxCellFaces + (useOwner ? v.lidxOwn[f] : v.lidxNei[f]);
        }
        if (firstKey >= 0) { v.faceV[f] = val; v.mark[firstKey] = f + 1; }
    }
};
struct V5ResetPos { int* pos; const int* list; QGD_HD void operator()(int i) const { pos[list[i]] = -1; } };
// ordered compaction of the flag array mark[0, M): thread t counts / fills its chunk; one thread scans the counts
struct V5CompactCount {
    const int* mark; int* cnt; int M, L;
    QGD_HD void operator()(int t) const
    {
        const long long lo = (long long)t * L;
        const long long hi = lo + L < M ? lo + L : M;
        int n = 0;
        for (long long i = lo; i < hi; ++i) n += mark[i] != 0;
        cnt[t] = n;
    }
};
struct V5CompactScan {
    int* cnt; int T; int* nOut;
    QGD_HD void operator()(int) const
    {
        int run = 0;
        for (int t = 0; t < T; ++t) { const int n = cnt[t]; cnt[t] = run; run += n; }
        *nOut = run;
    }
};
struct V5CompactFill {
    const int* mark; const int* cnt; int M, L; int* list; int* pos;
    QGD_HD void operator()(int t) const
    {
        const long long lo = (long long)t * L;
        const long long hi = lo + L < M ? lo + L : M;
        int o = cnt[t];
        for (long long i = lo; i < hi; ++i) {
            const int m = mark[i];
            if (m) { list[o] = m - 1; pos[m - 1] = o; ++o; }
        }
    }
};

// list = the ids flagged in mark[0, M) in index order, pos[id] = rank; returns the count (one host read)
template <class Exec> int v5Compact(Exec& ex, const VarSc5View& v, long long M, int* list, int* pos)
{
    if (M <= 0) return 0;
    if (M > 0x7fffffffLL) throw Error(QGD_ERR_UNSUPPORTED, "varScModel5: mesh too large for the 32-bit keys of the ordered compaction");
    int T = (int)((M + 63) / 64);
    if (T > kV5CompactThreads) T = kV5CompactThreads;
    const int L = (int)((M + T - 1) / T);
    ex.forEach(T, V5CompactCount{v.mark, v.cnt, (int)M, L});
    ex.forEach(1, V5CompactScan{v.cnt, T, v.nOut});
    ex.forEach(T, V5CompactFill{v.mark, v.cnt, (int)M, L, list, pos});
    return ex.readInt(v.nOut);
}

// fvc::smooth(ScQGD, smoothCoeff) (varScModel5.C:232); returns the number of completed FaceCellWave iterations
template <class Exec> int v5Smooth(Exec& ex, const VarSc5View& v)
{
    if (v.nI == 0) return 0;
    ex.fillInt(v.posC, -1, (size_t)v.nC);
    ex.forEach(v.nI, V5InitFaces{v});
    int nF = v5Compact(ex, v, v.nI, v.listF, v.posF);
    int iter = 0;
    while (nF > 0) {
        ex.fillInt(v.mark, 0, 2 * (size_t)nF);
        ex.forEach(v.nC, V5FaceToCell{v});
        ex.forEach(nF, V5ResetPos{v.posF, v.listF});
        const int nCh = v5Compact(ex, v, 2LL * nF, v.listC, v.posC);
        if (nCh == 0) break;
        ex.fillInt(v.mark, 0, (size_t)nCh * v.maxCellFaces);
        ex.forEach(v.nI, V5CellToFace{v});
        ex.forEach(nCh, V5ResetPos{v.posC, v.listC});
        nF = v5Compact(ex, v, (long long)nCh * v.maxCellFaces, v.listF, v.posF);
        if (nF == 0) break;
        ++iter;
    }
    return iter;
}

// varScModel5::correct (varScModel5.C:197-268) on top of a state closed by the ordinary kernels
template <class Exec> int v5Correct(Exec& ex, const VarSc5View& v)
{
    ex.forEach(v.nC, V5RhoCells{v});
    ex.forEach(v.nB, V5RhoBnd{v});
    ex.forEach(v.nC, V5GradCells{v});
    ex.forEach(v.nB, V5ScBnd{v});          // before the cells: both read the old ScQGD of their own entry only
    ex.forEach(v.nC, V5ScCells{v});
    const int iters = v5Smooth(ex, v);
    ex.forEach(v.nC, V5CloseCells{v});
    ex.forEach(v.nB, V5CloseBnd{v});
    return iters;
}

} // namespace qgd
