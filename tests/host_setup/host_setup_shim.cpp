// Test shim (CPU): exposes the product's host-side set-up code (qgdsolver_b200/csrc/qgd_host_setup.cpp: HostMesh::build,
// buildFaceRecords, buildLeastSquares) through a small C interface so that it can be checked against the oracle without a GPU.
// Compiled by tests/test_host_setup_cpu.py with g++ from the product sources where they lie; nothing here is shipped.
#include <cstring>

#include "qgd_internal.h"

namespace qgd { void setLastError(const std::string&) {} }

using qgd::HostMesh;

extern "C" {

void* hs_create(const qgd_mesh_desc* d)
{
    try { HostMesh* h = new HostMesh(); h->build(*d); return h; } catch (...) { return nullptr; }
}
void hs_destroy(void* p) { delete static_cast<HostMesh*>(p); }

void hs_lengths(void* p, double* hQGDf, double* hQGD)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    std::memcpy(hQGDf, h.hQGDf.data(), sizeof(double) * h.nFaces);
    std::memcpy(hQGD, h.hQGD.data(), sizeof(double) * h.nCells);
}

// vtx nFaces*4, flags nFaces, G 9*nFaces (SoA), halfDist nBnd
void hs_face_records(void* p, int reduced, int* vtx, int* flags, double* G, double* halfDist)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    std::vector<int> v, f;
    std::vector<double> g, hd;
    h.buildFaceRecords(reduced != 0, v, f, g, hd);
    std::memcpy(vtx, v.data(), sizeof(int) * v.size());
    std::memcpy(flags, f.data(), sizeof(int) * f.size());
    std::memcpy(G, g.data(), sizeof(double) * g.size());
    if (!hd.empty()) std::memcpy(halfDist, hd.data(), sizeof(double) * hd.size());
}

// point interpolation of one scalar field with the product's weights: interior points from cells, patch points from boundary values
void hs_point_interpolate(void* p, const double* cell, const double* bnd, double* out)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    for (int i = 0; i < h.nPoints; ++i) {
        double a = 0.0;
        for (int q = h.pcOff[i]; q < h.pcOff[i + 1]; ++q) a += h.pcW[q] * cell[h.pcCell[q]];
        out[i] = a;
    }
    for (size_t k = 0; k < h.patchPoints.size(); ++k) {
        double a = 0.0;
        for (int q = h.ppOff[k]; q < h.ppOff[k + 1]; ++q) a += h.ppW[q] * bnd[h.ppFace[q]];
        out[h.patchPoints[k]] = a;
    }
}

int hs_lsq_width(void* p, int opt)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    int W = 0; std::vector<int> c; std::vector<double> k; std::vector<char> d;
    h.buildLeastSquares(opt != 0, W, c, k, d);
    return W;
}
// cells W*nI, coef W*3*nI, deg nI  (column-major like the device arrays)
void hs_lsq(void* p, int opt, int* cells, double* coef, char* deg)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    int W = 0; std::vector<int> c; std::vector<double> k; std::vector<char> d;
    h.buildLeastSquares(opt != 0, W, c, k, d);
    std::memcpy(cells, c.data(), sizeof(int) * c.size());
    std::memcpy(coef, k.data(), sizeof(double) * k.size());
    std::memcpy(deg, d.data(), d.size());
}

// DIC blocks by recursive coordinate bisection (HostMesh::makePcgBlocks): block id per cell, returns the number of blocks
int hs_pcg_blocks(void* p, int target, int* out)
{
    HostMesh& h = *static_cast<HostMesh*>(p);
    h.makePcgBlocks(target);
    int nb = 0;
    for (int c = 0; c < h.nCells; ++c) { out[c] = h.pcgBlock[c]; nb = nb > out[c] + 1 ? nb : out[c] + 1; }
    return nb;
}

// cell -> faces rows (CSR) as the device ELL / tail builder consumes them
int hs_cell_faces(void* p, int* off, int* enc)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    if (off) std::memcpy(off, h.cfOff.data(), sizeof(int) * h.cfOff.size());
    if (enc) std::memcpy(enc, h.cfEnc.data(), sizeof(int) * h.cfEnc.size());
    return (int)h.cfEnc.size();
}

}

// ---------------------------------------------------------------- varScModel5: the product's device functors and sequences
// (qgdsolver_b200/csrc/qgd_varsc5.h) under a serial host executor - the same code the CUDA executor launches, one "kernel" at a time
#include <cstdlib>

#include "qgd_varsc5.h"

namespace {
struct HostExec {
    long long launches = 0;
    int order = 0;          // 0 ascending, 1 descending, 2 pseudo-random thread order: results must not depend on it
    template <class F> void forEach(int n, const F& f)
    {
        ++launches;
        if (order == 0) for (int i = 0; i < n; ++i) f(i);
        else if (order == 1) for (int i = n - 1; i >= 0; --i) f(i);
        else {
            // a full-period stride permutation: i -> (a i + b) mod n with gcd(a, n) = 1
            long long a = 7919 % (n > 0 ? n : 1);
            if (n > 1) { auto g = [](long long x, long long y) { while (y) { long long t = x % y; x = y; y = t; } return x; }; while (a < 2 || g(a, n) != 1) ++a; }
            for (int i = 0; i < n; ++i) f((int)(((long long)a * i + 3) % n));
        }
    }
    void fillInt(int* p, int value, size_t n) { for (size_t i = 0; i < n; ++i) p[i] = value ? -1 : 0; }
    int readInt(const int* p) { return *p; }
};

struct V5HostState {
    qgd::VarSc5Host hm;
    std::vector<double> pOld, pOldB, ScB, rho, rhoB, g, faceV;
    std::vector<int> posF, posC, listF, listC, mark, cnt, nOut;
    qgd::VarSc5View view(const qgd::HostMesh& h)
    {
        const int nC = hm.nC, nI = hm.nI, nB = hm.nB;
        pOld.assign(nC, 0.0); pOldB.assign(nB + 1, 0.0); ScB.assign(nB + 1, 0.0); rho.assign(nC, 0.0); rhoB.assign(nB + 1, 0.0);
        g.assign(3 * (size_t)nC, 0.0); faceV.assign(nI + 1, 0.0); posF.assign(nI + 1, -7); posC.assign(nC, -7); listF.assign(nI + 1, 0);
        listC.assign(nC, 0); mark.assign(std::max<size_t>(2 * (size_t)nI, (size_t)nC * hm.maxCellFaces) + 1, 12345);   // garbage: the code must initialise what it reads
        cnt.assign(qgd::kV5CompactThreads + 1, 0); nOut.assign(1, 0);
        qgd::VarSc5View v{};
        v.nC = nC; v.nI = nI; v.nB = nB; v.maxCellFaces = hm.maxCellFaces;
        v.own = hm.own.data(); v.nei = hm.nei.data(); v.w = hm.w.data(); v.Sf = hm.Sf.data(); v.bMagSf = hm.bMagSf.data(); v.bDC = hm.bDC.data();
        v.bHf = hm.bHf.data(); v.bKind = hm.bKind.data(); v.ccOff = hm.ccOff.data(); v.ccFace = hm.ccFace.data();
        v.lidxOwn = hm.lidxOwn.data(); v.lidxNei = hm.lidxNei.data(); v.V = h.V.data(); v.hQGD = h.hQGD.data(); v.cqSc = hm.cqSc.data();
        v.pOld = pOld.data(); v.pOldB = pOldB.data(); v.ScB = ScB.data(); v.rho = rho.data(); v.rhoB = rhoB.data(); v.g = g.data();
        v.faceV = faceV.data(); v.posF = posF.data(); v.posC = posC.data(); v.listF = listF.data(); v.listC = listC.data();
        v.mark = mark.data(); v.cnt = cnt.data(); v.nOut = nOut.data();
        return v;
    }
};
} // namespace

extern "C" {

// cqSc and aspect ratio of buildVarSc5Host; returns maxCellFaces
int hs_varsc5_quality(void* p, double badQualitySc, double maxAspectRatio, double* cqSc, double* aspectRatio)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    qgd::VarSc5Host hm;
    qgd::buildVarSc5Host(h, badQualitySc, maxAspectRatio, hm);
    std::memcpy(cqSc, hm.cqSc.data(), sizeof(double) * h.nCells);
    std::memcpy(aspectRatio, hm.aspectRatio.data(), sizeof(double) * h.nCells);
    return hm.maxCellFaces;
}

// fvc::smooth(field, coeff) through v5Smooth; returns the FaceCellWave iterations, *launches = kernel launches a device run would issue
int hs_varsc5_smooth(void* p, double* field, double coeff, int order, long long* launches)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    V5HostState st;
    qgd::buildVarSc5Host(h, 0.05, 1.5, st.hm);
    qgd::VarSc5View v = st.view(h);
    v.Sc = field; v.maxRatio = 1.0 + coeff;
    HostExec ex; ex.order = order;
    const int it = qgd::v5Smooth(ex, v);
    if (launches) *launches = ex.launches;
    return it;
}

// varScModel5::correct through v5Correct on a state closed by the ordinary kernels.  prm: R, Cp, mu, Pr, PrQGD, alphaEffGamma (0|1),
// rC, minSc, maxSc, ScDict, smoothCoeff, badQualitySc, maxAspectRatio.  S: [16][nC] in/out (fields 6 = T and 12 = c are read; 13, 14, 15
// written); bT, bC: boundary T and c in; bMu, bAlphaEff, bSlot: out; Sc / ScB in/out; constMask nC bytes or NULL.
int hs_varsc5_correct(void* p, const double* prm, double* S, const double* bT, const double* bC, const double* psiB, const double* aQGD,
                      const double* pOld, const double* pOldB, double* Sc, double* ScB, const unsigned char* constMask,
                      double* bMu, double* bAlphaEff, double* bSlot, int order)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    V5HostState st;
    qgd::buildVarSc5Host(h, prm[11], prm[12], st.hm);
    qgd::VarSc5View v = st.view(h);
    qgd::Consts k{};
    k.R = prm[0]; k.Cp = prm[1]; k.Cv = prm[1] - prm[0]; k.mu = prm[2]; k.Pr = prm[3]; k.PrQGD = prm[4]; k.gamma = k.Cp / k.Cv;
    k.alphaEffGamma = (int)prm[5]; k.transport = 0;
    v.k = k; v.rC = prm[6]; v.minSc = prm[7]; v.maxSc = prm[8]; v.ScDict = prm[9]; v.maxRatio = 1.0 + prm[10];
    const int nB = h.nBnd;
    std::vector<qgd::RecA> bA(nB + 1);
    std::vector<qgd::RecB> bB(nB + 1);
    for (int b = 0; b < nB; ++b) { bA[b] = qgd::RecA{0, 0, 0, 0, 0, 0, bT[b], 0}; bB[b] = qgd::RecB{0, 0, 0, 0, bC[b], -1.0, -1.0, -1.0}; }
    std::copy(pOld, pOld + h.nCells, st.pOld.begin()); std::copy(pOldB, pOldB + nB, st.pOldB.begin()); std::copy(ScB, ScB + nB, st.ScB.begin());
    v.S = S; v.bA = bA.data(); v.bB = bB.data(); v.psiB = psiB; v.aQGD = aQGD; v.Sc = Sc; v.scConst = constMask;
    HostExec ex; ex.order = order;
    const int it = qgd::v5Correct(ex, v);
    std::copy(st.ScB.begin(), st.ScB.begin() + nB, ScB);
    for (int b = 0; b < nB; ++b) { bMu[b] = bB[b].mu; bAlphaEff[b] = bB[b].alphaEff; bSlot[b] = bB[b].aByC; }
    return it;
}

} // extern "C"

// ---------------------------------------------------------------- wedge patches: the product's vertex list / normals and face tensor
#include "qgd_wedge.h"

extern "C" {

int hs_wedge_points(void* p, int* pts, double* R9)      // returns the count; pts / R9 (9 per vertex) may be NULL to query it
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    if (pts) std::memcpy(pts, h.wedgePts.data(), sizeof(int) * h.wedgePts.size());
    if (R9) std::memcpy(R9, h.wedgeR.data(), sizeof(double) * h.wedgeR.size());
    return (int)h.wedgePts.size();
}

void hs_wedge_face_t(const double* n, double* T)
{
    const double nn[3] = {n[0], n[1], n[2]};
    double t[9];
    qgd::wedgeFaceT(nn, t);
    std::memcpy(T, t, sizeof(t));
}

} // extern "C"
