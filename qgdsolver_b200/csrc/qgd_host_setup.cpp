// Host-side set-up: everything the reference derives once per mesh, recast as the compact
// device records the kernels stream.  Runs once in qgd_mesh_create / qgd_fvsc_create.
//
// Reference behaviour reproduced (formulas re-derived in vector form):
//   volPointInterpolation weights                [OF-v2312]  used at GaussVolPointBase3D.C:43-46
//   QGDCoeffs::updateQGDLength                   QGDCoeffs.C:195-199, 298-376
//   GaussVolPointBase3D ctor / tri / quad weights GaussVolPointBase3D.C:74-476
//   GaussVolPointBase2D ctor                     GaussVolPointBase2D.C:72-293
//   GaussVolPointBase1D, reduced                 GaussVolPointBase1D.C:49-79, reducedFaceNormalStencil.C:69-108
#include <algorithm>
#include <cmath>

#include "qgd_internal.h"
#include "qgd_varsc5.h"

namespace qgd {

namespace {
struct Vec3 {
    double v[3];
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
inline Vec3 at(const std::vector<double>& a, long i) { return {{a[3 * i], a[3 * i + 1], a[3 * i + 2]}}; }
inline Vec3 sub(Vec3 a, Vec3 b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
inline Vec3 add(Vec3 a, Vec3 b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
inline Vec3 scale(double s, Vec3 a) { return {{s * a[0], s * a[1], s * a[2]}}; }
inline double dot3(Vec3 a, Vec3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 cross3(Vec3 a, Vec3 b) { return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}}; }
inline double norm3(Vec3 a) { return std::sqrt(dot3(a, a)); }
} // namespace

void HostMesh::build(const qgd_mesh_desc& d)
{
    if (!d.points || !d.face_offsets || !d.face_verts || !d.owner || !d.C || !d.V || !d.Cf || !d.Sf || !d.magSf ||
        !d.weights || !d.deltaCoeffs || !d.nonOrthDeltaCoeffs || (d.n_internal_faces > 0 && !d.neighbour) ||
        (d.n_patches > 0 && (!d.patch_start || !d.patch_size || !d.patch_kind)))
        throw Error(QGD_ERR_INVALID, "qgd_mesh_create: null array in mesh descriptor");
    if (d.n_cells <= 0 || d.n_faces < d.n_internal_faces || d.n_points <= 0)
        throw Error(QGD_ERR_INVALID, "qgd_mesh_create: inconsistent mesh sizes");
    nCells = d.n_cells; nFaces = d.n_faces; nInternal = d.n_internal_faces; nPoints = d.n_points; nPatches = d.n_patches;
    nBnd = nFaces - nInternal;
    for (int i = 0; i < 3; ++i) gD[i] = d.geometric_d[i];
    nD = (gD[0] > 0) + (gD[1] > 0) + (gD[2] > 0);
    nOwned = (d.n_owned_cells > 0 && d.n_owned_cells <= nCells) ? d.n_owned_cells : nCells;
    if (d.coupled_internal_face) coupledFace.assign(d.coupled_internal_face, d.coupled_internal_face + nInternal);
    points.assign(d.points, d.points + 3 * (size_t)nPoints);
    faceOff.assign(d.face_offsets, d.face_offsets + nFaces + 1);
    faceVerts.assign(d.face_verts, d.face_verts + faceOff[nFaces]);
    owner.assign(d.owner, d.owner + nFaces);
    neighbour.assign(d.neighbour, d.neighbour + nInternal);
    patchStart.assign(d.patch_start, d.patch_start + nPatches);
    patchSize.assign(d.patch_size, d.patch_size + nPatches);
    patchKind.assign(d.patch_kind, d.patch_kind + nPatches);
    C.assign(d.C, d.C + 3 * (size_t)nCells);
    V.assign(d.V, d.V + nCells);
    Cf.assign(d.Cf, d.Cf + 3 * (size_t)nFaces);
    Sf.assign(d.Sf, d.Sf + 3 * (size_t)nFaces);
    magSf.assign(d.magSf, d.magSf + nFaces);
    w.assign(d.weights, d.weights + nFaces);
    dC.assign(d.deltaCoeffs, d.deltaCoeffs + nFaces);
    ndC.assign(d.nonOrthDeltaCoeffs, d.nonOrthDeltaCoeffs + nFaces);
    if (d.neighb_cell_centres) nbrCC.assign(d.neighb_cell_centres, d.neighb_cell_centres + 3 * (size_t)nBnd);
    else nbrCC.assign(3 * (size_t)nBnd, 0.0);

    // ---- validation of polyMesh invariants the kernels rely on
    for (int f = 0; f < nFaces; ++f)
        if (owner[f] < 0 || owner[f] >= nCells) throw Error(QGD_ERR_INVALID, "qgd_mesh_create: owner out of range");
    for (int f = 0; f < nInternal; ++f)
        if (neighbour[f] < 0 || neighbour[f] >= nCells) throw Error(QGD_ERR_INVALID, "qgd_mesh_create: neighbour out of range");
    for (size_t q = 0; q < faceVerts.size(); ++q)
        if (faceVerts[q] < 0 || faceVerts[q] >= nPoints) throw Error(QGD_ERR_INVALID, "qgd_mesh_create: face vertex out of range");
    bfacePatch.assign(nBnd, -1);
    for (int pi = 0; pi < nPatches; ++pi) {
        if (patchStart[pi] < nInternal || patchStart[pi] + patchSize[pi] > nFaces)
            throw Error(QGD_ERR_INVALID, "qgd_mesh_create: patch range outside the boundary faces");
        for (int i = 0; i < patchSize[pi]; ++i) bfacePatch[patchStart[pi] - nInternal + i] = pi;
    }
    for (int b = 0; b < nBnd; ++b)
        if (bfacePatch[b] < 0) throw Error(QGD_ERR_INVALID, "qgd_mesh_create: boundary face not covered by any patch");

    // ---- cell -> faces with side bit (ascending face index per cell)
    {
        std::vector<int> cnt(nCells + 1, 0);
        for (int f = 0; f < nFaces; ++f) cnt[owner[f] + 1]++;
        for (int f = 0; f < nInternal; ++f) cnt[neighbour[f] + 1]++;
        for (int c = 0; c < nCells; ++c) cnt[c + 1] += cnt[c];
        cfOff = cnt;
        cfEnc.assign(cnt[nCells], 0);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int f = 0; f < nFaces; ++f) {
            cfEnc[pos[owner[f]]++] = (f << 1);
            if (f < nInternal) cfEnc[pos[neighbour[f]]++] = (f << 1) | 1;
        }
    }

    // ---- point classification: points on non-empty, non-coupled boundary faces take boundary values
    std::vector<char> isPatchPoint(nPoints, 0);
    auto isPatchFace = [&](int b) { const int k = patchKind[bfacePatch[b]]; return k != QGD_PATCH_EMPTY && k != QGD_PATCH_PROCESSOR; };
    for (int b = 0; b < nBnd; ++b)
        if (isPatchFace(b))
            for (int q = faceOff[nInternal + b]; q < faceOff[nInternal + b + 1]; ++q) isPatchPoint[faceVerts[q]] = 1;

    // ---- point -> cells CSR (two passes, duplicates removed through a per-point sort)
    {
        std::vector<int> cnt(nPoints + 1, 0);
        for (int f = 0; f < nFaces; ++f) {
            const int m = (f < nInternal) ? 2 : 1;
            for (int q = faceOff[f]; q < faceOff[f + 1]; ++q) cnt[faceVerts[q] + 1] += m;
        }
        for (int p = 0; p < nPoints; ++p) cnt[p + 1] += cnt[p];
        std::vector<int> raw(cnt[nPoints]);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int f = 0; f < nFaces; ++f)
            for (int q = faceOff[f]; q < faceOff[f + 1]; ++q) {
                const int p = faceVerts[q];
                raw[pos[p]++] = owner[f];
                if (f < nInternal) raw[pos[p]++] = neighbour[f];
            }
        pcOff.assign(nPoints + 1, 0);
        pcCell.clear();
        pcCell.reserve(raw.size() / 3);
        for (int p = 0; p < nPoints; ++p) {
            if (!isPatchPoint[p]) {
                auto b = raw.begin() + cnt[p], e = raw.begin() + cnt[p + 1];
                std::sort(b, e);
                e = std::unique(b, e);
                pcCell.insert(pcCell.end(), b, e);
            }
            pcOff[p + 1] = (int)pcCell.size();
        }
        pcW.resize(pcCell.size());
        for (int p = 0; p < nPoints; ++p) {
            double sum = 0.0;
            const Vec3 x = at(points, p);
            for (int q = pcOff[p]; q < pcOff[p + 1]; ++q) { pcW[q] = 1.0 / norm3(sub(x, at(C, pcCell[q]))); sum += pcW[q]; }
            for (int q = pcOff[p]; q < pcOff[p + 1]; ++q) pcW[q] /= sum;
        }
    }
    // ---- patch points -> boundary faces
    {
        std::vector<int> slot(nPoints, -1);
        patchPoints.clear();
        for (int p = 0; p < nPoints; ++p) if (isPatchPoint[p]) { slot[p] = (int)patchPoints.size(); patchPoints.push_back(p); }
        const int nPP = (int)patchPoints.size();
        std::vector<int> cnt(nPP + 1, 0);
        for (int b = 0; b < nBnd; ++b)
            if (isPatchFace(b))
                for (int q = faceOff[nInternal + b]; q < faceOff[nInternal + b + 1]; ++q) cnt[slot[faceVerts[q]] + 1]++;
        for (int i = 0; i < nPP; ++i) cnt[i + 1] += cnt[i];
        ppOff = cnt;
        ppFace.assign(cnt[nPP], 0);
        ppW.assign(cnt[nPP], 0.0);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int b = 0; b < nBnd; ++b)
            if (isPatchFace(b))
                for (int q = faceOff[nInternal + b]; q < faceOff[nInternal + b + 1]; ++q) ppFace[pos[slot[faceVerts[q]]]++] = b;
        for (int i = 0; i < nPP; ++i) {
            double sum = 0.0;
            const Vec3 x = at(points, patchPoints[i]);
            for (int q = ppOff[i]; q < ppOff[i + 1]; ++q) { ppW[q] = 1.0 / norm3(sub(x, at(Cf, nInternal + ppFace[q]))); sum += ppW[q]; }
            for (int q = ppOff[i]; q < ppOff[i + 1]; ++q) ppW[q] /= sum;
        }
    }

    // ---- vertices of the constraint patches (wedge, symmetryPlane): every such patch a vertex lies on adds its planar normal (that of
    // the patch's first face, as the point patch fields take pointNormals()[0]) to the vertex's pointConstraint, in patch order
    {
        std::vector<int> cnt(nPoints, 0);
        std::vector<Vec3> dir(nPoints, Vec3{{0, 0, 0}});
        std::vector<int> stamp(nPoints, -1);
        for (int pi = 0; pi < nPatches; ++pi) {
            if ((patchKind[pi] != QGD_PATCH_WEDGE && patchKind[pi] != QGD_PATCH_SYMMETRY_PLANE) || patchSize[pi] == 0) continue;
            const int f0 = patchStart[pi];
            const Vec3 n = scale(1.0 / magSf[f0], at(Sf, f0));
            for (int f = f0; f < f0 + patchSize[pi]; ++f)
                for (int q = faceOff[f]; q < faceOff[f + 1]; ++q) {
                    const int p = faceVerts[q];
                    if (stamp[p] == pi) continue;
                    stamp[p] = pi;
                    if (cnt[p] == 0) { cnt[p] = 1; dir[p] = n; }                                   // pointConstraint::applyConstraint
                    else if (cnt[p] == 1) { const Vec3 pl = cross3(n, dir[p]); const double mp = norm3(pl); if (mp > 1e-3) { cnt[p] = 2; dir[p] = scale(1.0 / mp, pl); } }
                    else if (cnt[p] == 2) { if (std::fabs(dot3(n, dir[p])) > 1e-3) { cnt[p] = 3; dir[p] = Vec3{{0, 0, 0}}; } }
                }
        }
        wedgePts.clear(); wedgeR.clear();
        for (int p = 0; p < nPoints; ++p) {
            if (!cnt[p]) continue;
            wedgePts.push_back(p);
            for (int a = 0; a < 3; ++a)                                                            // constraintTransformation
                for (int b = 0; b < 3; ++b)
                    wedgeR.push_back(cnt[p] == 1 ? (a == b ? 1.0 : 0.0) - dir[p][a] * dir[p][b] : (cnt[p] == 2 ? dir[p][a] * dir[p][b] : 0.0));
        }
    }

    // ---- QGD length scales
    hQGDf.assign(nFaces, 0.0);
    for (int f = 0; f < nInternal; ++f) {
        const Vec3 cf = at(Cf, f);
        hQGDf[f] = 2.0 * std::min(norm3(sub(at(C, owner[f]), cf)), norm3(sub(at(C, neighbour[f]), cf)));
        if (!coupledFace.empty() && coupledFace[f]) hQGDf[f] = 1.0 / std::fabs(dC[f]);   // processor-patch rule
    }
    for (int b = 0; b < nBnd; ++b) {
        const int f = nInternal + b;
        const double h = 1.0 / std::fabs(dC[f]);
        hQGDf[f] = (patchKind[bfacePatch[b]] == QGD_PATCH_PROCESSOR) ? h : 2.0 * h;
    }
    hQGD.assign(nCells, 0.0);
    for (int c = 0; c < nCells; ++c) {
        double hs = 0.0, s = 0.0;
        for (int q = cfOff[c]; q < cfOff[c + 1]; ++q) {
            const int f = cfEnc[q] >> 1;
            if (f >= nInternal) {
                const int k = patchKind[bfacePatch[f - nInternal]];
                if (k == QGD_PATCH_EMPTY || k == QGD_PATCH_WEDGE) continue;
            }
            hs += hQGDf[f] * magSf[f];
            s += magSf[f];
        }
        hQGD[c] = hs / s;
    }
}

void HostMesh::buildFaceRecords(bool reduced, std::vector<int>& vtx, std::vector<int>& flags, std::vector<double>& G,
                                std::vector<double>& halfDist) const
{
    const size_t nF = (size_t)nFaces;
    vtx.assign(4 * nF, 0);
    flags.assign(nF, 0);
    G.assign(9 * nF, 0.0);
    halfDist.assign(nBnd, 0.0);
    auto setG = [&](size_t f, int slot, Vec3 g) { for (int i = 0; i < 3; ++i) G[(size_t)(3 * slot + i) * nF + f] = g[i]; };
    // 2D frame
    int ie1 = 0, ie2 = 1, ie3 = 2;
    if (nD == 2) {
        for (int d = 0; d < 3; ++d) if (gD[d] < 1) ie3 = d;
        ie1 = (ie3 == 0) ? 1 : 0;
        ie2 = (ie3 == 2) ? 1 : 2;
    }
    int ip1 = -1, ip3 = -1;   // the reference carries these over between internal faces (GaussVolPointBase2D.C:88,129-147)
    for (int f = 0; f < nFaces; ++f) {
        const bool internal = f < nInternal;
        const int b = f - nInternal;
        const int kind = internal ? -1 : patchKind[bfacePatch[b]];
        const Vec3 nf = scale(1.0 / magSf[f], at(Sf, f));
        const Vec3 cP = at(C, owner[f]);
        Vec3 cN;
        if (internal) cN = at(C, neighbour[f]);
        else if (kind == QGD_PATCH_PROCESSOR) cN = at(nbrCC, b);
        else cN = add(cP, scale(2.0, sub(at(Cf, f), cP)));
        const Vec3 d = sub(cN, cP);
        if (!internal) halfDist[b] = 0.5 * norm3(d);
        if (!internal && kind == QGD_PATCH_EMPTY) continue;
        const int nv = faceOff[f + 1] - faceOff[f];
        const int* fv = &faceVerts[faceOff[f]];
        const bool normalOnly = reduced || nD == 1 || (nD == 3 && nv != 3 && nv != 4);
        if (normalOnly) {
            if (internal) setG(f, 2, scale(-ndC[f], nf));       // nf*snGrad = -nf*delta*(phiP-phiN)
            else { flags[f] = FF_NORMAL_ONLY; setG(f, 2, nf); }
            continue;
        }
        if (nD == 2) {
            if (!internal && kind == QGD_PATCH_WEDGE) continue;                 // result stays zero
            if (!internal) { ip1 = -1; ip3 = -1; }
            const double zRef = internal ? cN[ie3] : cP[ie3];
            for (int q = 0; q < nv; ++q) if (points[3 * (size_t)fv[q] + ie3] >= zRef) { ip1 = fv[q]; break; }
            for (int q = 0; q < nv; ++q) if (points[3 * (size_t)fv[q] + ie3] >= zRef && fv[q] != ip1) { ip3 = fv[q]; break; }
            if (ip1 < 0 || ip3 < 0) throw Error(QGD_ERR_INVALID, "GaussVolPoint 2D: face without two vertices on the upper plane");
            const Vec3 v13 = sub(at(points, ip3), at(points, ip1));
            const double m42 = norm3(d), m13 = norm3(v13);
            const double cosa1 = d[ie1] / m42, sina1 = d[ie2] / m42, cosa2 = v13[ie1] / m13, sina2 = v13[ie2] / m13;
            const double den = sina2 * cosa1 - sina1 * cosa2;
            const double c1 = sina2 / den, c2 = sina1 / den, c3 = cosa1 / den, c4 = cosa2 / den;
            Vec3 gp{{0, 0, 0}}, g1{{0, 0, 0}};
            gp[ie1] = -c1 / m42; gp[ie2] = c4 / m42;           // multiplies (phiP - phiN)
            g1[ie1] = c2 / m13;  g1[ie2] = -c3 / m13;          // multiplies (phi[ip1] - phi[ip3])
            setG(f, 0, g1); setG(f, 2, gp);
            vtx[4 * (size_t)f] = ip1; vtx[4 * (size_t)f + 1] = ip1; vtx[4 * (size_t)f + 2] = ip3; vtx[4 * (size_t)f + 3] = ip1;
            flags[f] = FF_POINTS;
            continue;
        }
        // 3D diamond (bipyramid) Gauss gradient
        if (nv == 4) {
            const Vec3 p1 = at(points, fv[0]), p2 = at(points, fv[1]), p3 = at(points, fv[2]), p4 = at(points, fv[3]);
            const Vec3 e1 = sub(p2, p4), e2 = sub(p3, p1);
            const double D = dot3(e2, cross3(e1, d));
            setG(f, 0, scale(1.0 / D, cross3(d, e1)));
            setG(f, 1, scale(1.0 / D, cross3(d, e2)));
            setG(f, 2, scale(1.0 / D, cross3(e1, e2)));
            for (int q = 0; q < 4; ++q) vtx[4 * (size_t)f + q] = fv[q];
            flags[f] = FF_POINTS;
        } else {
            const Vec3 p1 = at(points, fv[0]), p2 = at(points, fv[1]), p3 = at(points, fv[2]);
            const Vec3 n2 = cross3(sub(p2, p1), sub(p3, p1));
            const double D = -dot3(n2, d);
            setG(f, 0, scale(1.0 / D, cross3(d, sub(p2, p3))));
            setG(f, 1, scale(1.0 / D, cross3(d, sub(p3, p1))));
            setG(f, 2, scale(1.0 / D, n2));
            vtx[4 * (size_t)f] = fv[0]; vtx[4 * (size_t)f + 1] = fv[1]; vtx[4 * (size_t)f + 2] = fv[2]; vtx[4 * (size_t)f + 3] = fv[2];
            flags[f] = FF_POINTS | (internal ? FF_TRI_QUIRK : 0);
        }
    }
}

// Recursive coordinate bisection: split the owned cells along the longest extent of their centres at the median until a part
// holds at most targetCells; parts are numbered in creation order (deterministic).  Halo cells get block -1.
void HostMesh::makePcgBlocks(int targetCells)
{
    if (targetCells < 1) targetCells = 1;
    pcgBlock.assign(nCells, -1);
    std::vector<int> ids(nOwned);
    for (int c = 0; c < nOwned; ++c) ids[c] = c;
    int next = 0;
    std::vector<std::pair<int, int>> stack{{0, nOwned}};
    while (!stack.empty()) {
        const auto [lo, hi] = stack.back();
        stack.pop_back();
        if (hi - lo <= targetCells) { for (int i = lo; i < hi; ++i) pcgBlock[ids[i]] = next; ++next; continue; }
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = lo; i < hi; ++i)
            for (int d = 0; d < 3; ++d) { const double x = C[3 * (size_t)ids[i] + d]; mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x); }
        int dir = 0;
        for (int d = 1; d < 3; ++d) if (gD[d] > 0 && (gD[dir] <= 0 || mx[d] - mn[d] > mx[dir] - mn[dir])) dir = d;
        const int mid = lo + (hi - lo) / 2;
        std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int a, int b) {
            const double xa = C[3 * (size_t)a + dir], xb = C[3 * (size_t)b + dir];
            return xa < xb || (xa == xb && a < b);
        });
        stack.push_back({mid, hi});
        stack.push_back({lo, mid});
    }
}

void HostMesh::buildLeastSquares(bool opt, int& W, std::vector<int>& cells, std::vector<double>& coef, std::vector<char>& deg) const
{
    const double SMALLv = 1e-15;
    // [OF-v2312] primitiveMesh::pointCells(): ascending cell order per point
    std::vector<std::vector<int>> pc(nPoints);
    for (int f = 0; f < nFaces; ++f)
        for (int q = faceOff[f]; q < faceOff[f + 1]; ++q) {
            pc[faceVerts[q]].push_back(owner[f]);
            if (f < nInternal) pc[faceVerts[q]].push_back(neighbour[f]);
        }
    for (auto& v : pc) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    std::vector<std::vector<int>> nbs(nInternal);
    W = 1;
    for (int f = 0; f < nInternal; ++f) {
        std::vector<int>& nb = nbs[f];
        for (int q = faceOff[f]; q < faceOff[f + 1]; ++q)
            for (int c : pc[faceVerts[q]]) if (std::find(nb.begin(), nb.end(), c) == nb.end()) nb.push_back(c);
        W = std::max(W, (int)nb.size());
    }
    const size_t nI = (size_t)std::max(nInternal, 1);
    cells.assign((size_t)W * nI, 0);
    coef.assign((size_t)W * 3 * nI, 0.0);
    deg.assign(nI, 0);
    for (int f = 0; f < nInternal; ++f) {
        const std::vector<int>& nb = nbs[f];
        const Vec3 cf = at(Cf, f);
        std::vector<Vec3> df(nb.size());
        std::vector<double> wf2(nb.size());
        double G[6] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < nb.size(); ++i) {
            df[i] = sub(at(C, nb[i]), cf);
            wf2[i] = 1.0 / dot3(df[i], df[i]);
            const Vec3 d = df[i];
            const double a[6] = {d[0] * d[0], d[0] * d[1], d[0] * d[2], d[1] * d[1], d[1] * d[2], d[2] * d[2]};
            for (int t = 0; t < 6; ++t) G[t] += a[t] * wf2[i];
        }
        double G0[6] = {0, 0, 0, 0, 0, 0};
        if (std::fabs(G[0]) < SMALLv) G0[0] = 1;
        if (std::fabs(G[3]) < SMALLv) G0[3] = 1;
        if (std::fabs(G[5]) < SMALLv) G0[5] = 1;
        for (int t = 0; t < 6; ++t) G[t] += G0[t];
        const double xx = G[0], xy = G[1], xz = G[2], yy = G[3], yz = G[4], zz = G[5];
        const double detG = xx * yy * zz + xy * yz * xz + xz * xy * yz - xx * yz * yz - xy * xy * zz - xz * yy * xz;
        for (size_t j = 0; j < (size_t)W; ++j) cells[j * nI + f] = owner[f];
        double Gi[6] = {G[0], G[1], G[2], G[3], G[4], G[5]};
        if (detG < 1) {
            if (!opt) { deg[f] = 1; continue; }      // leastSquares: nf*snGrad ; leastSquaresOpt: un-inverted (G+G0)&df
        } else {
        Gi[0] = (yy * zz - yz * yz) / detG; Gi[1] = (xz * yz - xy * zz) / detG; Gi[2] = (xy * yz - xz * yy) / detG;
        Gi[3] = (xx * zz - xz * xz) / detG; Gi[4] = (xy * xz - xx * yz) / detG; Gi[5] = (xx * yy - xy * xy) / detG;
        for (int t = 0; t < 6; ++t) Gi[t] -= G0[t];
        }
        for (size_t i = 0; i < nb.size(); ++i) {
            const Vec3 d = df[i];
            const double gd[3] = {Gi[0] * d[0] + Gi[1] * d[1] + Gi[2] * d[2], Gi[1] * d[0] + Gi[3] * d[1] + Gi[4] * d[2],
                                  Gi[2] * d[0] + Gi[4] * d[1] + Gi[5] * d[2]};
            cells[i * nI + f] = nb[i];
            for (int q = 0; q < 3; ++q) coef[(i * 3 + q) * nI + f] = wf2[i] * gd[q];
        }
    }
}

} // namespace qgd

// ---------------------------------------------------------------- varScModel5: time-constant host data (qgd_varsc5.h)
namespace qgd {

// mesh.cells() lists, the position of every internal face inside them, and the mesh-quality floor cqSc of varScModel5.C:112-132:
// primitiveMeshTools::cellClosedness [OF-v2312, as remembered]: per cell the component-wise sums of |Sf| over all its faces;
// aspectRatio = max / min of the sums over the solved directions, in 3D at least (1/6) (sum of all three) / V^(2/3);
// cqSc = badQualitySc * aspectRatio / maxAspectRatio where aspectRatio > maxAspectRatio, else 0
void buildVarSc5Host(const HostMesh& h, double badQualitySc, double maxAspectRatio, VarSc5Host& o)
{
    const int nC = h.nCells, nF = h.nFaces, nI = h.nInternal, nB = h.nBnd;
    o.nC = nC; o.nI = nI; o.nB = nB;
    o.own = h.owner; o.nei = h.neighbour;
    o.w.assign(h.w.begin(), h.w.begin() + nI);
    o.Sf = h.Sf;
    o.bMagSf.assign(h.magSf.begin() + nI, h.magSf.end());
    o.bDC.assign(h.dC.begin() + nI, h.dC.end());
    o.bHf.assign(h.hQGDf.begin() + nI, h.hQGDf.end());
    o.bKind.resize(nB);
    for (int b = 0; b < nB; ++b) o.bKind[b] = h.patchKind[h.bfacePatch[b]];
    // primitiveMesh::calcCells: owner loop over all faces, then neighbour loop over the internal faces
    o.ccOff.assign(nC + 1, 0);
    for (int f = 0; f < nF; ++f) o.ccOff[h.owner[f] + 1]++;
    for (int f = 0; f < nI; ++f) o.ccOff[h.neighbour[f] + 1]++;
    o.maxCellFaces = 0;
    for (int c = 0; c < nC; ++c) { o.maxCellFaces = std::max(o.maxCellFaces, o.ccOff[c + 1]); o.ccOff[c + 1] += o.ccOff[c]; }
    o.ccFace.assign(o.ccOff[nC], -1);
    o.lidxOwn.assign(nI, -1); o.lidxNei.assign(nI, -1);
    std::vector<int> fill(nC, 0);
    for (int f = 0; f < nF; ++f) {
        const int c = h.owner[f];
        if (f < nI) o.lidxOwn[f] = fill[c];
        o.ccFace[o.ccOff[c] + fill[c]++] = f;
    }
    for (int f = 0; f < nI; ++f) {
        const int c = h.neighbour[f];
        o.lidxNei[f] = fill[c];
        o.ccFace[o.ccOff[c] + fill[c]++] = f;
    }
    // cellClosedness
    std::vector<double> sumMag(3 * (size_t)nC, 0.0);
    for (int f = 0; f < nF; ++f)
        for (int d = 0; d < 3; ++d) sumMag[3 * (size_t)h.owner[f] + d] += std::fabs(h.Sf[3 * (size_t)f + d]);
    for (int f = 0; f < nI; ++f)
        for (int d = 0; d < 3; ++d) sumMag[3 * (size_t)h.neighbour[f] + d] += std::fabs(h.Sf[3 * (size_t)f + d]);
    const double ROOTVSMALL = 1.0e-150, VGREAT = 1.0e+300;
    o.aspectRatio.assign(nC, 1.0); o.cqSc.assign(nC, 0.0);
    for (int c = 0; c < nC; ++c) {
        double minC = VGREAT, maxC = -VGREAT;
        for (int d = 0; d < 3; ++d)
            if (h.gD[d] == 1) { minC = std::min(minC, sumMag[3 * (size_t)c + d]); maxC = std::max(maxC, sumMag[3 * (size_t)c + d]); }
        double ar = maxC / (minC + ROOTVSMALL);
        if (h.nD == 3) {
            const double v = std::max(ROOTVSMALL, h.V[c]);
            ar = std::max(ar, 1.0 / 6.0 * (sumMag[3 * (size_t)c] + sumMag[3 * (size_t)c + 1] + sumMag[3 * (size_t)c + 2]) / std::pow(v, 2.0 / 3.0));
        }
        o.aspectRatio[c] = ar;
        if (ar > maxAspectRatio) o.cqSc[c] = badQualitySc * ar / maxAspectRatio;
    }
}

} // namespace qgd
