// Internal declarations shared by the host set-up code, the kernels and the C-ABI layer.
// Device data model (DESIGN.md "Data layout in HBM"):
//   * gathered data   -> 64-byte AoS records (one or two 32-byte sectors per gather, 128-bit loads)
//   * streamed data   -> SoA arrays (perfectly coalesced)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qgd_b200.h"

namespace qgd {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void setLastError(const std::string& m);
#define QGD_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw qgd::Error(QGD_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " \
                                               + __FILE__ + ":" + std::to_string(__LINE__));             \
    } while (0)

// ---------------------------------------------------------------- runtime shared by the C-ABI translation units
void requireInit();                  // throws QGD_ERR_STATE before qgd_init
cudaStream_t runtimeStream();        // the library's compute stream
bool isCoeffsModel(const std::string& name);       // QGDCoeffs runTimeSelection table (QGDCoeffs.C:58-117)
std::string coeffsModelToc();
template <class F> int guarded(F&& fn)
{
    try { fn(); return QGD_OK; }
    catch (const Error& e) { setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { setLastError(e.what()); return QGD_ERR_INVALID; }
}

// ---------------------------------------------------------------- device buffers
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) QGD_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T>& h, cudaStream_t st = 0) {
        alloc(h.size());
        if (n) QGD_CUDA(cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, st));
        QGD_CUDA(cudaStreamSynchronize(st));
    }
    void zero(cudaStream_t st = 0) { if (n) QGD_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
};

// ---------------------------------------------------------------- multi-GPU plumbing shared by the solvers (qgd_abi.cu)
// Exchange lists of an extended sub-mesh: for neighbour k (rank nbr[k]) the local cells sendIds[sendOff[k] .. sendOff[k+1]) are
// packed and sent, the received block is scattered to recvIds[...]; ascending global id on both sides.
struct HaloLists {
    std::vector<int> nbr, sendOff, recvOff;
    DevBuf<int> sendIds, recvIds, offDev, roffDev;
    DevBuf<double> sendBuf, recvBuf;       // maxComp doubles per listed cell
    int maxComp = 0;
    bool active() const { return !nbr.empty(); }
    void set(int nn, const int* nbrRank, const int* sOff, const int* sIds, const int* rOff, const int* rIds, int nLocal, int nOwned,
             int maxComponents, cudaStream_t st);
};
int commRanks();                           // 1 before qgd_comm_init
// pack nComp fields (component k of cell c at base[k*stride + c]) -> grouped ncclSend/ncclRecv -> scatter; returns kernel launches
int commExchange(HaloLists& h, double* base, size_t stride, int nComp, cudaStream_t st);
enum CommOp { COMM_SUM = 0, COMM_MAX = 1, COMM_MIN = 2 };
void commAllReduce(double* dev, int count, CommOp op, cudaStream_t st);      // in place, device doubles

// ---------------------------------------------------------------- records
// cell / boundary-face state, gathered by the face kernels.  64 B each, 64-B aligned.
struct __align__(16) RecA { double rho, Ux, Uy, Uz, e, p, T, H; };
struct __align__(16) RecB { double rhoUx, rhoUy, rhoUz, rhoE, c, mu, alphaEff, aByC; };
// point values produced by the point gather (rho,U,e,p), 48 B
struct __align__(16) RecP { double rho, Ux, Uy, Uz, e, p; };

// face gradient record: grad(phi) = G1 (phi[v1]-phi[v3]) + G2 (phi[v2]-phi[v4]) + GP (phiP - phiN)
// which covers GaussVolPoint 3D quad (GaussVolPointBase3D.C:320-390), tri (:161-230), "other" faces
// (:760-768, G1=G2=0, GP=-nf*delta), 2D (GaussVolPointBase2D.C:122-169), 1D and `reduced`.
enum FaceFlags : int {
    FF_POINTS = 1,     // uses vertex values (G1/G2 non-zero)
    FF_TRI_QUIRK = 2,  // internal triangular face: vector-gradient index pattern of GaussVolPointBase3D.C:844-854
    FF_NORMAL_ONLY = 4, // boundary face evaluated as nf*snGrad (1D, reduced, other faces)
    FF_LSQ = 16        // internal face evaluated with the leastSquares cell stencil (FaceView::lsq*)
};

// ---------------------------------------------------------------- host-side derived mesh data
struct HostMesh {
    int nCells = 0, nFaces = 0, nInternal = 0, nPoints = 0, nPatches = 0, nBnd = 0;
    int nD = 3;
    int gD[3] = {1, 1, 1};
    int nOwned = 0;                                // cells owned by this rank (== nCells in serial runs)
    std::vector<int> coupledFace;                  // internal faces joining an owned and a halo cell (may be empty)
    std::vector<double> points, C, V, Cf, Sf, magSf, w, dC, ndC, nbrCC;
    std::vector<int> faceOff, faceVerts, owner, neighbour, patchStart, patchSize, patchKind;
    std::vector<int> bfacePatch;
    // derived
    std::vector<int> cfOff, cfEnc;                 // cell -> faces, enc = (face<<1)|isNeighbourSide
    std::vector<int> forcedDegFaces;               // faceSet degenerateStencilFaces (leastSquaresStencil.C:63-132), polyMesh face ids
    // DIC blocks: block id of every cell (empty = one block = the serial preconditioner).  In a decomposed reference run the DIC
    // factorisation and sweeps are local to each processor's lduMatrix; the device uses the same block-local form with blocks small
    // enough for one CTA (shared memory), so a sweep costs no grid-wide synchronisation (qgd_mesh_set_pcg_blocks / make_pcg_blocks)
    std::vector<int> pcgBlock;
    void makePcgBlocks(int targetCells);            // recursive coordinate bisection of the owned cells into compact tiles
    std::vector<int> pcOff, pcCell;                // non-patch point -> cells
    std::vector<double> pcW;
    std::vector<int> patchPoints;                  // list of patch points
    std::vector<int> ppOff, ppFace;                // patch point (by list position) -> boundary faces
    std::vector<double> ppW;
    std::vector<double> hQGDf, hQGD;
    // vertices of the constraint patches (wedge, symmetryPlane) and their constraint tensor R: volPointInterpolation turns a vector v
    // into R.v and a tensor T into R.T.R^T there [OF-v2312 pointConstraints: wedge / symmetryPlane point patch fields + constrainCorners;
    // pointConstraintI.H: one plane I - n n, two different planes d d with d along n1 x n2, three: 0]
    std::vector<int> wedgePts;
    std::vector<double> wedgeR;                    // 9 per listed point
    void build(const qgd_mesh_desc& d);
    // face gradient records for a scheme ("GaussVolPoint" / "reduced")
    // leastSquares scheme (extendedFaceStencilFindNeighbours.C:41-86, extendedFaceStencilCalculateWeights.C:43-155):
    // per internal face (polyMesh order) an ELL row of W neighbour cells and coefficient vectors wf2*Gdf; deg[f] = 1 where
    // the reference falls back to nf*snGrad (det G < 1)
    // opt: leastSquaresOpt keeps the (un-inverted) stencil on degenerate faces instead of the fallback
    void buildLeastSquares(bool opt, int& W, std::vector<int>& cells /*W*nInternal*/, std::vector<double>& coef /*W*3*nInternal*/,
                           std::vector<char>& deg /*nInternal*/) const;
    void buildFaceRecords(bool reduced, std::vector<int>& vtx /*nFaces*4*/, std::vector<int>& flags /*nFaces*/,
                          std::vector<double>& G /*9 arrays of nFaces, SoA: G[k*nFaces+f]*/,
                          std::vector<double>& halfDist /*nBnd*/) const;
};

} // namespace qgd

// ---------------------------------------------------------------- opaque handles
struct qgd_mesh {
    qgd::HostMesh h;
    // device copies
    // device face order: internal faces renumbered (tile, rank-in-owner, owner) so that a warp's faces have
    // consecutive owners AND consecutive neighbours/vertices; boundary faces keep their polyMesh position.
    std::vector<int> facePerm;     // device face -> polyMesh face
    std::vector<int> faceInv;      // polyMesh face -> device face
    qgd::DevBuf<int> facePermDev;
    int nIActive = 0;              // internal faces with an owned owner cell (first in device order)
    int faceStride = 0;            // column stride of the SoA arrays Sf[3] and G[9]: nFaces rounded up to a multiple of 16
    // ELL (column-major, width W) + CSR tail stencils: coalesced row access for thread-per-row kernels
    int pcEllW = 8, cfEllW = 6;
    qgd::DevBuf<int> pcEll, pcCount, pcTailOff, pcTailCell;      // point -> cells
    qgd::DevBuf<double> pcEllWt, pcTailW;
    qgd::DevBuf<int> cfEll, cfTailOff, cfTailEnc;                // cell -> faces, enc = (deviceFace<<1)|neighbourSide, -1 pad
    qgd::DevBuf<int> owner, neighbour, patchPoints, ppOff, ppFace, bfaceKind;
    qgd::DevBuf<double> ppW;
    qgd::DevBuf<double> Sf;        // SoA 3*faceStride
    qgd::DevBuf<double> magSf, w, dC, ndC, V, hQGDf, hQGD;
    qgd::DevBuf<int> wedgePts;     // HostMesh::wedgePts / wedgeR (empty without constraint patches)
    qgd::DevBuf<double> wedgeR;
};
