// Device-resident LDU PCG (lduMatrix PCG + DIC | diagonal | none, [OF-v2312], called at QHDpEqn.H:45).
// The whole solve — initial residual, normFactor, every iteration and the convergence test — runs inside ONE
// cooperative persistent kernel; vectors of a ~1M-cell problem stay resident in the 126 MB L2.
#pragma once
#include <functional>

#include "qgd_internal.h"

namespace qgd {

struct PcgResult { int iters; int pad; double res0, res, normFactor; };

// symmetric matrix in row form: ELL (column-major [W][n]) + CSR tail, entries of a row in ascending polyMesh face order;
// enc = (otherCell << 1) | lowerSide   (lowerSide: the other cell has the smaller index -> forward-sweep dependency)
struct PcgView {
    int n, W;
    const int* enc; const double* coef;
    const int* tailOff; const int* tailEnc; const double* tailCoef;
    const double* diag;          // incl. boundary (internalCoeffs) contribution
    const double* rD;            // precond 1: 1/diag ; precond 2: DIC reciprocal diagonal ; precond 0: unused
    const double* b;
    double* x;
    double* r; double* w; double* z; double* p0; double* p1;
    double* partials;            // [2 slots][2 values][gridDim.x]
    int nLevels; const int* lvlOff; const int* lvlCells;     // DIC level schedule (cells grouped by dependency depth)
    // block-local DIC (HostMesh::pcgBlock): cells ordered (block, level, id); position p of that order holds cell bCells[p].
    // Block b owns positions [bOff[b], bOff[b+1]) and the level offsets bLvlOff[bLvlStart[b] .. bLvlStart[b+1]] (absolute
    // positions).  Rows in block order, in-block entries only, ELL width Wb: bEnc[j*n + p] = (local position of the other cell
    // << 1) | lowerSide, -1 = none; bCoef the matching coefficient.  nBlocks = 0: the global level schedule above.
    int nBlocks, Wb, maxBlockCells;
    int dicWarp, dicUnitSmem;    // work unit of the block sweeps: a warp (8 blocks in flight per CTA) or the CTA; bytes of shared memory per unit
    const int* bOff; const int* bLvlStart; const int* bLvlOff; const int* bCells; const int* bEnc; const double* bCoef;
    double tol, relTol; int maxIter, precond;
    PcgResult* out;
};

struct PcgMatrix {
    int n = 0, W = 0, nLevels = 0, precond = -1;
    DevBuf<int> enc, tailOff, tailEnc, lvlOff, lvlCells;
    DevBuf<double> coef, tailCoef, diag, rD, b, x, r, w, z, p0, p1, partials;
    DevBuf<PcgResult> out;
    DevBuf<int> encFace, tailFace;   // device face id of every ELL / tail entry (-1: padding), for refresh()
    // block-local DIC (see PcgView)
    int nBlocks = 0, Wb = 0, maxBlockCells = 0, maxBlockLevels = 0;
    bool dicWarp = false;
    DevBuf<int> bOff, bLvlStart, bLvlOff, bCells, bEnc, bEncFace;
    DevBuf<double> bCoef;
    size_t dicUnitSmem() const { return nBlocks ? (((size_t)maxBlockCells * (16 + 12 * (size_t)Wb) + 4 * ((size_t)maxBlockLevels + 2) + 15) & ~(size_t)15) : 0; }
    size_t dicSmemBytes() const { return dicUnitSmem() * (dicWarp ? 8 : 1); }
    int gridBlocks = 0;
    double* bExternal = nullptr;     // when set, the right-hand side lives in the caller's array
    double* xExternal = nullptr;     // when set, the solution vector lives in the caller's array (e.g. the p slice of the QHD state)
    // diag: nCells (boundary contributions already added), upper: nInternal in polyMesh face order
    void build(const HostMesh& h, const double* diag, const double* upper, int precond, cudaStream_t st,
               const std::vector<int>* faceInv = nullptr);
    // new coefficients on the same addressing, all on the device: upper[f] = -faceCoef[deviceFace f], diagonal = diagDev;
    // the preconditioner is rebuilt (needs build(..., faceInv))
    void refresh(const double* faceCoef, const double* diagDev, cudaStream_t st);
    PcgView view(double tol, double relTol, int maxIter) const;
    // solves A x = b for the device vectors b, x (in place); returns kernel launches issued
    int solve(double tol, double relTol, int maxIter, cudaStream_t st);
};

int dicBlocksGrid(const PcgMatrix& A);
void launchDicBlocks(const PcgMatrix& A, const double* r, double* z, double* partials, const int* done, cudaStream_t st);

// ---- stepwise (one kernel per phase) form of the same solver for decomposed runs, see qgd_mpcg.cu
struct PcgHooks {
    std::function<void(double* vec, cudaStream_t st)> exchange;                   // fill the halo entries of a cell vector from their owners
    std::function<void(double* dev, int count, cudaStream_t st)> allreduceSum;    // in-place global sum of device doubles
};
struct StepwisePcg {
    int n = 0, nRows = 0, grid = 1;
    DevBuf<double> r, w, p, partials, red;
    DevBuf<int> state;
    void alloc(const PcgMatrix& A, int rows);     // rows = owned rows [0, rows) of the extended sub-mesh matrix (== A.n on one GPU)
    // b: rows device doubles ; x: A.n device doubles ; precond 0 none | 1 diagonal | 2 DIC on the matrix's blocks (PcgMatrix::nBlocks > 0) ;
    // hooks = nullptr on one GPU ; returns launches
    int solve(const PcgMatrix& A, const double* b, double* x, double tol, double relTol, int maxIter, int precond, cudaStream_t st,
              const PcgHooks* hooks, PcgResult* result);
};

} // namespace qgd
