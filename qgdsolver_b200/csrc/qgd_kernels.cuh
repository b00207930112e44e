// Kernel-side declarations: device views (POD structs of raw pointers passed by value) and launchers.
#pragma once
#include <functional>

#include "qgd_internal.h"

namespace qgd {

// thermo + model constants (perfectGas / hConst / sensibleInternalEnergy / constTransport / constScPrModel1)
struct Consts {
    double R, Cp, Cv, Tref, Hsref, mu, Pr, ScQGD, PrQGD, gamma;
    int alphaEffGamma, energyQuirk, reducedScheme;
    // QGDCoeffs model: 0 constScPrModel1, 1 constScPrModel1n, 2 constScPrModel2.
    // tauMode selects how tauQGDf is formed from the per-cell slot `aByC` of the state:
    //   0: I(alphaQGD/c)*hQGDf            slot = alphaQGD/c         constScPrModel1.C:103, constScPrModel2.C:82
    //   1: I(tauQGD)                      slot = tauQGD = alphaQGD*hQGD/(|U|+c)   constScPrModel1n.C:126-128
    //   2: I(alphaQGD)*hQGDf/I(c)         slot = alphaQGD   (constScPrModel1n before "U" is registered, i.e. the first step)
    int model, tauMode;
    int implicit;            // QGD::implicitDiffusion: the mu / alpha terms leave the explicit fluxes (updateFluxes.H:95-111,131-135)
    // varScModel6 / varScModel7 (model stays 0: tau as constScPrModel1): per-cell ScQGD from the pressure-jump sensor
    // (k_varsc), boundary ScQGD = ScB = the dictionary value clamped by model 7's minSc / maxSc
    int varSc;               // 0 | 6 | 7
    double cSc1, minSc, maxSc, ScB;
    // thermo-type instantiations of psiQGDThermos.C:65-111 besides const + hConst:
    //   transport 1 powerLaw (powerLawTransportI.H:120-150), 2 sutherland [OF-v2312 sutherlandTransportI.H]; eConst [OF-v2312 eConstThermoI.H]
    int transport, eConst;
    double mu0, T0, kExp, rPr, As, Ts, Esref;
};

struct FaceView {            // all faces: internal [0,nI) then boundary [nI,nF)
    int nI, nF, nB;
    int fs;                  // stride of the per-face SoA arrays G[9] and Sf[3] (nF rounded up to 16: every column 128-B aligned, so the
                             // cp.async.bulk tile copies of k_face_flux_tma are legal for any face count)
    int nIActive;            // internal faces whose owner is an owned cell (device order puts the others last)
    int zeroDivCmpt;         // 2D: out-of-plane component of Div(tensor) stays 0 (GaussVolPointBase2D.C:447-485), else -1
    const int* own;          // nF
    const int* nei;          // nI
    const int4* vtx;         // nF
    const int* flags;        // nF
    const double* Sf;        // SoA 3*nF
    const double* magSf;     // nF
    const double* w;         // nF
    const double* hf;        // nF  hQGDf
    const double* dC;        // nF  deltaCoeffs
    const double* ndC;       // nF  nonOrthDeltaCoeffs
    const double* G;         // SoA 7*fs: G1[3], G2[3], gpS with GP = gpS * Sf (in every fvsc scheme GP is parallel to the face area vector)
    int flagsUniform;        // >= 0: every active internal face carries these flags (the TMA kernel then skips the flags column)
    int lsqW;                // leastSquares: ELL width, cells [lsqW][nI] and coefficient vectors [lsqW][3][nI] (device face order)
    const int* lsqCells; const double* lsqCoef;
    const double* halfDist;  // nB
    const int* bKind;        // nB patch kind per boundary face
    const int* perm;         // nF device face -> polyMesh face (operator outputs are written in polyMesh order)
};

struct BndState {            // per boundary face
    RecA* A;                 // rho,U,e,p,T,H   (p = value used by interpolation, see quirk (h) in DESIGN.md)
    RecB* B;
    double* psi;
    double* pGrad;           // qgdFlux gradient
    double* pNew;            // p boundary value seen by fvsc::grad(p) inside the step
    double* phiw;            // phiwStar
    const int* bcU; const int* bcT; const int* bcP;     // kinds per boundary face
    const double* bvU; const double* bvT; const double* bvP;
    double* tauOutB;         // nB or null: boundary tauQGD as reported by models 1n and 2
};

struct StepScalars {         // device-resident time-step control
    double dt;
    double time;
    double coNum;            // last Courant number
    unsigned long long coMaxBits;   // max over faces of max(|Un+c|,|Un-c|)/h   (bit pattern of a non-negative double)
    unsigned long long tauMinBits;  // min over faces of tauQGDf
    double maxCo, maxDeltaT, cTau;
    int adjust;
    // QGDFoam.C:142-147: `if (min(e) <= 0 || min(rho) <= 0) { U.write(); e.write(); rho.write(); }` - the device records the
    // first step (1-based, counted by k_dt) whose update produced a non-positive e or rho; 0 = never
    int stepIndex, guardStep;
};

// ---- generic fvsc operator kernels (operator-level API)
void launchPointGather(cudaStream_t st, int K, const qgd_mesh& m, const double* cell, const double* bnd, double* pts);
void launchFvscGrad(cudaStream_t st, int K, const FaceView& fv, const double* cell, const double* pts,
                    const double* bnd, const double* bsg, const double* nbr, double* out);
void launchFvscDiv(cudaStream_t st, int K, const FaceView& fv, const double* cell, const double* pts,
                   const double* bnd, const double* bsg, const double* nbr, double* out);

// ---- QGDFoam step kernels
struct SolverView {
    int nCells, nPoints, nPatchPoints;
    int nOwned;                  // cells updated by this rank; [nOwned, nCells) are halo copies filled by the exchange
    // cell state, SoA: field k of cell c at S[k*nCells + c]; fields 0-7 = RecA (rho,Ux,Uy,Uz,e,p,T,H),
    // 8-15 = RecB (rhoUx,rhoUy,rhoUz,rhoE,c,mu,alphaEff,aByC).  SoA keeps the owner-ordered gathers of the face
    // kernel and the streaming point/cell kernels at one or two 128-B L1 wavefronts per warp-level load.
    double* S;
    double* P;                   // point values, SoA: field k (rho,Ux,Uy,Uz,e,p) of point i at P[k*nPoints + i]
    int pcEllW; const int* pcEll; const double* pcEllWt; const int* pcCount;   // ELL point -> cells [W][nPoints]
    const int* pcTailOff; const int* pcTailCell; const double* pcTailW;        // CSR tail for rows longer than W
    const int* patchPoints; const int* ppOff; const int* ppFace; const double* ppW;
    int cfEllW; const int* cfEll;                                              // ELL cell -> faces [W][nCells], -1 pad
    const int* cfTailOff; const int* cfTailEnc;
    const double* V; const double* hQGD; const double* aQGD;
    // multi-GPU overlap: points whose cells are all owned (gathered while the halo exchange is in flight) and the rest;
    // null = one launch over all points
    const int* ptsInterior; int nPtsInterior; const int* ptsHalo; int nPtsHalo;
    double* tauOut;              // nCells or null: tauQGD as the model reports it (models 1n and 2), for qgd_qgdfoam_get
    const double* su;            // [5][nCells] or null: explicit sources rhoSu, rhoUSu (3), rhoESu (volume-integrated; qgd_qgdfoam_set_sources)
    double* scVar;               // nCells or null: ScQGD of varScModel6/7, written by k_varsc before the cell thermo
    const unsigned char* scConst;// nCells or null: varScModel7 constScCellSet mask
    // face fluxes, 5 doubles per face (k = Fm, FUx, FUy, FUz, FE), SoA:
    //   internal face f : FI[k][slot(f)]            boundary face b : FB[k][b]
    // two-kernel form : one [5][nF] array, FI[k] = F + k*nF, FB[k] = FI[k] + nI, slot = f
    // pipelined form  : FI[k] = ring + k*ringSize, slot = f % ringSize (L2-resident ring) ; FB[k] = bnd + k*nB
    double* FI[5];
    double* FB[5];
    int ringSize;
    StepScalars* sc;
};

// scratch of the implicit-diffusion branch (QGDUEqn.H:54-75, QGDEEqn.H:53-64)
struct ImplicitView {
    double* GU0;      // [9][nCells] fvc::grad(U) of the old U (Gauss linear)
    double* GU1;      // [9][nCells] fvc::grad(U) of the solved U
    double* old;      // [4][nCells] U old (3), rho old
    double* FT;       // [3][nF] phiTauMC
    double* aU;       // [nF] muf |Sf| nonOrthDeltaCoeffs   (laplacian(muf,U) coefficient)
    double* aE;       // [nF] alphauf |Sf| nonOrthDeltaCoeffs
    double* Fs;       // [nF] phiSigmaDotU
    double* diagU; double* bU;      // [nCells], [3][nCells]
    double* diagE; double* bE;      // [nCells], [nCells]
};

// work plan of the pipelined face+cell kernel (k_face_cell_pipeline)
struct PipeView {
    int nChunks, lag, epoch;
    const int* cellOff;      // nChunks+1  owned cells of chunk k
    const int* faceOff;      // nChunks+1  device faces owned by chunk k
    const int* depOff;       // nChunks+1  C(k) needs F(depList[depOff[k] .. depOff[k+1]))  (owner chunks of its cells' faces)
    const int* depList;
    const int* ringOff;      // nChunks+1  F(k) rewrites ring slots last read by C(ringList[ringOff[k] .. ringOff[k+1]))
    const int* ringList;
    int* doneF; int* doneC;  // completion flags, epoch-valued
    int* queue;              // work-item counter, reset by k_dt
};

// wedge: the mesh has wedge patches (the wedge velocity condition is applied to the closed boundary state, k_wedge_bnd)
void launchInit(cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
                const double* U0, const double* T0, const double* p0, bool wedge = false);
// one QGDFoam.C:90-163 loop body; returns number of kernel launches issued
// ev (optional): 6 events recorded around k_points, k_face_flux, k_cell_update (begin/end pairs)
// hooks (optional, multi-GPU): midStep runs after the qgdFlux re-evaluation of p_b (exchange of halo p_b),
// beforeDt runs before the time-step kernel (all-reduce of the Courant max / tau min)
// waitHalo: called right before the first kernel that reads halo copies (the packed exchange of the previous step may
// still be in flight on the communication stream while the interior points are gathered)
// afterGrad (implicit branch, multi-GPU): runs after each cell-centred Gauss gradient of U with the gradient array, to fetch the
// gradients of the face-neighbour halo cells from their owners
// beforeBndPost: runs right before k_bnd_post overwrites the boundary state (varScModel5 keeps the p_b of that moment)
struct StepHooks { std::function<void()> midStep, beforeDt; std::function<void(cudaStream_t)> waitHalo; std::function<void(double*)> afterGrad;
                   std::function<void()> beforeBndPost; };
// Boundary work forked onto a second, high-priority stream: k_patch_points and k_bnd_flux of step n (and, on one GPU,
// k_bnd_post of step n-1) depend on the cell update only, not on the point gather, so they run beside k_points
// instead of in front of / behind the face kernel.  Joins: face kernel <- evPatch, k_dt <- evBndFlux, side <- evCell.
// The caller makes `side` wait for the main stream before the first step and joins evBndPost after the last one.
struct StepFork {
    cudaStream_t side;
    cudaEvent_t evEntry, evPatch, evBndFlux, evCell, evBndPost;
    bool postOnSide;         // k_bnd_post on the side stream too (single GPU: nothing on the main stream needs it before the next step)
};
// vertices of constraint patches (wedge, symmetryPlane): after the patch-point kernel their velocity is multiplied by the vertex's
// constraint tensor (pointConstraints [OF-v2312]); nrm = 9 doubles per listed vertex
struct WedgeView { int n; const int* pts; const double* nrm; };
int launchStep(cudaStream_t st, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
               bool anyQgdFlux, int gridFaces, bool adjust, cudaEvent_t* ev = nullptr, const StepHooks* hooks = nullptr,
               const PipeView* pipe = nullptr, int gridPipe = 0, const StepFork* fork = nullptr, const WedgeView* wedge = nullptr);
// implicit-diffusion step, phase by phase (the PCG solves run between the phases, see runStepsImplicit in qgd_abi.cu)
int launchImplicitPhase(cudaStream_t st, int phase, const Consts& c, const FaceView& fv, const SolverView& sv, const BndState& bs,
                        const ImplicitView& iv, bool anyQgdFlux, int gridFaces, bool adjust, const StepHooks* hooks = nullptr);
int faceKernelGrid();
void warmFaceKernel(bool adjust);   // sets the TMA face kernel's attributes / grid once (must not happen inside a stream capture)
int pipelineKernelGrid(int cfEllW);
void setFaceL2Hint(int bits); // env QGD_FACE_L2HINT: bit 0 = streamed constants / fluxes evict_first, bit 1 = state gathers evict_last
void setFaceTma(int on);      // env QGD_FACE_TMA: TMA-staged face kernel (default) vs register-prefetch kernel
const char* faceKernelName(const FaceView& fv);   // the internal-face kernel launchStep selects for this mesh
int faceKernelL2Hint();

} // namespace qgd

// one fvsc::fvscStencil instance (fvscStencil.H:46-137): face gradient records of a scheme on a mesh
struct qgd_fvsc {
    qgd_mesh* mesh = nullptr;
    bool reduced = false;
    qgd::DevBuf<int4> vtx;
    qgd::DevBuf<int> flags;
    qgd::DevBuf<double> G, halfDist;
    int flagsUniform = -1;
    bool lsq = false;              // leastSquares scheme
    int lsqW = 0;
    qgd::DevBuf<int> lsqCells;
    qgd::DevBuf<double> lsqCoef;
    // staging for operator-level calls (grown on demand)
    qgd::DevBuf<double> dCell, dBnd, dBsg, dNbr, dPts, dOut;
    qgd::FaceView view() const
    {
        const qgd_mesh& m = *mesh;
        qgd::FaceView v;
        v.nI = m.h.nInternal; v.nF = m.h.nFaces; v.nB = m.h.nBnd; v.fs = m.faceStride;
        v.nIActive = m.nIActive;
        v.zeroDivCmpt = -1;
        if (m.h.nD == 2 && !reduced) for (int d = 0; d < 3; ++d) if (m.h.gD[d] < 1) v.zeroDivCmpt = d;
        v.own = m.owner.p; v.nei = m.neighbour.p; v.vtx = vtx.p; v.flags = flags.p; v.Sf = m.Sf.p; v.magSf = m.magSf.p;
        v.w = m.w.p; v.hf = m.hQGDf.p; v.dC = m.dC.p; v.ndC = m.ndC.p; v.lsqW = lsq ? lsqW : 0; v.lsqCells = lsqCells.p; v.lsqCoef = lsqCoef.p; v.G = G.p; v.halfDist = halfDist.p; v.bKind = m.bfaceKind.p;
        v.perm = m.facePermDev.p; v.flagsUniform = flagsUniform;
        return v;
    }
};

namespace qgd {
// fvsc.C:47-85 (fvscOpName checks) + fvscStencil.C:59-95 (New); throws qgd::Error with the reference's messages
void fvscBuild(qgd_fvsc& op, qgd_mesh* mesh, const std::string& name);
}
