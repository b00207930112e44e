// ============================================================================
// CPU ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// Field-at-a-time CPU restatement of the QGDsolver hot path (fvsc face-centre
// derivatives, QGDCoeffs, QGDThermo state, QGDFoam explicit step, QHDFoam step
// with PCG).  Each function cites the reference listing it follows
// (Name.C:N == /root/reference/docs/html/Name_8C_source.html, source line N).
//
// "Parity unpinned": the reference ships no tests, golden vectors or tutorials
// for this path and cannot be compiled here (needs OpenFOAM v2312; SURVEY.md
// 8c), so this restatement is checked only against manufactured known-answer
// tests (tests/test_oracle_kat.py), not against reference outputs.
// OpenFOAM-internal semantics are isolated in functions tagged [OF-v2312].
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this.  The product (qgdsolver_b200/) never does.
// ============================================================================
#pragma once
#include <cstdint>

extern "C" {

// patch kinds (same numbering as include/qgd_b200.h)
enum { OR_PATCH_GENERIC = 0, OR_PATCH_EMPTY = 1, OR_PATCH_PROCESSOR = 2, OR_PATCH_WEDGE = 3,
       OR_PATCH_SYMMETRY_PLANE = 4 };   // polyPatch type symmetryPlane: an ordinary patch for the face derivatives, a constraint patch for the vertices
// boundary-condition kinds per patch and field
enum { OR_BC_FIXED_VALUE = 0, OR_BC_ZERO_GRADIENT = 1, OR_BC_FIXED_GRADIENT = 2, OR_BC_QGD_FLUX = 3,
       OR_BC_CALCULATED = 4, OR_BC_QHD_FLUX = 5, OR_BC_SLIP = 6,
       OR_BC_WEDGE = 7 };   // wedge [OF-v2312 wedgeFvPatchField]: U_b = faceT . U_P on a wedge patch (scalars there: zeroGradient)
// fvsc schemes
enum { OR_FVSC_GAUSSVOLPOINT = 0, OR_FVSC_REDUCED = 1, OR_FVSC_LEASTSQUARES = 2, OR_FVSC_LEASTSQUARESOPT = 3 };

typedef struct {
    int nCells, nFaces, nInternal, nPoints, nPatches;
    const double* points;        // nPoints*3
    const int* faceOff;          // nFaces+1
    const int* faceVerts;
    const int* owner;            // nFaces
    const int* neighbour;        // nInternal
    const int* patchStart;       // nPatches
    const int* patchSize;
    const int* patchKind;
    const double* C;             // nCells*3
    const double* V;             // nCells
    const double* Cf;            // nFaces*3
    const double* Sf;            // nFaces*3
    const double* magSf;         // nFaces
    const double* weights;       // nFaces
    const double* deltaCoeffs;   // nFaces
    const double* nonOrthDeltaCoeffs; // nFaces
    const double* neighbCellCentres;  // nBnd*3 (processor patches)
    int geometricD[3];
} or_mesh_t;

typedef struct {
    double R;        // specific gas constant  [perfectGas]
    double Cp;       // hConst
    double Hf;
    double Tref;     // hConst Tref / Hsref offsets [OF-v2312]
    double Hsref;
    double mu;       // constTransport
    double Pr;
    double ScQGD;    // constScPrModel1
    double PrQGD;
    int implicitDiffusion;     // QGD::implicitDiffusion (QGDThermo.C:61)
    int alphaEffGammaFactor;   // heThermo::alphaEff multiplies by gamma for internal energy [OF-v2312]
    int energyDdtRhoEQuirk;    // 1: QGDEEqn.H:67-72 as in the doc snapshot, fvm::ddt(rho,e) - fvc::ddt(rhoE)
                               // 0: fvm::ddt(rho,e) - fvc::ddt(rho,e)  (e keeps rhoE/rho - K)
    int qgdModel;              // 0 constScPrModel1, 1 constScPrModel1n, 2 constScPrModel2, 5 varScModel5, 6 varScModel6, 7 varScModel7
    // implicitDiffusion branch (QGDUEqn.H:54-75, QGDEEqn.H:53-64): fvSolution controls of the U and e solvers (PCG)
    double diffTol, diffRelTol;
    int diffMaxIter, diffPrecond;
    // varScModel7 dictionary entries (varScModel7.C:96-119): cSc1 (default 1), minSc / maxSc (default -1 = off)
    double varScCSc1, varScMinSc, varScMaxSc;
    // transport model (psiQGDThermos.C:65-111): 0 const (mu, Pr above), 1 powerLaw  mu = mu0 (T/T0)^k, alphah = mu * (1/Pr)
    // (powerLawTransportI.H:120-150), 2 sutherland  mu = As sqrt(T)/(1 + Ts/T), alphah = mu Cv (1.32 + 1.77 R/Cv)/Cp [OF-v2312
    // sutherlandTransportI.H]
    int transportModel;
    double mu0, T0, kExp;
    double As, Ts;
    // thermo model: 0 hConst (Cp, Hf, Tref, Hsref above), 1 eConst  Es = Cv (T - Tref) + Esref, Cp = Cv + R [OF-v2312 eConstThermoI.H]
    int thermoModel;
    double Cv, Esref;
    // varScModel5 dictionary entries (varScModel5.C:61-110; qgdModel 5): smoothCoeff (0.1), rC (0.5), badQualitySc (0.05),
    // maxAspectRatio (1.5); its minSc (0.05) / maxSc (1.0) travel in varScMinSc / varScMaxSc above
    double varSc5SmoothCoeff, varSc5RC, varSc5BadQualitySc, varSc5MaxAspectRatio;
} or_qgd_params_t;

typedef struct or_ctx or_ctx;

or_ctx* or_create(const or_mesh_t* mesh, int nThreads);
void    or_destroy(or_ctx*);

// ---- derived mesh data (for KATs / parity of setup kernels)
void or_get_hQGDf(or_ctx*, double* out /*nFaces*/);
void or_get_hQGD(or_ctx*, double* out /*nCells*/);

// ---- fvsc operators.  cell: nCells*k ; bnd: nBnd*k boundary values ; bndSnGrad: nBnd*k patch snGrad
//      nbr: nBnd*k patchNeighbourField (processor patches only, may be NULL) ; out: nFaces*(3k | k/3)
void or_fvsc_grad(or_ctx*, int scheme, int ncmpt /*1|3*/, const double* cell, const double* bnd,
                  const double* bndSnGrad, const double* nbr, double* out);
void or_fvsc_div(or_ctx*, int scheme, int ncmpt /*3|9*/, const double* cell, const double* bnd,
                 const double* bndSnGrad, const double* nbr, double* out);
void or_vol_point_interpolate(or_ctx*, int ncmpt, const double* cell, const double* bnd, double* outPoints);
void or_linear_interpolate(or_ctx*, int ncmpt, const double* cell, const double* bnd, double* out);

// ---- QGDFoam
//  bc kinds per patch for U, T, p ; fixed values per boundary face (U: nBnd*3, T,p: nBnd)
// varScModel7 "constScCellSet" (varScModel7.C:143-158,246-254): cells whose ScQGD is reset to the dictionary ScQGD each step.
// Call before or_qgd_init.
void or_qgd_set_const_sc_cells(or_ctx*, const int* cells, int n);
// [OF-v2312] fvc::smooth(field, coeff) (fvcSmooth/smooth.C: FaceCellWave<smoothData>) on a cell field, as varScModel5.C:232
// applies it to ScQGD; returns the number of FaceCellWave iterations.  Serial meshes only (no coupled patches).
int or_fvc_smooth(or_ctx*, double* field /*nCells, in/out*/, double coeff);
// varScModel5.C:112-132: the per-cell quality floor cqSc from primitiveMeshTools::cellClosedness [OF-v2312]; aspectRatio may be NULL
void or_varsc5_cell_quality(or_ctx*, double badQualitySc, double maxAspectRatio, double* cqSc /*nCells*/, double* aspectRatio);
// Explicit source matrices of QGDRhoEqn.H:46 / QGDUEqn.H:62,85 / QGDEEqn.H:60,71 (rhoSu, rhoUSu, rhoESu; zero in QGDFoam,
// createZeroSources.H:28-44; the Lagrangian cloud's Srho/SU/Sh in particlesQGDFoam): the volume-integrated explicit
// source of each cell, i.e. minus the fvMatrix::source() of the matrix on the right-hand side.  suRho, suE: nCells,
// suU: nCells*3; NULL = zero.  They stay in force until changed.
void or_qgd_set_sources(or_ctx*, const double* suRho, const double* suU, const double* suE);
void or_qgd_init(or_ctx*, const or_qgd_params_t*, int fvscScheme,
                 const int* bcU, const int* bcT, const int* bcP,
                 const double* bvU, const double* bvT, const double* bvP,
                 const double* U0, const double* T0, const double* p0, const double* alphaQGD /*nCells or NULL*/,
                 double deltaT0);
//  nSteps explicit steps.  adjustTimeStep: 0 fixed dt.  Returns last Courant number (or -1).
double or_qgd_step(or_ctx*, int nSteps, int adjustTimeStep, double maxCo, double maxDeltaT, double cTau);
double or_qgd_deltaT(or_ctx*);
//  field ids: 0 rho, 1 rhoU(3), 2 rhoE, 3 U(3), 4 e, 5 p, 6 T, 7 c, 8 mu, 9 alpha, 10 tauQGD
//  cells: nCells*k ; bnd (may be NULL): nBnd*k
void or_qgd_get(or_ctx*, int field, double* cells, double* bnd);
//  face fields of the last step: 0 phiJm, 1 phiJmU(3), 2 phiP(3), 3 phiPi(3), 4 phiJmH, 5 phiQ, 6 phiPiU,
//  7 tauQGDf, 8 gradUf(9), 9 gradef(3), 10 gradRhof(3), 11 gradPf(3), 12 phiwStar
void or_qgd_get_face(or_ctx*, int field, double* out /*nFaces*k*/);

// ---- QHDFoam (explicit branch).  heRhoQGDThermo with rhoConst + hConst + constTransport (+ beta), laminar.
typedef struct {
    double rho0;               // rhoConst
    double mu, Pr, beta;       // constTransport ; beta from mixture.transport (QHDFoam/createFields.H:110-115)
    double g[3];               // constant/gravitationalProperties
    int qgdModel;              // 0 constTau, 1 H2bynuQHD, 2 HbyUQHD, 3 T0byGr
    double Tau, UQHD, Gr, T0;  // model coefficients
    int implicitDiffusion;     // QGD::implicitDiffusion
    double pTol, pRelTol;      // fvSolution::solvers::p
    int pMaxIter, pPrecond;    // precond: 0 none, 1 diagonal, 2 DIC
    int pRefCell; double pRefValue;   // setRefCell(p, thermo.subDict("QGD"), ...)
    // implicitDiffusion branch (QHDUEqn.H:46-65, QHDTEqn.H:69-80): fvSolution controls of the U and T solvers (PCG)
    double diffTol, diffRelTol;
    int diffMaxIter, diffPrecond;
    // 1: the scalarTransportQHDFoam loop (scalarTransportQHDFoam.C:70-135): U and tau frozen, no pressure or momentum
    // equation, phi = phiu, T equation (implicitDiffusion only) with the extra -fvc::Sp(fvc::div(phiu),T) term,
    // Courant number from mag(Uf)/hQGDf
    int scalarTransport;
} or_qhd_params_t;
//  bc kinds per patch: fixedValue | zeroGradient | fixedGradient (qhdFlux behaves as fixedGradient in QHDFoam, see
//  DESIGN.md quirk (i)); bv*: value on fixedValue faces, gradient on fixedGradient faces
void or_qhd_init(or_ctx*, const or_qhd_params_t*, int fvscScheme, const int* bcU, const int* bcT, const int* bcP,
                 const double* bvU, const double* bvT, const double* bvP, const double* U0, const double* T0,
                 const double* p0, const double* alphaQGD, double deltaT0);
double or_qhd_step(or_ctx*, int nSteps, int adjustTimeStep, double maxCo, double maxDeltaT, double cTau);
double or_qhd_deltaT(or_ctx*);
//  cell fields: 0 U(3), 1 T, 2 p, 3 tauQGD ; face fields: 0 phi, 1 phiu, 2 phiwo, 3 tauQGDf, 4 gradPf(3), 5 gradUf(9), 6 gradTf(3)
void or_qhd_get(or_ctx*, int field, double* cells, double* bnd);
void or_qhd_get_face(or_ctx*, int field, double* out);
void or_qhd_solver_info(or_ctx*, int* iters, double* res0, double* res);

// ---- LDU PCG [OF-v2312 PCG + DIC / diagonal]  (QHDpEqn.H:45)
//  symmetric: lower == upper.  precond: 0 none, 1 diagonal, 2 DIC.  returns iterations.
int or_pcg_solve(or_ctx*, const double* diag, const double* upper, const double* b, double* x,
                 double tol, double relTol, int maxIter, int precond, double* initRes, double* finalRes);
// the decomposed-run form: global matrix, global reductions, preconditioner local to each processor block
// (cellBlock[c] = processor of cell c; NULL = serial)
int or_pcg_solve_blocks(or_ctx*, const double* diag, const double* upper, const double* b, double* x,
                        double tol, double relTol, int maxIter, int precond, double* initRes, double* finalRes,
                        const int* cellBlock);

// faceSet "degenerateStencilFaces" (leastSquaresStencil.C:63-132): internal faces whose leastSquares gradient is replaced by
// nf*snGrad, in addition to the faces found degenerate by det(G) < 1.  polyMesh face ids; boundary faces are ignored here
// (the reference only keeps those on processor patches).
void or_set_degenerate_faces(or_ctx*, const int* faces, int n);
// every PCG solve of the following QGDFoam (implicit branch) / QHDFoam steps runs in the decomposed-run form above
// (NULL: back to serial).  The rest of the step is decomposition-independent in exact arithmetic.
void or_set_pcg_blocks(or_ctx*, const int* cellBlock);

} // extern "C"
