/* ============================================================================
 * qgd_b200.h — C ABI of the B200-native QGD/QHD hot path (libqgd_b200.so)
 *
 * The reference (unicfdlab/QGDsolver) has no FFI: its boundary is OpenFOAM
 * runTimeSelection + objectRegistry in-process C++.  This header
 * is what an OpenFOAM-side shim library binds instead; every entry point cites
 * the reference interface it replaces as  File.C:line  ==
 * /root/reference/docs/html/File_8C_source.html, original source line.
 *
 * Rules: plain pointers and sizes only; the caller owns all host arrays; the
 * library copies on *_create / *_set and owns device memory behind opaque
 * handles; every call returns 0 on success or a negative qgd_status, and
 * qgd_last_error() returns the message the shim forwards to
 * FatalErrorInFunction (the reference aborts with FatalError at
 * fvscStencil.C:72-78, QGDCoeffs.C:72-78, fvsc.C:62).  Nothing here ever falls
 * back to a CPU path: without a usable CUDA device every compute call fails.
 * One host thread per handle; not thread-safe per handle (fvscStencil.C:57).
 * ==========================================================================*/
#ifndef QGD_B200_H
#define QGD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QGD_OK = 0,
    QGD_ERR_INVALID = -1,        /* bad argument / inconsistent mesh            */
    QGD_ERR_UNKNOWN_MODEL = -2,  /* runTimeSelection lookup failed               */
    QGD_ERR_UNSUPPORTED = -3,    /* recognised but not implemented on device     */
    QGD_ERR_CUDA = -4,           /* CUDA runtime / no device                     */
    QGD_ERR_COMM = -5,           /* NCCL                                         */
    QGD_ERR_STATE = -6           /* call order (e.g. step before init_fields)    */
} qgd_status;

/* polyPatch kinds the path distinguishes (emptyFvPatch, wedgeFvPatch,
 * processorFvPatch tests at QGDCoeffs.C:346-348, GaussVolPointBase2D.C:175-199,
 * GaussVolPointBase3D.C:76-79) */
typedef enum { QGD_PATCH_GENERIC = 0, QGD_PATCH_EMPTY = 1, QGD_PATCH_PROCESSOR = 2, QGD_PATCH_WEDGE = 3,
               QGD_PATCH_SYMMETRY_PLANE = 4   /* polyPatch type symmetryPlane: an ordinary patch for the face derivatives and length
                                                 scales; its vertices are constrained like those of wedge patches [OF-v2312
                                                 pointConstraints] and leastSquares leaves its faces at zero
                                                 (extendedFaceStencilScalarGrad.C:86-109).  Field condition: QGD_BC_SLIP. */
} qgd_patch_kind;

/* boundary-condition kinds of the closed device-native set */
typedef enum {
    QGD_BC_FIXED_VALUE = 0,      /* fixedValue                                   */
    QGD_BC_ZERO_GRADIENT = 1,    /* zeroGradient                                 */
    QGD_BC_FIXED_GRADIENT = 2,   /* fixedGradient                                */
    QGD_BC_QGD_FLUX = 3,         /* qgdFlux  qgdFluxFvPatchScalarField.C:159-208 */
    QGD_BC_CALCULATED = 4,       /* calculated                                   */
    QGD_BC_QHD_FLUX = 5,         /* qhdFlux  qhdFluxFvPatchScalarField.C:159-219 */
    QGD_BC_SLIP = 6,             /* slip | symmetryPlane | symmetry for U [OF-v2312 basicSymmetryFvPatchField]:
                                    U_b = U_P - n (n . U_P); scalars on such patches are zeroGradient.  QGDFoam, explicit
                                    branch (the Mach-3 forward-facing-step walls, BASELINE configs[1])              */
    QGD_BC_WEDGE = 7             /* wedge for U on a QGD_PATCH_WEDGE patch [OF-v2312 wedgeFvPatchField::evaluate]:
                                    U_b = faceT . U_P with faceT = rotationTensor(centre-plane normal, patch normal)
                                    (wedgePolyPatch); scalars on wedge patches are zeroGradient.  The face derivatives
                                    stay zero there (GaussVolPointBase2D.C:175-179) and the vertices of wedge patches
                                    lose the patch-normal component of interpolated vectors / tensors [OF-v2312
                                    pointConstraints].  QGDFoam, explicit branch.                                  */
} qgd_bc_kind;

typedef struct qgd_mesh qgd_mesh;       /* fvMesh image on the device             */
typedef struct qgd_fvsc qgd_fvsc;       /* one fvsc::fvscStencil instance         */
typedef struct qgd_solver qgd_solver;   /* QGDFoam time loop state                */
typedef struct qgd_qhd_solver qgd_qhd_solver;   /* QHDFoam time loop state        */

/* ---- library / device ------------------------------------------------------ */
int         qgd_init(int device);                 /* cudaSetDevice + stream; call once per process (rank) */
const char* qgd_last_error(void);                 /* message of the last failing call on this thread     */
int         qgd_version(void);
int         qgd_device_synchronize(void);

/* ---- mesh: what fvMesh hands to the reference ------------------------------ */
typedef struct {
    int n_cells, n_faces, n_internal_faces, n_points, n_patches;
    const double* points;              /* n_points*3            polyMesh::points()                 */
    const int*    face_offsets;        /* n_faces+1             faceList as CSR                    */
    const int*    face_verts;
    const int*    owner;               /* n_faces                                                   */
    const int*    neighbour;           /* n_internal_faces                                          */
    const int*    patch_start;         /* n_patches             polyPatch::start()                  */
    const int*    patch_size;
    const int*    patch_kind;          /* qgd_patch_kind                                            */
    const double* C;                   /* n_cells*3             fvMesh::C()                         */
    const double* V;                   /* n_cells               fvMesh::V()                         */
    const double* Cf;                  /* n_faces*3             fvMesh::Cf() incl. boundary         */
    const double* Sf;                  /* n_faces*3             fvMesh::Sf()                        */
    const double* magSf;               /* n_faces                                                   */
    const double* weights;             /* n_faces               surfaceInterpolation::weights()     */
    const double* deltaCoeffs;         /* n_faces                                                   */
    const double* nonOrthDeltaCoeffs;  /* n_faces                                                   */
    const double* neighb_cell_centres; /* n_bnd*3, processor patches: neighbFaceCellCentres() or NULL */
    int geometric_d[3];                /* fvMesh::geometricD()                                      */
    /* multi-GPU extended sub-mesh (qgdsolver_b200/decompose.py): cells [0,n_owned_cells) are owned by this
     * rank, the rest are halo copies; internal faces flagged here join an owned and a halo cell and follow the
     * reference's processor-patch rules (hQGDf = 1/deltaCoeffs, QGDCoeffs.C:195-199,310-317).  0 / NULL = serial. */
    int n_owned_cells;
    const int* coupled_internal_face;  /* n_internal_faces flags or NULL                             */
} qgd_mesh_desc;

int qgd_mesh_create(const qgd_mesh_desc* desc, qgd_mesh** out);
/* faceSet "degenerateStencilFaces" (constant/polyMesh/sets, leastSquaresStencil.C:63-132): polyMesh ids of faces whose
 * `leastSquares` gradient is replaced by nf*snGrad, in addition to those found degenerate by det(G) < 1.  Call before
 * qgd_fvsc_create / qgd_*foam_create on this mesh; boundary faces in the list are ignored (serial meshes). */
int qgd_mesh_set_degenerate_stencil_faces(qgd_mesh* mesh, const int* faces, int n);
/* DIC blocks.  In `mpirun -np N` runs of the reference the DIC preconditioner of PCG (QHDpEqn.H:45; QGDUEqn.H:54-75,
 * QGDEEqn.H:53-64) is factorised and swept on each processor's own lduMatrix: faces between cells of different processors
 * are left out of it [OF-v2312 DICPreconditioner].  The device uses the same block-local form, with blocks small enough to
 * be swept by one CTA in shared memory (no grid-wide synchronisation inside the preconditioner); without blocks the
 * preconditioner is the serial one (exact operation order, level-scheduled: one grid synchronisation per level - only
 * usable on small meshes).  cell_block: block id >= 0 of every cell (halo cells of a sub-mesh are ignored), NULL clears.
 * make: recursive coordinate bisection of the owned cells into compact tiles of at most target_cells cells.
 * Call before qgd_pcg_solve / qgd_*foam_init_fields on this mesh.  get: -1 for halo cells or when no blocks are set. */
int qgd_mesh_set_pcg_blocks(qgd_mesh* mesh, const int* cell_block);
int qgd_mesh_make_pcg_blocks(qgd_mesh* mesh, int target_cells, int* n_blocks);
int qgd_mesh_get_pcg_blocks(qgd_mesh* mesh, int* cell_block);
int qgd_mesh_destroy(qgd_mesh* mesh);
/* derived fields, for write-back / parity checks.  what: 0 hQGDf (n_faces)  QGDCoeffs.C:298-318
 *                                                        1 hQGD  (n_cells)  QGDCoeffs.C:320-362 */
int qgd_mesh_get(qgd_mesh* mesh, int what, double* out);

/* ---- fvsc face-centre derivative operators (operator-level integration) ----
 * qgd_fvsc_create  replaces fvscStencil::New / lookupOrNew (fvscStencil.C:59-118); scheme_name is the
 *   fvSchemes::fvsc entry (fvsc.C:47-58): "GaussVolPoint" | "reduced" | "leastSquares" | "leastSquaresOpt" (the last
 *   two on 2D/1D meshes: they are rejected in 3D exactly like fvsc.C:60-63);
 *   anything else -> QGD_ERR_UNKNOWN_MODEL with the reference's "Unknown Model type" message.
 * qgd_fvsc_grad    replaces fvscStencil::Grad(volScalarField|volVectorField)  (fvscStencil.H:105-116,
 *   GaussVolPointStencil.C:71-99, reducedFaceNormalStencil.C:69-88).  ncmpt = 1 | 3.
 * qgd_fvsc_div     replaces fvscStencil::Div(volVectorField|volTensorField)   (fvscStencil.H:119-130,
 *   GaussVolPointStencil.C:101-129).  ncmpt = 3 | 9.
 * Array layout is OpenFOAM's (array of vectors/tensors):
 *   cell        n_cells*ncmpt   primitiveField()
 *   bnd         n_bnd*ncmpt     boundaryField() values, already evaluated (the shim calls
 *                               correctBoundaryConditions() itself, GaussVolPointStencil.C:73)
 *   bnd_sngrad  n_bnd*ncmpt     boundaryField()[patchi].snGrad()  (GaussVolPointBase3D.C:791)
 *   nbr         n_bnd*ncmpt     patchNeighbourField() on processor patches, or NULL (:785-786)
 *   out         n_faces*(3*ncmpt) for grad, n_faces*(ncmpt/3) for div; internal faces then boundary faces.
 * All pointers are HOST pointers; the call does H2D, kernels, D2H and returns when `out` is valid. */
int qgd_fvsc_create(qgd_mesh* mesh, const char* scheme_name, qgd_fvsc** out);
int qgd_fvsc_destroy(qgd_fvsc* op);
int qgd_fvsc_grad(qgd_fvsc* op, int ncmpt, const double* cell, const double* bnd, const double* bnd_sngrad,
                  const double* nbr, double* out);
int qgd_fvsc_div(qgd_fvsc* op, int ncmpt, const double* cell, const double* bnd, const double* bnd_sngrad,
                 const double* nbr, double* out);

/* ---- QGDFoam (solver-level integration) -------------------------------------
 * Dictionary content the reference reads: thermophysicalProperties (thermoType hePsiQGDThermo /
 * pureMixture / const / hConst / perfectGas / sensibleInternalEnergy, psiQGDThermos.C:65-111),
 * its QGD sub-dictionary (QGDThermo.C:48-82, QGDCoeffs.C:58-117, constScPrModel1.C:58-89),
 * fvSchemes::fvsc, controlDict (QGDCourantNo.H:36, setDeltaT-QGDQHD.H:41-58).
 * fvSchemes assumed by the fused face kernels: ddt Euler, grad Gauss linear, interpolationSchemes linear (or default none)
 * and divSchemes none | Gauss linear - the branches of qgdInterpolate / qgdFlux<T> (QGDInterpolate.H:38-118) that return
 * linearInterpolate(psi) and flux*psif, or fvc::interpolate / fvc::flux with the `linear` scheme, which give the same values.
 * Any other interpolation or convection scheme (upwind, limited ...) needs OpenFOAM's scheme objects and a registered flux
 * field; the shim must refuse such an fvSchemes (qgdsolver_b200/runcase.py does), it cannot be expressed through this ABI. */
typedef struct {
    const char* fvsc_scheme;        /* fvSchemes::fvsc::default                                     */
    const char* qgd_coeffs_model;   /* QGD::QGDCoeffs : "constScPrModel1" | "constScPrModel1n" | "constScPrModel2" |
                                       "varScModel5" | "varScModel6" | "varScModel7"                */
    double R;                       /* perfectGas: R = 8314.47/W                                    */
    double Cp, Hf, Tref, Hsref;     /* hConst                                                       */
    double mu, Pr;                  /* constTransport                                               */
    double ScQGD, PrQGD;            /* constScPrModel1.C:58-89 (default 1, 1)                       */
    int implicit_diffusion;         /* QGD::implicitDiffusion (QGDThermo.C:61: default true): QGDUEqn.H:54-75,
                                       QGDEEqn.H:53-64 with the PCG controls below                  */
    int alpha_eff_gamma_factor;     /* heThermo::alphaEff for internal energy [OF-v2312]            */
    int energy_ddt_rhoE_quirk;      /* 1 = QGDEEqn.H:67-72 literally (default), 0 = fvc::ddt(rho,e) */
    /* controlDict */
    int adjust_time_step;           /* QGDCourantNo.H:36                                            */
    double max_co, max_delta_t, c_tau;   /* readTimeControls.H ; setDeltaT-QGDQHD.H:45 (cTau 0.75)  */
    double delta_t;                 /* initial deltaT                                               */
    /* fvSolution::solvers::"(U|e)" for the implicit-diffusion branch: PCG + preconditioner                  */
    double diff_tolerance, diff_rel_tol;
    int diff_max_iter;
    const char* diff_preconditioner;   /* "DIC" | "diagonal" | "none" ; NULL = "DIC"                */
    /* QGD sub-dictionary entries of varScModel7 (varScModel7.C:96-119): cSc1 (default 1), minSc / maxSc
     * (default -1 = off).  Read only when qgd_coeffs_model is "varScModel7"; "varScModel6" has none.  */
    double varsc_cSc1, varsc_minSc, varsc_maxSc;
    /* thermoType instantiations of psiQGDThermos.C:65-111 (all pureMixture / perfectGas / sensibleInternalEnergy):
     *   transport_model  NULL | "const" (mu, Pr above) | "sutherland" (As, Ts [OF-v2312 sutherlandTransportI.H]) |
     *                    "powerLaw" (mu0, T0, k_exp, Pr; powerLawTransportI.H:120-150)
     *   thermo_model     NULL | "hConst" (Cp, Hf, Tref, Hsref above) | "eConst" (Cv, Hf, Tref, Esref; const transport only) */
    const char* transport_model;
    double As, Ts;
    double mu0, T0, k_exp;
    const char* thermo_model;
    double Cv, Esref;
    /* QGD sub-dictionary entries of varScModel5 (varScModel5.C:61-110), read only when qgd_coeffs_model is "varScModel5":
     * smoothCoeff (0.1), rC (0.5), badQualitySc (0.05), maxAspectRatio (1.5); its minSc (0.05) / maxSc (1.0) travel in
     * varsc_minSc / varsc_maxSc above.  The shim passes the dictionary value or the default given in brackets.
     * varScModel5 on the device: explicit and implicit-diffusion branch, serial mesh; ScQGD = rC |grad(psi p)| hQGD / (psi p) + (1 - rC) ScQGD, clamped,
     * floored by the cellClosedness aspect-ratio value, smoothed with fvc::smooth (FaceCellWave, reference visiting order). */
    double varsc5_smoothCoeff, varsc5_rC, varsc5_badQualitySc, varsc5_maxAspectRatio;
} qgd_qgdfoam_desc;

int qgd_qgdfoam_create(qgd_mesh* mesh, const qgd_qgdfoam_desc* desc, qgd_solver** out);
int qgd_qgdfoam_destroy(qgd_solver* s);
/* varScModel7 / varScModel5 "constScCellSet" (varScModel7.C:143-158,246-254; varScModel5.C:134-149,222-230): polyMesh cell
 * ids whose ScQGD is reset to the dictionary ScQGD after every sensor evaluation (model 5: before the smoothing).  Call
 * before qgd_qgdfoam_init_fields.  n = 0 clears the set. */
int qgd_qgdfoam_set_const_sc_cells(qgd_solver* s, const int* cells, int n);
/* Explicit source matrices of the conservative equations: rhoSu (QGDRhoEqn.H:46), rhoUSu (QGDUEqn.H:62,85), rhoESu
 * (QGDEEqn.H:60,71).  QGDFoam builds them as zero matrices (createZeroSources.H:28-44); particlesQGDFoam fills them
 * from the Lagrangian cloud.  Each array holds the volume-integrated explicit source of every cell (= minus the
 * fvMatrix::source() of the right-hand-side matrix; implicit Sp parts are not supported): rhoSu n_cells [kg/s],
 * rhoUSu n_cells*3 [N], rhoESu n_cells [W]; a NULL array is zero, all NULL removes the sources.  The values stay in
 * force for every following step until the next call.  As in the reference's explicit branch the momentum source
 * corrects U only, rhoU keeps the value of the conservative update (QGDUEqn.H:79-89). */
int qgd_qgdfoam_set_sources(qgd_solver* s, const double* rhoSu, const double* rhoUSu, const double* rhoESu);
/* boundary conditions of U, T, p per patch (0/U, 0/T, 0/p): kinds n_patches each (qgd_bc_kind),
 * fixed values per boundary face: U n_bnd*3, T n_bnd, p n_bnd (read only where the kind is fixedValue). */
int qgd_qgdfoam_set_bcs(qgd_solver* s, const int* bc_U, const int* bc_T, const int* bc_p,
                        const double* val_U, const double* val_T, const double* val_p);
/* initial internal fields as read from 0/ (QGDFoam createFields.H:3-109): U n_cells*3, T, p n_cells,
 * alphaQGD n_cells or NULL (QGDCoeffs.C:119-160: uniform 0.5).  Runs thermo construction on the device. */
int qgd_qgdfoam_init_fields(qgd_solver* s, const double* U, const double* T, const double* p, const double* alphaQGD);
/* n_steps passes of the QGDFoam.C:90-163 loop body, fully on the device (no host sync inside). */
int qgd_qgdfoam_step(qgd_solver* s, int n_steps);
/* Same loop, but with HOST state buffers (operator / "plug-in" style hand-off on every call): the full cell
 * state is uploaded before the steps and downloaded after them, inside the call.  This is the state the
 * reference keeps in its registered fields (createFields.H:10-74 + thermo:mu); boundary fields stay on the
 * device.  Buffers should be page-locked (cudaHostAlloc / cudaHostRegister) for full PCIe speed. */
typedef struct {
    double* rho;   /* n_cells   */
    double* U;     /* n_cells*3 */
    double* e;     /* n_cells   */
    double* p;     /* n_cells   */
    double* T;     /* n_cells   */
    double* rhoU;  /* n_cells*3 */
    double* rhoE;  /* n_cells   */
    double* mu;    /* n_cells   thermo:mu incl. muQGD (QGDThermo.C:91-98) */
} qgd_state_host;
#define QGD_STATE_DOUBLES_PER_CELL 12
int qgd_qgdfoam_step_host(qgd_solver* s, int n_steps, const qgd_state_host* in, const qgd_state_host* out);
/* Field-level hand-off with HOST buffers: the caller passes the fields the reference reads from a time directory
 * (QGDFoam createFields.H:3-109: U n_cells*3, T, p n_cells); the library re-creates the thermodynamic and conserved
 * state from them exactly as qgd_qgdfoam_init_fields does (boundary conditions and alphaQGD as set before), runs
 * n_steps of the loop and writes back what the reference writes at a write time: U, T, p and, where the pointer is
 * non-NULL, rho, rhoU, rhoE.  This is the reference's restart semantics (write + read of a time directory), i.e. 5
 * doubles per cell each way instead of the 12 of qgd_qgdfoam_step_host; `in` NULL keeps the device state, `out` NULL
 * skips the download.  Buffers should be page-locked. */
typedef struct {
    double* U;     /* n_cells*3 */
    double* T;     /* n_cells   */
    double* p;     /* n_cells   */
    double* rho;   /* n_cells   out only, may be NULL */
    double* rhoU;  /* n_cells*3 out only, may be NULL */
    double* rhoE;  /* n_cells   out only, may be NULL */
} qgd_fields_host;
int qgd_qgdfoam_step_fields_host(qgd_solver* s, int n_steps, const qgd_fields_host* in, const qgd_fields_host* out);
/* name of the internal-face kernel the step launches on this solver's mesh ("k_face_flux_tma" | "k_face_flux" |
 * "k_face_flux_lsq" | "k_face_cell_pipeline") and its L2 cache-policy bits (QGD_FACE_L2HINT); bench bookkeeping */
const char* qgd_qgdfoam_face_kernel(qgd_solver* s, int* l2hint);
/* fields: 0 rho, 1 rhoU(3), 2 rhoE, 3 U(3), 4 e, 5 p, 6 T, 7 c, 8 mu, 9 alpha, 10 tauQGD, 11 H, 12 ScQGD.
 * cells: n_cells*k host buffer, bnd: n_bnd*k host buffer or NULL. */
int qgd_qgdfoam_get(qgd_solver* s, int field, double* cells, double* bnd);
/* face flux of the last step: 0 phiJm, 1 momentum flux (phiJmU+phiP-phiPi, 3), 2 energy flux
 * (phiJmH+phiQ-phiPiU); n_faces*k host buffer */
int qgd_qgdfoam_get_flux(qgd_solver* s, int which, double* out);
int qgd_qgdfoam_get_scalars(qgd_solver* s, double* delta_t, double* courant, double* time);
/* The reference dumps U, e, rho when an update leaves min(e) <= 0 or min(rho) <= 0 and carries on (QGDFoam.C:142-147).  The
 * device records the first such step (1-based count of steps taken by this solver, 0 = never); the shim polls this after a
 * batch of steps and performs the writes (qgd_qgdfoam_get fields 3, 4, 0). */
int qgd_qgdfoam_state_guard(qgd_solver* s, int* first_step);
/* Step form.  mode 0 (default): two kernels (faces, then cells) with a full-size flux array; always used with
 * adjustTimeStep, whose global Courant maximum must be known before any cell is updated; needed for
 * qgd_qgdfoam_get_flux.  mode 1 (fixed deltaT only): internal faces and cells are processed by ONE persistent kernel whose
 * work queue interleaves face chunks and cell chunks; the 5 flux doubles per face live in an L2-resident ring instead of
 * an HBM array (DESIGN.md "flux ring").  chunk_cells / lag / ring_slots: 0 / -1 / 0 select the defaults.  Results are
 * bit-identical in both forms. */
int qgd_qgdfoam_set_pipeline(qgd_solver* s, int mode, int chunk_cells, int lag, int ring_slots);
int qgd_qgdfoam_get_pipeline(qgd_solver* s, int* mode, int* chunk_cells, int* lag, int* ring_slots, int* n_chunks, int* grid);
/* Opt-in, environment QGD_STEP_GRAPH=1: qgd_qgdfoam_step captures one step (all its kernels on both streams) as a CUDA graph
 * and replays it n_steps times; the adaptive time step stays on the device, so no re-capture is needed.  Single GPU, two-kernel
 * step form, explicit branch, not with varScModel5 (its smoothing loop is host-driven) or per-kernel profiling; other
 * configurations silently use the stream launches.  Results are bit-identical.  Returns the steps replayed from a graph so far. */
long long qgd_qgdfoam_graph_steps(qgd_solver* s);
/* implicit-diffusion branch: PCG iterations of the last Ux, Uy, Uz and e solves */
int qgd_qgdfoam_diffusion_iterations(qgd_solver* s, int iters[4]);
/* kernel launches issued by this solver so far (bench bookkeeping) */
long long qgd_qgdfoam_launch_count(qgd_solver* s);
/* per-kernel CUDA-event timing on the solver stream: enable, run steps, then read the summed durations (ms) of the
 * point-gather, face-flux and cell-update kernels over the profiled steps */
int qgd_qgdfoam_profile(qgd_solver* s, int enable);
int qgd_qgdfoam_kernel_times(qgd_solver* s, double* ms_points, double* ms_face, double* ms_cell, int* n_steps);
/* CUDA-event timing of the device step loop: call begin, steps, end -> milliseconds on the solver stream */
int qgd_timer_begin(void);
int qgd_timer_end(float* ms);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink replaces Pstream ------------------
 * qgd_comm_unique_id: rank 0 creates the 128-byte NCCL id, the launcher broadcasts it (torch.distributed / MPI).
 * qgd_comm_init: every rank, after qgd_init.  qgd_qgdfoam_set_halo: exchange lists of this rank's extended sub-mesh:
 * for neighbour k (rank nbr_rank[k]) cells  send_cells[send_cell_off[k] .. send_cell_off[k+1])  are packed and sent,
 * and the received block is scattered to recv_cells[...]; same for the boundary-face state of physical boundary
 * faces of halo cells (boundary-face indices = face - n_internal_faces).  Order on both sides: ascending global id.
 * After set_halo every qgd_qgdfoam_step does: cell update -> ONE packed exchange (16 doubles/cell, 20/boundary face)
 * -> next step; with adjust_time_step the Courant max / tau min are all-reduced on the device before k_dt. */
int qgd_comm_unique_id(void* out128);
int qgd_comm_init(int rank, int n_ranks, const void* id128);
int qgd_comm_finalize(void);
int qgd_qgdfoam_set_halo(qgd_solver* s, int n_neighbours, const int* nbr_rank,
                         const int* send_cell_off, const int* send_cells, const int* recv_cell_off, const int* recv_cells,
                         const int* send_bf_off, const int* send_bfaces, const int* recv_bf_off, const int* recv_bfaces);

/* implicitDiffusion true on extended sub-meshes additionally needs the face-neighbour subset of the halo (same list form):
 * the search direction of every PCG iteration of the U and e solves (QGDUEqn.H:54-75, QGDEEqn.H:53-64) and the cell-centred
 * fvc::grad(U) are exchanged over it, the dot products are all-reduced (preconditioner diagonal | none on sub-meshes). */
int qgd_qgdfoam_set_halo_faces(qgd_solver* s, int n_neighbours, const int* nbr_rank, const int* send_off, const int* send_cells,
                               const int* recv_off, const int* recv_cells);

/* ---- QHDFoam (solver-level integration) -------------------------------------
 * Replaces the loop body QHDFoam.C:83-139 (explicit branch) with rhoQGDThermo::New -> heRhoQGDThermo<rhoConst,
 * hConst, const> (rhoQGDThermos.C:76-140, heRhoQGDThermo.C:38-139), a QHD-family QGDCoeffs model
 * (constTau.C:48-85, H2bynuQHD.C:77-83, HbyUQHD.C:79-84, T0byGr.C:83-88) and the fvSolution PCG controls of p.
 * fvSchemes assumed: ddt Euler, grad Gauss linear, laplacian Gauss linear (un)corrected on orthogonal meshes,
 * interpolation linear, no div(phi,U)/div(phi,T) entries (QGDInterpolate.H:86-104 -> flux*psif). */
typedef struct {
    const char* fvsc_scheme;        /* fvSchemes::fvsc::default                                     */
    const char* qgd_coeffs_model;   /* "constTau" | "H2bynuQHD" | "HbyUQHD" | "T0byGr"              */
    double rho0;                    /* rhoConst                                                     */
    double mu, Pr, beta;            /* constTransport + beta (QHDFoam/createFields.H:110-115)       */
    double g[3];                    /* constant/gravitationalProperties (createFields.H:108)        */
    double Tau, UQHD, Gr, T0;       /* model coefficients                                           */
    int implicit_diffusion;         /* QGD::implicitDiffusion: QHDUEqn.H:46-65, QHDTEqn.H:69-80      */
    double p_tolerance, p_rel_tol;  /* fvSolution::solvers::p                                       */
    int p_max_iter;
    const char* p_preconditioner;   /* "DIC" | "diagonal" | "none"                                  */
    int p_ref_cell; double p_ref_value;   /* setRefCell(p, thermo.subDict("QGD"), ...) createFields.H:162-165 */
    int adjust_time_step;           /* QHDCourantNo.H:37                                            */
    double max_co, max_delta_t, c_tau, delta_t;
    /* fvSolution::solvers::"(U|T)" for the implicit-diffusion branch: PCG + preconditioner ("DIC" | "diagonal" | "none") */
    double diff_tolerance, diff_rel_tol;
    int diff_max_iter;
    const char* diff_preconditioner;
    /* 1 = the loop body of scalarTransportQHDFoam (scalarTransportQHDFoam.C:70-135) instead of QHDFoam.C:83-139: U, p
     * and tau stay as initialised, phiu = Sf & Uf is the transporting flux, only
     *   fvm::ddt(T) + fvc::div(phiu*Tf) - fvc::Sp(fvc::div(phiu),T) - fvm::laplacian(Hif,T) - fvc::div(tauQGDf*phiu*(Uf & gradTf)) == 0
     * is solved, and only when implicit_diffusion is set (:114); Courant number from mag(Uf)/hQGDf (:88-96). */
    int scalar_transport;
} qgd_qhdfoam_desc;
int qgd_qhdfoam_create(qgd_mesh* mesh, const qgd_qhdfoam_desc* desc, qgd_qhd_solver** out);
int qgd_qhdfoam_destroy(qgd_qhd_solver* s);
/* kinds per patch for U, T, p: fixedValue | zeroGradient | fixedGradient | qhdFlux (a fixed-gradient patch in this solver:
 * no field is registered as "phiwStar" in QHDFoam, qhdFluxFvPatchScalarField.C:166-206).  val_*: per boundary face, the
 * value on fixedValue patches, the gradient on fixedGradient / qhdFlux patches (U: n_bnd*3). */
int qgd_qhdfoam_set_bcs(qgd_qhd_solver* s, const int* bc_U, const int* bc_T, const int* bc_p,
                        const double* val_U, const double* val_T, const double* val_p);
/* U n_cells*3, T, p n_cells, alphaQGD n_cells or NULL (0.5).  Assembles the (time-constant) pressure matrix and its
 * preconditioner on the device. */
int qgd_qhdfoam_init_fields(qgd_qhd_solver* s, const double* U, const double* T, const double* p, const double* alphaQGD);
/* Decomposed run (one extended sub-mesh per GPU, after qgd_comm_init; call before qgd_qhdfoam_init_fields).  Two list sets in
 * the form of qgd_qgdfoam_set_halo: the vertex-ring halo (state U, T, p after every step, p after the pressure solve) and its
 * face-neighbour subset (search direction of every PCG iteration, fvc::grad(U)).  Replaces the processor-patch updates and
 * the reduce() calls of lduMatrix::solver PCG in a `mpirun -np N QHDFoam -parallel` run (QHDpEqn.H:45).  On
 * sub-meshes: explicit branch, p preconditioner diagonal | none, p_ref_cell = local id on the owning rank and -1 elsewhere. */
int qgd_qhdfoam_set_halo(qgd_qhd_solver* s, int n_neighbours, const int* nbr_rank, const int* send_off, const int* send_cells,
                         const int* recv_off, const int* recv_cells, int n_face_neighbours, const int* nbr_rank_face,
                         const int* fsend_off, const int* fsend_cells, const int* frecv_off, const int* frecv_cells);
int qgd_qhdfoam_step(qgd_qhd_solver* s, int n_steps);
/* fields: 0 U(3), 1 T, 2 p, 3 tauQGD */
int qgd_qhdfoam_get(qgd_qhd_solver* s, int field, double* cells, double* bnd);
/* phi = phiu - phiwo + pEqn.flux() of the last step (QHDpEqn.H:47), n_faces */
int qgd_qhdfoam_get_flux(qgd_qhd_solver* s, double* phi);
int qgd_qhdfoam_get_scalars(qgd_qhd_solver* s, double* delta_t, double* courant, double* time);
/* SolverPerformance of the last pEqn.solve(): nIterations, initialResidual, finalResidual */
int qgd_qhdfoam_solver_info(qgd_qhd_solver* s, int* iters, double* initial_residual, double* final_residual);
long long qgd_qhdfoam_launch_count(qgd_qhd_solver* s);

/* ---- LDU PCG (QHDpEqn.H:45 -> lduMatrix::solver PCG + DIC|diagonal) ---------- */
/* symmetric LDU matrix on the mesh addressing (lower == upper); precond: 0 none, 1 diagonal (Jacobi),
 * 2 DIC (device: level-scheduled, bit-identical operation order to the sequential sweeps).  Host pointers; the whole
 * solve is one cooperative kernel launch.  Returns iterations in *iters. */
int qgd_pcg_solve(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x,
                  double tolerance, double rel_tol, int max_iter, int precond,
                  int* iters, double* initial_residual, double* final_residual);

/* The same solver cut into one kernel per phase (initial residual, preconditioning + wA.rA, search-direction update, SpMV +
 * wA.pA, solution / residual update), all scalars and the convergence flag on the device: the form a decomposed run needs, with
 * a halo exchange of the search direction and all-reduces of the dot products between the phases.  This entry point runs it on
 * one GPU (no communication); precond 0 | 1 | 2 (2 = DIC on the mesh's DIC blocks, which must be set).  Parity: tests/test_gpu_extra.py,
 * tests/test_gpu_qhd.py::test_block_local_dic_matches_the_decomposed_run_oracle. */
int qgd_pcg_solve_stepwise(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x,
                           double tolerance, double rel_tol, int max_iter, int precond,
                           int* iters, double* initial_residual, double* final_residual);

/* The decomposed run of that solver (after qgd_comm_init; every rank calls it): `mesh` is the rank's extended sub-mesh (created
 * with n_owned_cells), diag / b / x have n_cells entries (the rows of the owned cells are solved, halo entries of x are refreshed
 * over NCCL), upper follows the local internal faces.  Exchange lists = the face-neighbour subset of the halo, per neighbour k
 * (rank nbr_rank[k]): local cell ids send_cells[send_off[k] .. send_off[k+1]) and recv_cells[...], ascending global id on both
 * sides.  precond 0 | 1 | 2 (DIC blocks inside the owned cells).  Parity at N = 2 and N = 8: tests/multi_gpu_pcg_worker.py. */
int qgd_pcg_solve_multi(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x,
                        double tolerance, double rel_tol, int max_iter, int precond,
                        int n_neighbours, const int* nbr_rank, const int* send_off, const int* send_cells,
                        const int* recv_off, const int* recv_cells,
                        int* iters, double* initial_residual, double* final_residual);

#ifdef __cplusplus
}
#endif
#endif /* QGD_B200_H */
