// qgd_host_demo — the reference's call sites, written against the host mirror (fvsc::grad/div, QGDFoam / QHDFoam loop
// bodies) and executed on the device through libqgd_b200.so.  Driven by tests/test_host_mirror.py.
//   qgd_host_demo mesh  nx ny nz out.bin                 dump the generated polyMesh + geometry (no GPU needed)
//   qgd_host_demo fvsc  nx ny nz                         operator-level checks on linear fields, error behaviour
//   qgd_host_demo qgdfoam nx ny nz steps in.bin out.bin  QGDFoam: U,T,p from in.bin -> rho,rhoU,rhoE,p to out.bin
//   qgd_host_demo qhdfoam nx ny steps in.bin out.bin     QHDFoam cavity (2D): U,T,p from in.bin -> U,T,p to out.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "QGDFoam.H"
#include "hexMesh.H"

using namespace Foam;

template<class T> static void wr(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }
template<class T> static void rd(FILE* f, std::vector<T>& v) { if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) { std::fprintf(stderr, "short read\n"); std::exit(2); } }

static int fail(const std::string& m) { std::cout << "FAIL: " << m << std::endl; return 1; }

static int runFvsc(label nx, label ny, label nz)
{
    fvMesh mesh;
    hexBoxMesh(mesh, nx, ny, nz, 1.0, 0.8, 0.6);
    mesh.schemes.subDict("fvsc").add("default", word("GaussVolPoint"));
    mesh.schemes.subDict("fvsc").add("grad(q)", word("reduced"));
    qgdCheck(qgd_init(0));
    const vector a{0.7, -1.3, 2.1};
    auto lin = [&](const vector& x) { return 0.5 + a[0] * x[0] + a[1] * x[1] + a[2] * x[2]; };
    const label nI = mesh.nInternalFaces;
    // GaussVolPoint is exact for linear fields wherever the vertex values are: interior points (symmetric inverse-distance
    // weights on this uniform box).  Patch points average boundary-face values, which is not linear-exact at edges, so the
    // exactness checks below use the faces that have no vertex on the boundary.
    std::vector<char> bndPoint(mesh.nPoints(), 0), interiorFace(mesh.nFaces(), 1);
    for (label f = nI; f < mesh.nFaces(); ++f) for (label q = mesh.faceOffsets[f]; q < mesh.faceOffsets[f + 1]; ++q) bndPoint[mesh.faceVerts[q]] = 1;
    for (label f = 0; f < mesh.nFaces(); ++f) for (label q = mesh.faceOffsets[f]; q < mesh.faceOffsets[f + 1]; ++q) if (bndPoint[mesh.faceVerts[q]]) interiorFace[f] = 0;
    // ---- grad(volScalarField): exact for a linear field (fvsc.C:87-101)
    volScalarField phi("phi", mesh);
    for (auto& t : phi.patchTypes) t.type = "fixedValue";
    for (label c = 0; c < mesh.nCells; ++c) phi.internal[c] = lin(mesh.C[c]);
    for (label b = 0; b < mesh.nBoundaryFaces(); ++b) phi.boundary[b] = lin(mesh.Cf[nI + b]);
    tmp<surfaceVectorField> g = fvsc::grad(phi);
    double err = 0;
    for (label f = 0; f < mesh.nFaces(); ++f) if (interiorFace[f]) for (int i = 0; i < 3; ++i) err = std::max(err, std::fabs((*g)[f][i] - a[i]));
    if (err > 1e-11) return fail("grad(phi) of a linear field, err " + std::to_string(err));
    // ---- grad / div of a linear vector field (fvsc.C:109-147): grad(U)_ij = d_i U_j = M_ji, div U = trace
    const double M[3][3] = {{0.3, -0.2, 0.5}, {1.1, 0.4, -0.7}, {-0.6, 0.9, 0.2}};
    auto linU = [&](const vector& x) { vector u; for (int j = 0; j < 3; ++j) u[j] = 0.1 * j + M[j][0] * x[0] + M[j][1] * x[1] + M[j][2] * x[2]; return u; };
    volVectorField U("U", mesh);
    for (auto& t : U.patchTypes) t.type = "fixedValue";
    for (label c = 0; c < mesh.nCells; ++c) U.internal[c] = linU(mesh.C[c]);
    for (label b = 0; b < mesh.nBoundaryFaces(); ++b) U.boundary[b] = linU(mesh.Cf[nI + b]);
    tmp<surfaceTensorField> gU = fvsc::grad(U);
    tmp<surfaceScalarField> dU = fvsc::div(U);
    err = 0;
    for (label f = 0; f < mesh.nFaces(); ++f) {
        if (!interiorFace[f]) continue;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) err = std::max(err, std::fabs((*gU)[f][3 * i + j] - M[j][i]));
        err = std::max(err, std::fabs((*dU)[f] - (M[0][0] + M[1][1] + M[2][2])));
    }
    if (err > 1e-11) return fail("grad(U)/div(U) of a linear field, err " + std::to_string(err));
    // ---- div(volTensorField): (div T)_j = d_i T_ij
    volTensorField TT("TT", mesh);
    for (auto& t : TT.patchTypes) t.type = "fixedValue";
    auto linT = [&](const vector& x) { tensor t; for (int q = 0; q < 9; ++q) t[q] = 0.01 * q + (q + 1) * 0.1 * x[q % 3]; return t; };
    for (label c = 0; c < mesh.nCells; ++c) TT.internal[c] = linT(mesh.C[c]);
    for (label b = 0; b < mesh.nBoundaryFaces(); ++b) TT.boundary[b] = linT(mesh.Cf[nI + b]);
    tmp<surfaceVectorField> dT = fvsc::div(TT);
    err = 0;
    for (label f = 0; f < mesh.nFaces(); ++f)
        for (int j = 0; j < 3 && interiorFace[f]; ++j) {
            double ex = 0;      // sum_i d_i T_ij, T_ij = 0.01 q + (q+1) 0.1 x_{q%3}, q = 3i+j -> d_i only if q%3 == i
            for (int i = 0; i < 3; ++i) { const int q = 3 * i + j; if (q % 3 == i) ex += (q + 1) * 0.1; }
            err = std::max(err, std::fabs((*dT)[f][j] - ex));
        }
    if (err > 1e-11) return fail("div(T) of a linear tensor field, err " + std::to_string(err));
    // ---- per-term scheme selection (fvsc.C:51-58): "grad(q)" -> reduced = nf*snGrad
    volScalarField q("q", mesh);
    for (auto& t : q.patchTypes) t.type = "fixedValue";
    q.internal = phi.internal; q.boundary = phi.boundary;
    tmp<surfaceVectorField> gq = fvsc::grad(q);
    err = 0;
    for (label f = 0; f < mesh.nFaces(); ++f) {
        double an = 0;
        for (int i = 0; i < 3; ++i) an += a[i] * mesh.Sf[f][i] / mesh.magSf[f];
        for (int i = 0; i < 3; ++i) err = std::max(err, std::fabs((*gq)[f][i] - an * mesh.Sf[f][i] / mesh.magSf[f]));
    }
    if (err > 1e-11) return fail("reduced grad(q), err " + std::to_string(err));
    if (mesh.stencilRegistry.size() != 2) return fail("lookupOrNew must keep one stencil per scheme name");
    // ---- error behaviour (fvscStencil.C:72-78, fvsc.C:60-63, QGDCoeffs.C:72-78)
    auto expectFatal = [&](const std::function<void()>& fn, const std::string& needle) {
        try { fn(); } catch (const FatalErrorException& e) { return std::string(e.what()).find(needle) != std::string::npos; }
        return false;
    };
    if (!expectFatal([&] { fvsc::fvscStencil::New("noSuchScheme", mesh); }, "Unknown Model type noSuchScheme")) return fail("unknown scheme message");
    mesh.schemes.subDict("fvsc").add("grad(w)", word("leastSquares"));
    volScalarField w("w", mesh);
    if (!expectFatal([&] { fvsc::grad(w); }, "Can't use leastSquares or leastSquaresOpt in 3D case.")) return fail("leastSquares 3D message");
    dictionary dd;
    if (!expectFatal([&] { qgd::QGDCoeffs::New("noModel", mesh, dd); }, "Unknown QGD coeffs evaluation approach type noModel")) return fail("QGDCoeffs message");
    std::cout << "PASS fvsc host mirror: grad/div exact on linear fields, per-term scheme selection, reference error messages" << std::endl;
    return 0;
}

int main(int argc, char** argv)
{
    try {
        if (argc < 2) { std::cerr << "usage: qgd_host_demo mesh|fvsc|qgdfoam|qhdfoam ..." << std::endl; return 2; }
        const std::string mode = argv[1];
        if (mode == "mesh") {
            fvMesh mesh;
            hexBoxMesh(mesh, atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
            FILE* f = fopen(argv[5], "wb");
            const int hdr[4] = {mesh.nCells, mesh.nFaces(), mesh.nInternalFaces, mesh.nPoints()};
            fwrite(hdr, sizeof(int), 4, f);
            wr(f, mesh.points); wr(f, mesh.faceVerts); wr(f, mesh.owner); wr(f, mesh.neighbour);
            wr(f, mesh.C); wr(f, mesh.V); wr(f, mesh.Cf); wr(f, mesh.Sf); wr(f, mesh.magSf); wr(f, mesh.weights); wr(f, mesh.deltaCoeffs);
            fclose(f);
            return 0;
        }
        if (mode == "fvsc") return runFvsc(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
        if (mode == "qgdfoam") {
            const label nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]), steps = atoi(argv[5]);
            fvMesh mesh;
            hexBoxMesh(mesh, nx, ny, nz);
            mesh.schemes.subDict("fvsc").add("default", word("GaussVolPoint"));
            qgdCheck(qgd_init(0));
            volVectorField U("U", mesh); volScalarField T("T", mesh), p("p", mesh);      // 0/U, 0/T, 0/p: zeroGradient everywhere
            FILE* f = fopen(argv[6], "rb");
            rd(f, U.internal); rd(f, T.internal); rd(f, p.internal);
            fclose(f);
            dictionary thermo, control;
            dictionary& mix = thermo.subDict("mixture");
            mix.add("R", 1.0); mix.add("Cp", 3.5); mix.add("Hf", 0.0); mix.add("mu", 1.0e-3); mix.add("Pr", 0.71);
            thermo.subDict("QGD").add("QGDCoeffs", word("constScPrModel1"));
            thermo.subDict("QGD").add("implicitDiffusion", word("false"));
            thermo.subDict("QGD").subDict("constScPrModel1Dict").add("ScQGD", 1.0);
            thermo.subDict("QGD").subDict("constScPrModel1Dict").add("PrQGD", 1.0);
            control.add("deltaT", 2.0e-4);
            QGDFoamSolver solver(mesh, thermo, control, U, T, p);
            solver.solve(steps);                                             // the loop body QGDFoam.C:90-163, `steps` times
            volScalarField rho("rho", mesh), rhoE("rhoE", mesh); volVectorField rhoU("rhoU", mesh);
            solver.read(rho, 0); solver.read(rhoU, 1); solver.read(rhoE, 2); solver.read(p, 5);
            f = fopen(argv[7], "wb");
            wr(f, rho.internal); wr(f, rhoU.internal); wr(f, rhoE.internal); wr(f, p.internal);
            fclose(f);
            std::cout << "PASS qgdfoam host mirror: " << steps << " steps, t = " << solver.time() << std::endl;
            return 0;
        }
        if (mode == "qhdfoam") {
            const label nx = atoi(argv[2]), ny = atoi(argv[3]), steps = atoi(argv[4]);
            fvMesh mesh;
            hexBoxMesh(mesh, nx, ny, 1, 1.0, 1.0, 0.1, {{"zMin", "empty"}, {"zMax", "empty"}});
            mesh.schemes.subDict("fvsc").add("default", word("GaussVolPoint"));
            qgdCheck(qgd_init(0));
            volVectorField U("U", mesh); volScalarField T("T", mesh), p("p", mesh);
            for (size_t pi = 0; pi < mesh.boundaryMesh.size(); ++pi) {
                const word& nm = mesh.boundaryMesh[pi].name;
                U.patchTypes[pi].type = "fixedValue";                                    // no-slip walls
                T.patchTypes[pi].type = (nm == "xMin" || nm == "xMax") ? "fixedValue" : "zeroGradient";
                p.patchTypes[pi].type = "qhdFlux";
                if (nm == "xMin") for (label i = 0; i < mesh.boundaryMesh[pi].size; ++i) T.boundary[mesh.boundaryMesh[pi].start + i - mesh.nInternalFaces] = 1.0;
            }
            FILE* f = fopen(argv[5], "rb");
            rd(f, U.internal); rd(f, T.internal); rd(f, p.internal);
            fclose(f);
            dictionary thermo, control, pSolver;
            dictionary& mix = thermo.subDict("mixture");
            mix.add("rho", 1.0); mix.add("mu", 1.0e-2); mix.add("Pr", 0.71); mix.add("beta", 3.0e-3);
            thermo.subDict("QGD").add("QGDCoeffs", word("constTau"));
            thermo.subDict("QGD").add("implicitDiffusion", word("false"));
            thermo.subDict("QGD").subDict("constTauDict").add("Tau", 1.0e-3);
            pSolver.add("tolerance", 1e-13); pSolver.add("relTol", 0.0); pSolver.add("maxIter", 5000); pSolver.add("preconditioner", word("DIC"));
            control.add("deltaT", 1.0e-3);
            QHDFoamSolver solver(mesh, thermo, vector{0.0, -9.81, 0.0}, pSolver, control, U, T, p);
            solver.solve(steps);                                             // the loop body QHDFoam.C:83-139
            solver.read(U, T, p);
            f = fopen(argv[6], "wb");
            wr(f, U.internal); wr(f, T.internal); wr(f, p.internal);
            fclose(f);
            std::cout << "PASS qhdfoam host mirror: " << steps << " steps, last pEqn iterations = " << solver.pIterations() << std::endl;
            return 0;
        }
        std::cerr << "unknown mode " << mode << std::endl;
        return 2;
    } catch (const FatalErrorException& e) {
        std::cerr << e.what() << std::endl;      // the reference prints and aborts (exit(FatalError))
        return 1;
    }
}
