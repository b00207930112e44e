// Per-solver device state of varScModel5 (see qgd_varsc5.h for the algorithm)
#pragma once
#include <algorithm>

#include "qgd_varsc5.h"

namespace qgd {

struct VarSc5Device {
    int nC = 0, nI = 0, nB = 0, maxCellFaces = 0;
    DevBuf<int> own, nei, bKind, ccOff, ccFace, lidxOwn, lidxNei, posF, posC, listF, listC, mark, cnt, nOut;
    DevBuf<double> w, Sf, bMagSf, bDC, bHf, cqSc, pOld, pOldB, ScB, rho, rhoB, g, faceV;
    Consts k{};
    double rC = 0.5, minSc = 0.05, maxSc = 1.0, ScDict = 1.0, maxRatio = 1.1;
    int lastSmoothIters = 0;             // FaceCellWave iterations of the last fvc::smooth
    // dictionary entries of varScModel5.C:61-110; builds the host tables (buildVarSc5Host) and uploads them
    void create(const HostMesh& h, const Consts& k, double rC, double minSc, double maxSc, double ScDict, double smoothCoeff,
                double badQualitySc, double maxAspectRatio, cudaStream_t st);
    VarSc5View view(double* S, RecA* bA, RecB* bB, const double* psiB, const double* aQGD, const double* V, const double* hQGD,
                    double* Sc, const unsigned char* scConst) const;
    // varScModel5::correct on a state closed by the ordinary kernels (pOld / pOldB filled by the caller); returns kernel launches
    long long correct(const VarSc5View& v, cudaStream_t st);
};

} // namespace qgd
