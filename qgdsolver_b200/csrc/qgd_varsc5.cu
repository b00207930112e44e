// varScModel5 on the device: CUDA executor of the functors / sequences of qgd_varsc5.h and the per-solver device state.
#include "qgd_varsc5_dev.h"

namespace qgd {

namespace {

template <class F> __global__ void __launch_bounds__(256) k_v5_for(int n, F f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

// one kernel launch per forEach; the sizes of the changed lists come back to the host (one 4-byte read per half-sweep)
struct DeviceExec {
    cudaStream_t st;
    long long launches = 0;
    template <class F> void forEach(int n, const F& f)
    {
        if (n <= 0) return;
        k_v5_for<F><<<(n + 255) / 256, 256, 0, st>>>(n, f);
        ++launches;
    }
    void fillInt(int* p, int value, size_t n)          // value 0 or -1 (byte pattern)
    {
        if (n) QGD_CUDA(cudaMemsetAsync(p, value ? 0xFF : 0, n * sizeof(int), st));
    }
    int readInt(const int* p)
    {
        int v = 0;
        QGD_CUDA(cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, st));
        QGD_CUDA(cudaStreamSynchronize(st));
        return v;
    }
};

} // namespace

void VarSc5Device::create(const HostMesh& h, const Consts& k, double rC, double minSc, double maxSc, double ScDict, double smoothCoeff,
                          double badQualitySc, double maxAspectRatio, cudaStream_t st)
{
    VarSc5Host hm;
    buildVarSc5Host(h, badQualitySc, maxAspectRatio, hm);
    nC = hm.nC; nI = hm.nI; nB = hm.nB; maxCellFaces = hm.maxCellFaces;
    auto upI = [&](DevBuf<int>& d, std::vector<int> v) { if (v.empty()) v.push_back(0); d.upload(v, st); };
    auto upD = [&](DevBuf<double>& d, std::vector<double> v) { if (v.empty()) v.push_back(0.0); d.upload(v, st); };
    upI(own, hm.own); upI(nei, hm.nei); upD(w, hm.w); upD(Sf, hm.Sf); upD(bMagSf, hm.bMagSf); upD(bDC, hm.bDC); upD(bHf, hm.bHf);
    upI(bKind, hm.bKind); upI(ccOff, hm.ccOff); upI(ccFace, hm.ccFace); upI(lidxOwn, hm.lidxOwn); upI(lidxNei, hm.lidxNei);
    upD(cqSc, hm.cqSc);
    this->k = k; this->rC = rC; this->minSc = minSc; this->maxSc = maxSc; this->ScDict = ScDict; this->maxRatio = 1.0 + smoothCoeff;
    pOld.alloc(nC); pOldB.alloc(nB + 1); ScB.alloc(nB + 1); rho.alloc(nC); rhoB.alloc(nB + 1); g.alloc(3 * (size_t)nC);
    faceV.alloc(nI + 1); posF.alloc(nI + 1); posC.alloc(nC); listF.alloc(nI + 1); listC.alloc(nC);
    const size_t nMark = std::max<size_t>(2 * (size_t)nI, (size_t)nC * maxCellFaces) + 1;
    mark.alloc(nMark); cnt.alloc(kV5CompactThreads + 1); nOut.alloc(1);
    std::vector<double> scb(nB + 1, ScDict);           // ScQGD_.boundaryFieldRef() = ScQGD (varScModel5.C:79)
    QGD_CUDA(cudaMemcpyAsync(ScB.p, scb.data(), scb.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    QGD_CUDA(cudaStreamSynchronize(st));
}

VarSc5View VarSc5Device::view(double* S, RecA* bA, RecB* bB, const double* psiB, const double* aQGD, const double* V, const double* hQGD,
                              double* Sc, const unsigned char* scConst) const
{
    VarSc5View v{};
    v.nC = nC; v.nI = nI; v.nB = nB; v.maxCellFaces = maxCellFaces;
    v.own = own.p; v.nei = nei.p; v.w = w.p; v.Sf = Sf.p; v.bMagSf = bMagSf.p; v.bDC = bDC.p; v.bHf = bHf.p; v.bKind = bKind.p;
    v.ccOff = ccOff.p; v.ccFace = ccFace.p; v.lidxOwn = lidxOwn.p; v.lidxNei = lidxNei.p;
    v.V = V; v.hQGD = hQGD; v.cqSc = cqSc.p; v.scConst = scConst;
    v.rC = rC; v.minSc = minSc; v.maxSc = maxSc; v.ScDict = ScDict; v.maxRatio = maxRatio; v.k = k;
    v.S = S; v.bA = bA; v.bB = bB; v.psiB = psiB; v.aQGD = aQGD; v.pOld = pOld.p; v.pOldB = pOldB.p; v.Sc = Sc; v.ScB = ScB.p;
    v.rho = rho.p; v.rhoB = rhoB.p; v.g = g.p; v.faceV = faceV.p; v.posF = posF.p; v.posC = posC.p; v.listF = listF.p; v.listC = listC.p;
    v.mark = mark.p; v.cnt = cnt.p; v.nOut = nOut.p;
    return v;
}

long long VarSc5Device::correct(const VarSc5View& v, cudaStream_t st)
{
    DeviceExec ex{st};
    lastSmoothIters = v5Correct(ex, v);
    QGD_CUDA(cudaGetLastError());
    return ex.launches;
}

} // namespace qgd
