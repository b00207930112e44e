// C-ABI layer of libqgd_b200.so (include/qgd_b200.h): opaque handles, model selection with the reference's
// error behaviour, H2D/D2H staging for the operator-level calls.  No compute happens on the host here.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

#include <memory>

#include "qgd_kernels.cuh"
#include "qgd_pcg.cuh"
#include "qgd_varsc5_dev.h"

namespace qgd {

static thread_local std::string g_lastError;
void setLastError(const std::string& m) { g_lastError = m; }

static cudaStream_t g_stream = nullptr;
static cudaStream_t g_commStream = nullptr;        // halo exchange overlapped with the interior point gather
static cudaStream_t g_sideStream = nullptr;        // boundary kernels beside the point gather (StepFork), high priority
static cudaEvent_t g_evFork[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static int g_bndFork = 1;                          // env QGD_BND_FORK=0 keeps the boundary kernels in line on the main stream
static cudaEvent_t g_evStep = nullptr, g_evHalo = nullptr;
static bool g_initialised = false;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

// ---- NCCL, bound at run time (dlopen) so that single-GPU users need no NCCL and the copy already loaded by the
// launcher process (e.g. torch's) is reused
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
[[maybe_unused]] static int g_rank = 0;
static int g_nranks = 1;

static void loadNccl()
{
    if (g_nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) throw Error(QGD_ERR_COMM, std::string("cannot load libnccl: ") + dlerror());
    auto sym = [&](const char* n) { void* p = dlsym(g_nccl.lib, n); if (!p) throw Error(QGD_ERR_COMM, std::string("libnccl lacks ") + n); return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
}
#define QGD_NCCL(expr)                                                                                        \
    do {                                                                                                      \
        ncclResult_t _r = (expr);                                                                             \
        if (_r != ncclSuccess)                                                                                \
            throw qgd::Error(QGD_ERR_COMM, std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " at " + \
                                               __FILE__ + ":" + std::to_string(__LINE__));                    \
    } while (0)

void requireInit()
{
    if (!g_initialised) throw Error(QGD_ERR_STATE, "qgd_init(device) must be called first (no CPU fallback exists)");
}
cudaStream_t runtimeStream() { return g_stream; }

// runTimeSelection tables of the path (names only; the OpenFOAM shim registers the same TypeNames)
static const char* const kFvscTable[] = {"GaussVolPoint", "leastSquares", "leastSquaresOpt", "reduced"};          // sortedToc order
static const char* const kCoeffsTable[] = {"H2bynuQHD", "HbyUQHD", "T0byGr", "constScPrModel1", "constScPrModel1n",
                                           "constScPrModel2", "constTau", "varScModel5", "varScModel6", "varScModel7"};

template <size_t N> static std::string toc(const char* const (&t)[N])
{
    std::string s = std::to_string(N) + "\n(\n";
    for (size_t i = 0; i < N; ++i) { s += t[i]; s += "\n"; }
    s += ")\n";
    return s;
}
template <size_t N> static bool inTable(const char* const (&t)[N], const std::string& n)
{
    for (size_t i = 0; i < N; ++i) if (n == t[i]) return true;
    return false;
}
bool isCoeffsModel(const std::string& n) { return inTable(kCoeffsTable, n); }
std::string coeffsModelToc() { return toc(kCoeffsTable); }

} // namespace qgd

using namespace qgd;

struct qgd_solver {
    qgd_mesh* mesh = nullptr;
    std::unique_ptr<qgd_fvsc> fvsc;
    qgd_qgdfoam_desc desc{};
    Consts k{};
    DevBuf<RecA> bA;
    DevBuf<RecB> bB;
    DevBuf<double> S, P;    // cell state 16 x nCells, point values 6 x nPoints (SoA)
    DevBuf<double> aQGD, Fflux, psiB, pGrad, pNew, phiw, bvU, bvT, bvP, stage, tauOut, tauOutB, scVar;
    DevBuf<unsigned char> scConst;   // varScModel7 / varScModel5 constScCellSet mask
    std::unique_ptr<VarSc5Device> v5; // varScModel5: its own pass after every step (qgd_varsc5.h); scVar holds its ScQGD
    VarSc5View v5view()
    {
        return v5->view(S.p, bA.p, bB.p, psiB.p, aQGD.p, mesh->V.p, mesh->hQGD.p, scVar.p, scConst.n ? scConst.p : nullptr);
    }
    DevBuf<double> su;               // [5][nCells] explicit sources (qgd_qgdfoam_set_sources) or empty
    int stepsDone = 0;
    long long graphSteps = 0;        // steps replayed from a captured CUDA graph (QGD_STEP_GRAPH=1)
    // implicit-diffusion branch
    struct Implicit {
        DevBuf<double> GU0, GU1, old, FT, aU, aE, Fs, diagU, bU, diagE, bE;
        PcgMatrix AU, AE;
        int precond = 2;
        bool built = false;
    } impl;
    ImplicitView iview()
    {
        ImplicitView v;
        v.GU0 = impl.GU0.p; v.GU1 = impl.GU1.p; v.old = impl.old.p; v.FT = impl.FT.p; v.aU = impl.aU.p; v.aE = impl.aE.p; v.Fs = impl.Fs.p;
        v.diagU = impl.diagU.p; v.bU = impl.bU.p; v.diagE = impl.diagE.p; v.bE = impl.bE.p;
        return v;
    }
    // face fluxes (layout: SolverView::F): one [5][nF] array (two-kernel form) or [5][ring] + [5][nB] (pipelined form)
    size_t strideI = 0, strideB = 0, bndOff = 0;
    int ringSize = 0x7fffffff;
    // pipelined face+cell kernel (fixed deltaT): plan + flags.  mode: 0 two kernels, 1 pipelined
    struct Pipe {
        int mode = 0, chunkCells = 512, lag = 0, ringSlots = 0, nChunks = 0, epoch = 0, grid = 0;
        bool windowSet = false;
        DevBuf<int> cellOff, faceOff, depOff, depList, ringOff, ringList, doneF, doneC, queue;
    } pipe;
    PipeView pview()
    {
        PipeView v;
        v.nChunks = pipe.nChunks; v.lag = std::min(pipe.lag, pipe.nChunks); v.epoch = pipe.epoch;
        v.cellOff = pipe.cellOff.p; v.faceOff = pipe.faceOff.p; v.depOff = pipe.depOff.p; v.depList = pipe.depList.p;
        v.ringOff = pipe.ringOff.p; v.ringList = pipe.ringList.p;
        v.doneF = pipe.doneF.p; v.doneC = pipe.doneC.p; v.queue = pipe.queue.p;
        return v;
    }
    DevBuf<int> bcU, bcT, bcP;
    DevBuf<StepScalars> sc;
    long long launches = 0;
    bool anyQgdFlux = false, bcsSet = false, fieldsSet = false;
    int gridFaces = 148;
    // halo exchange (multi-GPU): per neighbour rank, contiguous slices of the id lists / buffers
    struct Halo {
        std::vector<int> nbr, sendCellOff, recvCellOff, sendBfOff, recvBfOff;
        DevBuf<int> sendCells, recvCells, sendBf, recvBf;
        DevBuf<double> sendBuf, recvBuf;      // [cells: 16 doubles each][bfaces: 20 doubles each] per neighbour
        DevBuf<double> midSend, midRecv;      // p_b of halo boundary faces (qgdFlux mid-step refresh)
        DevBuf<int> planSendItem, planSendCell, planSendBf, planRecvItem, planRecvCell, planRecvBf;   // k_halo_all plans
        DevBuf<long long> planSendBuf, planRecvBuf;
        DevBuf<int> ptsInterior, ptsHalo;      // point lists for the overlapped exchange
        bool pending = false;                  // an exchange is in flight on the communication stream
        bool active = false;
    } halo;
    HaloLists haloFace;                        // face-neighbour subset of the halo (implicit branch: PCG search direction, fvc::grad(U))
    StepwisePcg sw;                            // decomposed linear solves of the implicit branch
    // per-kernel CUDA-event timing (bench): events of the profiled steps, 6 per step
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    size_t eventsUsed = 0;
    ~qgd_solver() { for (cudaEvent_t e : events) cudaEventDestroy(e); }
    SolverView sview() const
    {
        const qgd_mesh& m = *mesh;
        SolverView s;
        s.nCells = m.h.nCells; s.nPoints = m.h.nPoints; s.nPatchPoints = (int)m.h.patchPoints.size();
        s.nOwned = m.h.nOwned;
        s.S = S.p; s.P = P.p;
        s.pcEllW = m.pcEllW; s.pcEll = m.pcEll.p; s.pcEllWt = m.pcEllWt.p; s.pcCount = m.pcCount.p;
        s.pcTailOff = m.pcTailOff.p; s.pcTailCell = m.pcTailCell.p; s.pcTailW = m.pcTailW.p;
        s.patchPoints = m.patchPoints.p; s.ppOff = m.ppOff.p; s.ppFace = m.ppFace.p; s.ppW = m.ppW.p;
        s.cfEllW = m.cfEllW; s.cfEll = m.cfEll.p; s.cfTailOff = m.cfTailOff.p; s.cfTailEnc = m.cfTailEnc.p; s.V = m.V.p; s.hQGD = m.hQGD.p; s.aQGD = aQGD.p;
        s.tauOut = tauOut.n ? tauOut.p : nullptr;
        s.scVar = scVar.n ? scVar.p : nullptr; s.scConst = scConst.n ? scConst.p : nullptr;
        s.su = su.n ? su.p : nullptr;
        const bool split = halo.active && halo.ptsInterior.n > 0 && halo.ptsHalo.n > 0;
        s.ptsInterior = split ? halo.ptsInterior.p : nullptr; s.nPtsInterior = split ? (int)halo.ptsInterior.n : 0;
        s.ptsHalo = split ? halo.ptsHalo.p : nullptr; s.nPtsHalo = split ? (int)halo.ptsHalo.n : 0;
        for (int q = 0; q < 5; ++q) { s.FI[q] = Fflux.p + q * strideI; s.FB[q] = Fflux.p + bndOff + q * strideB; }
        s.ringSize = ringSize; s.sc = sc.p;
        return s;
    }
    BndState bview() const
    {
        BndState b;
        b.A = bA.p; b.B = bB.p; b.psi = psiB.p; b.pGrad = pGrad.p; b.pNew = pNew.p; b.phiw = phiw.p;
        b.bcU = bcU.p; b.bcT = bcT.p; b.bcP = bcP.p; b.bvU = bvU.p; b.bvT = bvT.p; b.bvP = bvP.p;
        b.tauOutB = tauOutB.n ? tauOutB.p : nullptr;
        return b;
    }
};

namespace {

template <class T> void h2d(DevBuf<T>& d, const T* h, size_t n)
{
    if (d.n < n) d.alloc(n);
    if (n) QGD_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, g_stream));
}
template <class T> void d2h(T* h, const T* d, size_t n)
{
    if (n) QGD_CUDA(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, g_stream));
}

// state hand-off kernels for qgd_qgdfoam_step_host
__global__ void k_pack_state(Consts k, int n, double* S, const double* __restrict__ aQGD, const double* __restrict__ hQGD,
                             const double* __restrict__ st)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const size_t N = n;
    const double rho = st[c], Ux = st[N + 3 * (size_t)c], Uy = st[N + 3 * (size_t)c + 1], Uz = st[N + 3 * (size_t)c + 2];
    const double e = st[4 * N + c], p = st[5 * N + c], T = st[6 * N + c];
    const double rUx = st[7 * N + 3 * (size_t)c], rUy = st[7 * N + 3 * (size_t)c + 1], rUz = st[7 * N + 3 * (size_t)c + 2];
    const double rhoE = st[10 * N + c], mu = st[11 * N + c];
    const double psi = 1.0 / (k.R * T);
    const double cs = sqrt(k.gamma / psi);
    // thermo:mu = mu(T) + muQGD (QGDThermo.C:91-98): the molecular part follows the transport model, the rest is muQGD
    double muT = k.mu, aT = k.mu / k.Pr;
    if (k.transport == 1) { muT = k.mu0 * pow(T / k.T0, k.kExp); aT = muT * k.rPr; }
    else if (k.transport == 2) { muT = k.As * sqrt(T) / (1.0 + k.Ts / T); aT = muT * k.Cv * (1.32 + 1.77 * k.R / k.Cv) / k.Cp; }
    const double alpha = aT + (mu - muT) / k.PrQGD;
    const double v[16] = {rho, Ux, Uy, Uz, e, p, T, (rhoE + p) / rho,
                          rUx, rUy, rUz, rhoE, cs, mu, k.alphaEffGamma ? k.gamma * alpha : alpha,
                          k.tauMode == 1 ? aQGD[c] * hQGD[c] / (sqrt(Ux * Ux + Uy * Uy + Uz * Uz) + cs) : (k.tauMode == 2 ? aQGD[c] : aQGD[c] / cs)};
#pragma unroll
    for (int j = 0; j < 16; ++j) S[j * N + c] = v[j];
}
__global__ void k_unpack_state(int n, const double* __restrict__ S, double* st)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const size_t N = n;
    st[c] = S[c];
    st[N + 3 * (size_t)c] = S[N + c]; st[N + 3 * (size_t)c + 1] = S[2 * N + c]; st[N + 3 * (size_t)c + 2] = S[3 * N + c];
    st[4 * N + c] = S[4 * N + c]; st[5 * N + c] = S[5 * N + c]; st[6 * N + c] = S[6 * N + c];
    st[7 * N + 3 * (size_t)c] = S[8 * N + c]; st[7 * N + 3 * (size_t)c + 1] = S[9 * N + c]; st[7 * N + 3 * (size_t)c + 2] = S[10 * N + c];
    st[10 * N + c] = S[11 * N + c]; st[11 * N + c] = S[13 * N + c];
}

// U (AoS), T, p [, rho, rhoU (AoS), rhoE] of the owned and halo cells into the staging buffer: [3n | n | n | n | 3n | n]
__global__ void k_unpack_fields(int n, const double* __restrict__ S, double* __restrict__ st, int conserved)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const size_t N = n;
    st[3 * (size_t)c] = S[N + c]; st[3 * (size_t)c + 1] = S[2 * N + c]; st[3 * (size_t)c + 2] = S[3 * N + c];
    st[3 * N + c] = S[6 * N + c]; st[4 * N + c] = S[5 * N + c];
    if (conserved) {
        st[5 * N + c] = S[c];
        st[6 * N + 3 * (size_t)c] = S[8 * N + c]; st[6 * N + 3 * (size_t)c + 1] = S[9 * N + c]; st[6 * N + 3 * (size_t)c + 2] = S[10 * N + c];
        st[9 * N + c] = S[11 * N + c];
    }
}

// ---- halo exchange: pack (gather) -> ncclSend/ncclRecv in one group -> unpack (scatter)
constexpr int kCellDoubles = 16, kBfDoubles = 20;
// all neighbours in one launch: item i of the concatenated (cells | boundary faces) lists belongs to neighbour k with
// itemOff[k] <= i < itemOff[k+1]; its block starts at bufOff[k] doubles and holds nC_k cells then nB_k boundary faces
struct HaloPlan { int nn; const int* itemOff; const int* cellOff; const int* bfOff; const long long* bufOff; };
template <bool PACK>
__global__ void k_halo_all(HaloPlan hp, const int* __restrict__ cells, const int* __restrict__ bfaces, double* __restrict__ S, size_t nCells,
                           RecA* __restrict__ bA, RecB* __restrict__ bB, double* __restrict__ psi, double* __restrict__ pGrad,
                           double* __restrict__ pNew, double* __restrict__ phiw, double* __restrict__ buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hp.itemOff[hp.nn]) return;
    int k = 0;
    while (i >= hp.itemOff[k + 1]) ++k;                    // <= 26 neighbours
    const int nC = hp.cellOff[k + 1] - hp.cellOff[k], j = i - hp.itemOff[k];
    double* blk = buf + hp.bufOff[k];
    if (j < nC) {
        const int c = cells[hp.cellOff[k] + j];
#pragma unroll
        for (int q = 0; q < kCellDoubles; ++q) {
            if (PACK) blk[(size_t)q * nC + j] = S[(size_t)q * nCells + c];
            else S[(size_t)q * nCells + c] = blk[(size_t)q * nC + j];
        }
    } else {
        const int jb = j - nC;
        const int b = bfaces[hp.bfOff[k] + jb];
        double* o = blk + (size_t)kCellDoubles * nC + (size_t)jb * kBfDoubles;
        double* a = reinterpret_cast<double*>(bA + b);
        double* bb = reinterpret_cast<double*>(bB + b);
        if (PACK) {
#pragma unroll
            for (int q = 0; q < 8; ++q) { o[q] = a[q]; o[8 + q] = bb[q]; }
            o[16] = psi[b]; o[17] = pGrad[b]; o[18] = pNew[b]; o[19] = phiw[b];
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) { a[q] = o[q]; bb[q] = o[8 + q]; }
            psi[b] = o[16]; pGrad[b] = o[17]; pNew[b] = o[18]; phiw[b] = o[19];
        }
    }
}

__global__ void k_gather1(int n, const int* __restrict__ ids, const double* __restrict__ src, double* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[ids[i]];
}
__global__ void k_scatter1(int n, const int* __restrict__ ids, const double* __restrict__ src, double* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = src[i];
}

// generic cell-field exchange (QHDFoam state, PCG search direction, Gauss gradients): component-major blocks per neighbour
__global__ void k_halo_fields(int nn, const int* __restrict__ off, const int* __restrict__ ids, double* __restrict__ base, size_t stride,
                              int nComp, int maxComp, double* __restrict__ buf, int pack)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= off[nn]) return;
    int k = 0;
    while (i >= off[k + 1]) ++k;
    const int cnt = off[k + 1] - off[k], j = i - off[k];
    double* blk = buf + (size_t)maxComp * off[k];
    const int c = ids[i];
    for (int q = 0; q < nComp; ++q) {
        if (pack) blk[(size_t)q * cnt + j] = base[(size_t)q * stride + c];
        else base[(size_t)q * stride + c] = blk[(size_t)q * cnt + j];
    }
}

// full state exchange after the cell update (SURVEY 5.8 C1+C3 in one message per neighbour)
int haloExchange(qgd_solver* s, cudaStream_t g_stream = qgd::g_stream)
{
    qgd_solver::Halo& h = s->halo;
    if (!h.active) return 0;
    const size_t nCells = s->mesh->h.nCells;
    const int nn = (int)h.nbr.size();
    int launches = 0;
    auto blockOff = [&](const std::vector<int>& cOff, const std::vector<int>& bOff, int k) {
        return (size_t)kCellDoubles * cOff[k] + (size_t)kBfDoubles * bOff[k];
    };
    {
        HaloPlan hp{nn, h.planSendItem.p, h.planSendCell.p, h.planSendBf.p, h.planSendBuf.p};
        const int items = h.sendCellOff[nn] + h.sendBfOff[nn];
        if (items) {
            k_halo_all<true><<<(items + 255) / 256, 256, 0, g_stream>>>(hp, h.sendCells.p, h.sendBf.p, s->S.p, nCells, s->bA.p, s->bB.p, s->psiB.p,
                                                                        s->pGrad.p, s->pNew.p, s->phiw.p, h.sendBuf.p);
            ++launches;
        }
    }
    QGD_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < nn; ++k) {
        const size_t ns = blockOff(h.sendCellOff, h.sendBfOff, k + 1) - blockOff(h.sendCellOff, h.sendBfOff, k);
        const size_t nr = blockOff(h.recvCellOff, h.recvBfOff, k + 1) - blockOff(h.recvCellOff, h.recvBfOff, k);
        if (ns) QGD_NCCL(g_nccl.Send(h.sendBuf.p + blockOff(h.sendCellOff, h.sendBfOff, k), ns, ncclDouble, h.nbr[k], g_comm, g_stream));
        if (nr) QGD_NCCL(g_nccl.Recv(h.recvBuf.p + blockOff(h.recvCellOff, h.recvBfOff, k), nr, ncclDouble, h.nbr[k], g_comm, g_stream));
    }
    QGD_NCCL(g_nccl.GroupEnd());
    {
        HaloPlan hp{nn, h.planRecvItem.p, h.planRecvCell.p, h.planRecvBf.p, h.planRecvBuf.p};
        const int items = h.recvCellOff[nn] + h.recvBfOff[nn];
        if (items) {
            k_halo_all<false><<<(items + 255) / 256, 256, 0, g_stream>>>(hp, h.recvCells.p, h.recvBf.p, s->S.p, nCells, s->bA.p, s->bB.p, s->psiB.p,
                                                                         s->pGrad.p, s->pNew.p, s->phiw.p, h.recvBuf.p);
            ++launches;
        }
    }
    return launches;
}

// mid-step refresh of p_b on halo boundary faces after the qgdFlux re-evaluation (QGDFoam/updateFluxes.H:63-65)
int haloExchangeMid(qgd_solver* s)
{
    qgd_solver::Halo& h = s->halo;
    if (!h.active) return 0;
    const int nn = (int)h.nbr.size();
    int launches = 0;
    const int nS = h.sendBfOff[nn], nR = h.recvBfOff[nn];
    if (nS) { k_gather1<<<(nS + 255) / 256, 256, 0, g_stream>>>(nS, h.sendBf.p, s->pNew.p, h.midSend.p); ++launches; }
    QGD_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < nn; ++k) {
        const int ns = h.sendBfOff[k + 1] - h.sendBfOff[k], nr = h.recvBfOff[k + 1] - h.recvBfOff[k];
        if (ns) QGD_NCCL(g_nccl.Send(h.midSend.p + h.sendBfOff[k], ns, ncclDouble, h.nbr[k], g_comm, g_stream));
        if (nr) QGD_NCCL(g_nccl.Recv(h.midRecv.p + h.recvBfOff[k], nr, ncclDouble, h.nbr[k], g_comm, g_stream));
    }
    QGD_NCCL(g_nccl.GroupEnd());
    if (nR) { k_scatter1<<<(nR + 255) / 256, 256, 0, g_stream>>>(nR, h.recvBf.p, h.midRecv.p, s->pNew.p); ++launches; }
    return launches;
}

// (Re)build the flux storage and, for mode 1, the work plan of k_face_cell_pipeline.
//   chunkCells: cells per work item (multiple of 32; 0 = default) ; lag: chunks between F(k) and C(k) in the queue (<0 = default)
//   ringSlots: ring size in faces (power of two; 0 = sized from the mesh band width and the L2 persisting carve-out)
void configurePipeline(qgd_solver* s, int mode, int chunkCells, int lag, int ringSlots)
{
    const qgd_mesh& m = *s->mesh;
    const HostMesh& h = m.h;
    const int nI = h.nInternal, nIA = m.nIActive, nOwn = h.nOwned;
    qgd_solver::Pipe& P = s->pipe;
    if (mode == 1 && s->desc.adjust_time_step) mode = 0;        // the global Courant max is needed before any cell update
    if (mode == 1 && (nIA == 0 || s->fvsc->lsq)) mode = 0;
    auto fullArrays = [&] {
        P.mode = 0;
        s->strideI = s->strideB = (size_t)h.nFaces; s->bndOff = (size_t)nI;
        s->ringSize = 0x7fffffff;
        s->Fflux.alloc(5 * (size_t)h.nFaces + 1); s->Fflux.zero(g_stream);
    };
    if (mode == 0) { fullArrays(); return; }
    if (const char* v = getenv("QGD_PIPE_CHUNK")) if (!chunkCells) chunkCells = atoi(v);
    if (const char* v = getenv("QGD_PIPE_LAG")) if (lag < 0) lag = atoi(v);
    if (const char* v = getenv("QGD_PIPE_RING")) if (!ringSlots) ringSlots = atoi(v);
    if (chunkCells <= 0) chunkCells = 1024;
    chunkCells = ((chunkCells + 31) / 32) * 32;
    P.grid = pipelineKernelGrid(m.cfEllW);
    if (lag < 0) lag = P.grid / 2 + 8;
    const int nCh = (nOwn + chunkCells - 1) / chunkCells;
    lag = std::min(lag, nCh);
    // device-order face -> owner chunk must be non-decreasing (faces are sorted by owner tile of 32)
    std::vector<int> cellOff(nCh + 1), faceOff(nCh + 1, 0), depLo(nCh);
    std::vector<std::vector<int>> deps(nCh), cons(nCh);           // C(k) <- F(deps[k]) ; faces owned by chunk i are read by C(cons[i])
    auto addUnique = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int k = 0; k <= nCh; ++k) cellOff[k] = std::min(k * chunkCells, nOwn);
    for (int k = 0; k < nCh; ++k) depLo[k] = k;
    int prev = 0;
    for (int f = 0; f < nIA; ++f) {
        const int pf = m.facePerm[f];
        const int co = h.owner[pf] / chunkCells;
        if (co < prev) { fullArrays(); return; }                  // unexpected order: keep the two-kernel form
        prev = co;
        faceOff[co + 1]++;
        addUnique(deps[co], co); addUnique(cons[co], co);
        const int nb = h.neighbour[pf];
        if (nb < nOwn) {
            const int cn = nb / chunkCells;
            if (cn < co) { fullArrays(); return; }                // not upper-triangular
            addUnique(deps[cn], co); addUnique(cons[co], cn);
            depLo[cn] = std::min(depLo[cn], co);
        }
    }
    for (int k = 0; k < nCh; ++k) faceOff[k + 1] += faceOff[k];
    // ring size
    int maxChunkFaces = 1;
    for (int k = 0; k < nCh; ++k) maxChunkFaces = std::max(maxChunkFaces, faceOff[k + 1] - faceOff[k]);
    int R = ringSlots;
    if (R <= 0) {
        // faces that must coexist: from the oldest face a C item still needs to the newest face in flight
        // (lag chunks queued ahead of C(j) + one F item per two resident CTAs)
        long need = 0;
        for (int j = 0; j < nCh; ++j) {
            const int hiChunk = std::min(nCh, j + lag + P.grid / 2 + 8);
            need = std::max(need, (long)faceOff[hiChunk] - faceOff[depLo[j]]);
        }
        R = (int)std::min<long>(need, nIA);
    }
    R = std::max(R, 2 * maxChunkFaces);
    R = ((R + 31) / 32) * 32;
    std::vector<std::vector<int>> ring(nCh);
    if (R >= nIA) R = ((nIA + 31) / 32) * 32;                    // no wrap-around
    else {
        // F(k) rewrites the slots of faces [faceOff[k]-R, faceOff[k+1]-R): their consumers must be done
        auto chunkOf = [&](int f) { return (int)(std::upper_bound(faceOff.begin(), faceOff.end(), f) - faceOff.begin()) - 1; };
        std::vector<int> posF(nCh), posC(nCh);                    // queue position of every item
        for (int k = 0; k < nCh; ++k) posF[k] = k < lag ? k : lag + 2 * (k - lag);
        for (int j = 0; j < nCh; ++j) posC[j] = j < nCh - lag ? lag + 2 * j + 1 : lag + 2 * (nCh - lag) + (j - (nCh - lag));
        bool ok = true;
        for (int k = 0; k < nCh && ok; ++k) {
            const int a = faceOff[k] - R, b = faceOff[k + 1] - R;
            if (b <= 0 || faceOff[k + 1] == faceOff[k]) continue;
            const int i0 = chunkOf(std::max(a, 0)), i1 = chunkOf(b - 1);
            for (int i = i0; i <= i1; ++i)
                for (int c : cons[i]) { addUnique(ring[k], c); if (posC[c] >= posF[k]) ok = false; }   // later item: deadlock
        }
        if (!ok) {
            if (ringSlots > 0) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_pipeline: ring too small for this mesh band width / lag");
            fullArrays();
            return;
        }
    }
    auto csr = [&](const std::vector<std::vector<int>>& v, std::vector<int>& off, std::vector<int>& list) {
        off.assign(nCh + 1, 0);
        list.clear();
        for (int k = 0; k < nCh; ++k) { list.insert(list.end(), v[k].begin(), v[k].end()); off[k + 1] = (int)list.size(); }
        if (list.empty()) list.push_back(0);
    };
    std::vector<int> depOff, depList, ringOff, ringList;
    csr(deps, depOff, depList); csr(ring, ringOff, ringList);
    P.mode = 1; P.chunkCells = chunkCells; P.lag = lag; P.ringSlots = R; P.nChunks = nCh; P.epoch = 0;
    s->strideI = (size_t)R; s->strideB = (size_t)std::max(h.nBnd, 1); s->bndOff = 5 * (size_t)R; s->ringSize = R;
    s->Fflux.alloc(5 * (size_t)R + 5 * s->strideB); s->Fflux.zero(g_stream);
    P.cellOff.upload(cellOff, g_stream); P.faceOff.upload(faceOff, g_stream);
    P.depOff.upload(depOff, g_stream); P.depList.upload(depList, g_stream); P.ringOff.upload(ringOff, g_stream); P.ringList.upload(ringList, g_stream);
    P.doneF.alloc(nCh); P.doneC.alloc(nCh); P.queue.alloc(1);
    P.doneF.zero(g_stream); P.doneC.zero(g_stream); P.queue.zero(g_stream);
    // keep the ring resident: L2 persisting carve-out + access-policy window on the library stream
    {
        int dev = 0, maxPersist = 0, maxWindow = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t bytes = 5 * (size_t)R * sizeof(double);
        const bool wrap = (size_t)R < (size_t)nIA;
        if (wrap && maxPersist > 0 && maxWindow > 0 && !getenv("QGD_PIPE_NO_PERSIST")) {
            const size_t carve = std::min((size_t)maxPersist, bytes);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = s->Fflux.p;
            attr.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)maxWindow);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)std::min(bytes, (size_t)maxWindow));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) P.windowSet = true;
            cudaGetLastError();
        }
    }
    QGD_CUDA(cudaStreamSynchronize(g_stream));
}

void runStepsImplicit(qgd_solver* s, int n)
{
    const HostMesh& h = s->mesh->h;
    const size_t nC = h.nCells, nF = h.nFaces;
    qgd_solver::Implicit& I = s->impl;
    const bool multi = s->halo.active && g_nranks > 1;
    if (multi && !s->haloFace.active())
        throw Error(QGD_ERR_STATE, "implicitDiffusion true on an extended sub-mesh needs the face-neighbour lists (qgd_qgdfoam_set_halo_faces)");
    if (!I.built) {
        I.GU0.alloc(9 * nC); I.GU1.alloc(9 * nC); I.old.alloc(4 * nC); I.FT.alloc(3 * nF); I.aU.alloc(nF); I.aE.alloc(nF); I.Fs.alloc(nF);
        I.diagU.alloc(nC); I.bU.alloc(3 * nC); I.diagE.alloc(nC); I.bE.alloc(nC);
        I.diagU.zero(g_stream); I.diagE.zero(g_stream);
        std::vector<double> ones(nC, 1.0), zeros(std::max(h.nInternal, 1), 0.0);
        I.AU.build(h, ones.data(), zeros.data(), I.precond, g_stream, &s->mesh->faceInv);
        I.AE.build(h, ones.data(), zeros.data(), I.precond, g_stream, &s->mesh->faceInv);
        if (multi) s->sw.alloc(I.AU, h.nOwned);
        I.built = true;
    }
    const FaceView fv = s->fvsc->view();
    const SolverView sv = s->sview();
    const BndState bs = s->bview();
    const ImplicitView iv = s->iview();
    const bool adjust = s->desc.adjust_time_step != 0;
    const double tol = s->desc.diff_tolerance, rel = s->desc.diff_rel_tol;
    const int maxIter = s->desc.diff_max_iter > 0 ? s->desc.diff_max_iter : 1000;
    StepHooks hooks;
    PcgHooks ph;
    const bool model5 = (bool)s->v5;
    if (model5)        // varScModel5 reads the p_b the closing correctBoundaryConditions() is about to replace (varScModel5.C:255-263)
        hooks.beforeBndPost = [s] {
            if (s->mesh->h.nBnd)
                QGD_CUDA(cudaMemcpyAsync(s->v5->pOldB.p, s->pNew.p, (size_t)s->mesh->h.nBnd * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
        };
    const StepHooks* hp = (multi || model5) ? &hooks : nullptr;
    if (multi) {
        hooks.midStep = [s] { s->launches += haloExchangeMid(s); };
        hooks.afterGrad = [s, nC](double* G) { s->launches += commExchange(s->haloFace, G, nC, 9, g_stream); };
        if (adjust)
            hooks.beforeDt = [s] {
                StepScalars* sc = s->sc.p;
                QGD_NCCL(g_nccl.GroupStart());
                QGD_NCCL(g_nccl.AllReduce(&sc->coMaxBits, &sc->coMaxBits, 1, ncclDouble, ncclMax, g_comm, g_stream));
                QGD_NCCL(g_nccl.AllReduce(&sc->tauMinBits, &sc->tauMinBits, 1, ncclDouble, ncclMin, g_comm, g_stream));
                QGD_NCCL(g_nccl.GroupEnd());
            };
        ph.exchange = [s](double* vec, cudaStream_t cs) { s->launches += commExchange(s->haloFace, vec, 0, 1, cs); };
        ph.allreduceSum = [](double* dev, int count, cudaStream_t cs) { commAllReduce(dev, count, COMM_SUM, cs); };
    }
    // one linear solve: the cooperative persistent kernel on one GPU, the stepwise solver with exchange / all-reduce on sub-meshes
    auto solveSys = [&](PcgMatrix& A, double* b, double* x, int slot) {
        A.bExternal = b; A.xExternal = x;
        if (multi) {
            PcgResult r;
            s->launches += s->sw.solve(A, b, x, tol, rel, maxIter, I.precond, g_stream, &ph, &r);
            QGD_CUDA(cudaMemcpyAsync(s->stage.p + slot * 4, &r, sizeof(PcgResult), cudaMemcpyHostToDevice, g_stream));
            QGD_CUDA(cudaStreamSynchronize(g_stream));
        } else {
            s->launches += A.solve(tol, rel, maxIter, g_stream);
            QGD_CUDA(cudaMemcpyAsync(s->stage.p + slot * 4, A.out.p, sizeof(PcgResult), cudaMemcpyDeviceToDevice, g_stream));
        }
    };
    for (int i = 0; i < n; ++i) {
        if (s->k.model == 1) s->k.tauMode = s->stepsDone == 0 ? 2 : 1;
        ++s->stepsDone;
        s->launches += launchImplicitPhase(g_stream, 0, s->k, fv, sv, bs, iv, s->anyQgdFlux, s->gridFaces, adjust, hp);
        I.AU.refresh(iv.aU, iv.diagU, g_stream);
        for (int j = 0; j < 3; ++j) solveSys(I.AU, iv.bU + j * nC, s->S.p + (1 + j) * nC, j);      // QGDUEqn.H:65-68, segregated components
        if (multi) s->launches += commExchange(s->haloFace, s->S.p + nC, nC, 3, g_stream);          // solved U of the face neighbours
        s->launches += 2 + launchImplicitPhase(g_stream, 1, s->k, fv, sv, bs, iv, s->anyQgdFlux, s->gridFaces, adjust, hp);
        I.AE.refresh(iv.aE, iv.diagE, g_stream);
        s->launches += 2;
        solveSys(I.AE, iv.bE, s->S.p + 4 * nC, 3);                                                 // QGDEEqn.H:55-61
        if (model5)    // the pressure of the old step (p is rewritten by the closing cell kernel of phase 2)
            QGD_CUDA(cudaMemcpyAsync(s->v5->pOld.p, s->S.p + 5 * nC, nC * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
        s->launches += launchImplicitPhase(g_stream, 2, s->k, fv, sv, bs, iv, s->anyQgdFlux, s->gridFaces, adjust, hp);
        if (model5) s->launches += s->v5->correct(s->v5view(), g_stream);      // varScModel5::correct on the closed state
        if (multi) s->launches += haloExchange(s);
    }
    QGD_CUDA(cudaGetLastError());
}

static int h_nB(const qgd_solver* s) { return s->mesh->h.nBnd; }

// varScModel5 at start-up: QGDCoeffs::correct runs twice before the first step - in the thermo constructor
// (hePsiQGDThermo ctor -> calculate()) and in thermo.correct() (QGDFoam/createFields.H:8) - each time relaxing ScQGD from the
// dictionary value towards the sensor and smoothing it.  p is the field as read, p_b the boundary value k_init_bnd left in pNew.
static void varSc5Init(qgd_solver* s)
{
    const HostMesh& h = s->mesh->h;
    VarSc5Device& m = *s->v5;
    std::vector<double> sc0(h.nCells, s->desc.ScQGD), scb(h.nBnd + 1, s->desc.ScQGD);     // varScModel5.C:76-80
    QGD_CUDA(cudaMemcpyAsync(s->scVar.p, sc0.data(), (size_t)h.nCells * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    QGD_CUDA(cudaMemcpyAsync(m.ScB.p, scb.data(), scb.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    QGD_CUDA(cudaMemcpyAsync(m.pOld.p, s->S.p + 5 * (size_t)h.nCells, (size_t)h.nCells * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    if (h.nBnd) QGD_CUDA(cudaMemcpyAsync(m.pOldB.p, s->pNew.p, (size_t)h.nBnd * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    QGD_CUDA(cudaStreamSynchronize(g_stream));          // sc0 / scb are host temporaries
    const VarSc5View v = s->v5view();
    s->launches += m.correct(v, g_stream);
    s->launches += m.correct(v, g_stream);
}

void runSteps(qgd_solver* s, int n)
{
    if (s->k.implicit) {
        if (s->stage.n < 16) s->stage.alloc(QGD_STATE_DOUBLES_PER_CELL * (size_t)s->mesh->h.nCells + 16);
        runStepsImplicit(s, n);
        return;
    }
    const FaceView fv = s->fvsc->view();
    const SolverView sv = s->sview();
    const BndState bs = s->bview();
    StepHooks hooks;
    const bool multi = s->halo.active && g_nranks > 1;
    const bool model5 = (bool)s->v5;
    if (model5)        // varScModel5 reads the p_b the closing correctBoundaryConditions() is about to replace (varScModel5.C:255-263)
        hooks.beforeBndPost = [s] {
            if (s->mesh->h.nBnd)
                QGD_CUDA(cudaMemcpyAsync(s->v5->pOldB.p, s->pNew.p, (size_t)s->mesh->h.nBnd * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
        };
    if (multi) {
        hooks.midStep = [s] { s->launches += haloExchangeMid(s); };
        if (s->desc.adjust_time_step)
            hooks.beforeDt = [s] {
                // Courant max / tau min over all ranks, on the device (gMax/gMin of QGDCourantNo.H:50, setDeltaT-QGDQHD.H:46)
                StepScalars* sc = s->sc.p;
                QGD_NCCL(g_nccl.GroupStart());
                QGD_NCCL(g_nccl.AllReduce(&sc->coMaxBits, &sc->coMaxBits, 1, ncclDouble, ncclMax, g_comm, g_stream));
                QGD_NCCL(g_nccl.AllReduce(&sc->tauMinBits, &sc->tauMinBits, 1, ncclDouble, ncclMin, g_comm, g_stream));
                QGD_NCCL(g_nccl.GroupEnd());
            };
    }
    // overlap the packed exchange with the next step's interior point gather (fixed deltaT, no mid-step exchange)
    const bool overlap = multi && g_commStream && !s->desc.adjust_time_step && !s->anyQgdFlux && s->halo.ptsInterior.n > 0 &&
                         !(getenv("QGD_HALO_OVERLAP") && atoi(getenv("QGD_HALO_OVERLAP")) == 0);
    if (multi)
        hooks.waitHalo = [s](cudaStream_t st) {
            if (s->halo.pending) QGD_CUDA(cudaStreamWaitEvent(st, g_evHalo, 0));
        };
    { const char* v = getenv("QGD_BND_FORK"); g_bndFork = v ? atoi(v) : 1; }
    StepFork fork{g_sideStream, g_evFork[0], g_evFork[1], g_evFork[2], g_evFork[3], g_evFork[4], !multi};
    // measured on one B200: neutral at 256^3 (the boundary chain's time moves into k_points), -3 % per step at 128^3; multi-GPU: opt-in
    // (QGD_BND_FORK=2) until measured - there the side stream can only start after the halo wait
    const bool useFork = (multi ? g_bndFork == 2 : g_bndFork != 0) && g_sideStream && h_nB(s) > 0 && !s->anyQgdFlux && !model5 &&
                         !(s->pipe.mode == 1 && !s->desc.adjust_time_step && !s->k.varSc);
    const WedgeView wedge{(int)s->mesh->h.wedgePts.size(), s->mesh->wedgePts.p, s->mesh->wedgeR.p};
    // ---- opt-in (QGD_STEP_GRAPH=1): the step as a CUDA graph.  One step is captured from the very launch sequence below (side
    // stream included: it joins the capture through evEntry and is joined back before the capture ends) and replayed n times; the
    // time-step control lives on the device (k_dt), so adaptive deltaT needs no re-capture.  Single GPU, two-kernel step form, no
    // per-kernel profiling events; constScPrModel1n takes its first step (tauMode 2) outside the graph.
    int i0 = 0;
    {
        const char* gv = getenv("QGD_STEP_GRAPH");
        const bool wantGraph = gv && atoi(gv) != 0 && !multi && !model5 && !s->profiling && s->pipe.mode == 0 && n >= 2;
        if (wantGraph) {
            if (s->k.model == 1 && s->stepsDone == 0) {             // the first correct() of constScPrModel1n differs (constScPrModel1n.C:102-129)
                s->k.tauMode = 2;
                ++s->stepsDone;
                s->launches += launchStep(g_stream, s->k, fv, sv, bs, s->anyQgdFlux, s->gridFaces, s->desc.adjust_time_step != 0, nullptr,
                                          nullptr, nullptr, 0, useFork ? &fork : nullptr, wedge.n ? &wedge : nullptr);
                if (useFork && fork.postOnSide) QGD_CUDA(cudaStreamWaitEvent(g_stream, fork.evBndPost, 0));
                i0 = 1;
            }
            if (s->k.model == 1) s->k.tauMode = 1;
            warmFaceKernel(s->desc.adjust_time_step != 0);             // kernel attributes / occupancy query: not inside a capture
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            QGD_CUDA(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
            int perStep = 0;
            try {
                perStep = launchStep(g_stream, s->k, fv, sv, bs, s->anyQgdFlux, s->gridFaces, s->desc.adjust_time_step != 0, nullptr,
                                     nullptr, nullptr, 0, useFork ? &fork : nullptr, wedge.n ? &wedge : nullptr);
                if (useFork && fork.postOnSide) QGD_CUDA(cudaStreamWaitEvent(g_stream, fork.evBndPost, 0));
            } catch (...) {
                cudaStreamEndCapture(g_stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            QGD_CUDA(cudaStreamEndCapture(g_stream, &graph));
            cudaError_t ge = cudaGraphInstantiate(&exec, graph, 0);
            if (ge != cudaSuccess) { cudaGraphDestroy(graph); QGD_CUDA(ge); }
            for (int i = i0; i < n; ++i) {
                ge = cudaGraphLaunch(exec, g_stream);
                if (ge != cudaSuccess) break;
                ++s->stepsDone;
                s->launches += perStep;
            }
            if (ge == cudaSuccess) ge = cudaStreamSynchronize(g_stream);     // the executable graph must outlive its launches
            cudaGraphExecDestroy(exec);
            cudaGraphDestroy(graph);
            QGD_CUDA(ge);
            QGD_CUDA(cudaGetLastError());
            s->graphSteps += n - i0;
            return;
        }
    }
    for (int i = 0; i < n; ++i) {
        cudaEvent_t* ev = nullptr;
        if (s->profiling) {
            if (s->eventsUsed + 6 > s->events.size())
                for (int j = 0; j < 6; ++j) { cudaEvent_t e; QGD_CUDA(cudaEventCreate(&e)); s->events.push_back(e); }
            ev = &s->events[s->eventsUsed];
            s->eventsUsed += 6;
        }
        if (s->k.model == 1) s->k.tauMode = s->stepsDone == 0 ? 2 : 1;      // constScPrModel1n.C:102-129
        ++s->stepsDone;
        PipeView pv;
        const bool usePipe = s->pipe.mode == 1 && !s->desc.adjust_time_step;
        if (usePipe) { ++s->pipe.epoch; pv = s->pview(); }
        if (model5)    // the pressure of the old step: thermo.correct() runs before p = rho/psi (QGDFoam.C:149-154)
            QGD_CUDA(cudaMemcpyAsync(s->v5->pOld.p, s->S.p + 5 * (size_t)s->mesh->h.nCells, (size_t)s->mesh->h.nCells * sizeof(double),
                                     cudaMemcpyDeviceToDevice, g_stream));
        s->launches += launchStep(g_stream, s->k, fv, sv, bs, s->anyQgdFlux, s->gridFaces, s->desc.adjust_time_step != 0, ev,
                                  (multi || model5) ? &hooks : nullptr, usePipe ? &pv : nullptr, s->pipe.grid, useFork ? &fork : nullptr,
                                  wedge.n ? &wedge : nullptr);
        if (model5) s->launches += s->v5->correct(s->v5view(), g_stream);      // varScModel5::correct on the closed state
        s->halo.pending = false;
        if (multi) {
            if (overlap) {
                QGD_CUDA(cudaEventRecord(g_evStep, g_stream));
                QGD_CUDA(cudaStreamWaitEvent(g_commStream, g_evStep, 0));
                s->launches += haloExchange(s, g_commStream);
                QGD_CUDA(cudaEventRecord(g_evHalo, g_commStream));
                s->halo.pending = true;
            } else s->launches += haloExchange(s);
        }
    }
    if (multi && s->halo.pending) { QGD_CUDA(cudaStreamWaitEvent(g_stream, g_evHalo, 0)); s->halo.pending = false; }
    if (useFork && n > 0 && fork.postOnSide) QGD_CUDA(cudaStreamWaitEvent(g_stream, fork.evBndPost, 0));   // join: boundary state closed
    QGD_CUDA(cudaGetLastError());
}

} // namespace

// ---- multi-GPU plumbing shared with qgd_qhd.cu
int qgd::commRanks() { return g_comm ? g_nranks : 1; }

void qgd::HaloLists::set(int nn, const int* nbrRank, const int* sOff, const int* sIds, const int* rOff, const int* rIds, int nLocal,
                         int nOwned, int maxComponents, cudaStream_t st)
{
    if (nn < 0 || (nn > 0 && (!nbrRank || !sOff || !sIds || !rOff || !rIds))) throw Error(QGD_ERR_INVALID, "halo lists: bad arguments");
    if (nn > 0 && !g_comm) throw Error(QGD_ERR_STATE, "halo lists: call qgd_comm_init first");
    nbr.assign(nbrRank, nbrRank + nn);
    sendOff.assign(sOff, sOff + nn + 1); recvOff.assign(rOff, rOff + nn + 1);
    for (int i = 0; i < sendOff[nn]; ++i) if (sIds[i] < 0 || sIds[i] >= nOwned) throw Error(QGD_ERR_INVALID, "halo lists: send cell is not an owned cell");
    for (int i = 0; i < recvOff[nn]; ++i) if (rIds[i] < nOwned || rIds[i] >= nLocal) throw Error(QGD_ERR_INVALID, "halo lists: recv cell is not a halo cell");
    auto up = [&](DevBuf<int>& d, const int* p, int n) { std::vector<int> v(p, p + n); if (v.empty()) v.push_back(0); d.upload(v, st); };
    up(sendIds, sIds, sendOff[nn]); up(recvIds, rIds, recvOff[nn]);
    maxComp = maxComponents;
    sendBuf.alloc((size_t)maxComp * sendOff[nn] + 1); recvBuf.alloc((size_t)maxComp * recvOff[nn] + 1);
    offDev.upload(sendOff, st); roffDev.upload(recvOff, st);
}

int qgd::commExchange(HaloLists& h, double* base, size_t stride, int nComp, cudaStream_t st)
{
    if (!h.active()) return 0;
    if (nComp > h.maxComp) throw Error(QGD_ERR_INVALID, "commExchange: more components than the halo buffers hold");
    const int nn = (int)h.nbr.size();
    int launches = 0;
    const int nS = h.sendOff[nn], nR = h.recvOff[nn];
    if (nS) { k_halo_fields<<<(nS + 255) / 256, 256, 0, st>>>(nn, h.offDev.p, h.sendIds.p, base, stride, nComp, h.maxComp, h.sendBuf.p, 1); ++launches; }
    QGD_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < nn; ++k) {
        const size_t ns = (size_t)nComp * (h.sendOff[k + 1] - h.sendOff[k]), nr = (size_t)nComp * (h.recvOff[k + 1] - h.recvOff[k]);
        if (ns) QGD_NCCL(g_nccl.Send(h.sendBuf.p + (size_t)h.maxComp * h.sendOff[k], ns, ncclDouble, h.nbr[k], g_comm, st));
        if (nr) QGD_NCCL(g_nccl.Recv(h.recvBuf.p + (size_t)h.maxComp * h.recvOff[k], nr, ncclDouble, h.nbr[k], g_comm, st));
    }
    QGD_NCCL(g_nccl.GroupEnd());
    if (nR) { k_halo_fields<<<(nR + 255) / 256, 256, 0, st>>>(nn, h.roffDev.p, h.recvIds.p, base, stride, nComp, h.maxComp, h.recvBuf.p, 0); ++launches; }
    QGD_CUDA(cudaGetLastError());
    return launches;
}

void qgd::commAllReduce(double* dev, int count, CommOp op, cudaStream_t st)
{
    if (!g_comm || g_nranks <= 1) return;
    QGD_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, op == COMM_SUM ? ncclSum : (op == COMM_MAX ? ncclMax : ncclMin), g_comm, st));
}

void qgd::fvscBuild(qgd_fvsc& op, qgd_mesh* mesh, const std::string& name)
{
    // fvsc.C:47-85 (fvscOpName checks) + fvscStencil.C:59-95 (New)
    if (!inTable(kFvscTable, name))
        throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown Model type " + name + "\n\nValid model types are:\n" + toc(kFvscTable));
    if ((name == "leastSquares" || name == "leastSquaresOpt")) {
        if (mesh->h.nD == 3) throw Error(QGD_ERR_INVALID, "Can't use leastSquares or leastSquaresOpt in 3D case.");
        // extended sub-meshes: every cell sharing a point with a face of an owned cell is present (vertex-ring halo), so the stencil of
        // every face this rank evaluates is complete (extendedFaceStencilFindNeighbours.C:41-86); faces owned by halo cells are never computed
    }
    if (name == "GaussVolPoint") {               // fvsc.C:65-82: wedge patches + prism cells are rejected
        const HostMesh& h = mesh->h;
        bool wedge = false;
        for (int pk : h.patchKind) wedge = wedge || pk == QGD_PATCH_WEDGE;
        if (wedge)
            for (int c = 0; c < h.nCells; ++c) {     // prismMatcher [OF-v2312]: 5 faces, two triangles and three quads
                if (h.cfOff[c + 1] - h.cfOff[c] != 5) continue;
                int tri = 0, quad = 0;
                for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q) {
                    const int f = h.cfEnc[q] >> 1, nv = h.faceOff[f + 1] - h.faceOff[f];
                    tri += nv == 3; quad += nv == 4;
                }
                if (tri == 2 && quad == 3)
                    throw Error(QGD_ERR_INVALID, "GaussVolPoint scheme does not support solving axisymmetric cases with wedge BC and prism cells.\n"
                                                 "Try to set leastSquares scheme.");
            }
    }
    op.mesh = mesh;
    op.lsq = (name == "leastSquares" || name == "leastSquaresOpt");
    op.reduced = (name == "reduced") || op.lsq;      // boundary faces nf*snGrad, no point values (extendedFaceStencilScalarGrad.C:86-109)
    std::vector<int> vtx, flags;
    std::vector<double> G, hd;
    mesh->h.buildFaceRecords(op.reduced, vtx, flags, G, hd);
    const int nF = mesh->h.nFaces;
    if (op.lsq)      // leastSquares leaves the faces of constraint patches at zero (extendedFaceStencilScalarGrad.C:86-109: empty, wedge, coupled ...)
        for (int b = 0; b < mesh->h.nBnd; ++b)
            if (mesh->h.patchKind[mesh->h.bfacePatch[b]] == QGD_PATCH_WEDGE || mesh->h.patchKind[mesh->h.bfacePatch[b]] == QGD_PATCH_SYMMETRY_PLANE)
                for (int k = 6; k < 9; ++k) G[(size_t)k * nF + mesh->h.nInternal + b] = 0.0;
    const std::vector<int>& perm = mesh->facePerm;
    std::vector<int4> v4(nF);
    std::vector<int> fl(nF);
    const size_t fs = (size_t)mesh->faceStride;
    // device record: G1[3], G2[3] and the scalar gpS with GP = gpS * Sf.  GP is parallel to the face area vector in every scheme
    // (quad: (e1 x e2)/D with the diagonals e1, e2; tri: A_f/(3 vt); 2D: the in-plane normal of v13; 1D / reduced / polygon faces:
    // -nf delta), so 7 doubles per face carry the 9 of the full record (GaussVolPointBase3D.C:186-229,346-389; ...2D.C:122-169)
    std::vector<double> Gp(7 * fs, 0.0);
    for (int f = 0; f < nF; ++f) {
        const size_t o = perm[f];
        v4[f] = make_int4(vtx[4 * o], vtx[4 * o + 1], vtx[4 * o + 2], vtx[4 * o + 3]);
        fl[f] = flags[o];
        for (int k = 0; k < 6; ++k) Gp[(size_t)k * fs + f] = G[(size_t)k * nF + o];
        const double* S = &mesh->h.Sf[3 * o];
        const double gP[3] = {G[(size_t)6 * nF + o], G[(size_t)7 * nF + o], G[(size_t)8 * nF + o]};
        const double SS = S[0] * S[0] + S[1] * S[1] + S[2] * S[2];
        const double gS = SS > 0.0 ? (gP[0] * S[0] + gP[1] * S[1] + gP[2] * S[2]) / SS : 0.0;
        double dev = 0.0, nrm = 0.0;
        for (int i = 0; i < 3; ++i) { dev = std::max(dev, std::fabs(gP[i] - gS * S[i])); nrm = std::max(nrm, std::fabs(gP[i])); }
        if (dev > 1e-9 * nrm)
            throw Error(QGD_ERR_INVALID, "fvsc: face " + std::to_string(o) + " has a gradient record whose cell-difference vector is not "
                                         "parallel to the face area vector (degenerate face geometry)");
        Gp[(size_t)6 * fs + f] = gS;
    }
    op.flagsUniform = -1;
    if (mesh->nIActive > 0) {
        op.flagsUniform = fl[0];
        for (int f = 1; f < mesh->nIActive; ++f) if (fl[f] != fl[0]) { op.flagsUniform = -1; break; }
    }
    if (op.lsq) {
        // internal faces: least-squares cell stencil; the G record keeps the nf*snGrad fallback of degenerate faces
        int W = 0;
        std::vector<int> cells;
        std::vector<double> coef;
        std::vector<char> deg;
        mesh->h.buildLeastSquares(name == "leastSquaresOpt", W, cells, coef, deg);
        const int nI = mesh->h.nInternal;
        if (name == "leastSquares")                     // leastSquaresStencil.C:63-132: user faceSet of faces forced to nf*snGrad
            for (int f : mesh->h.forcedDegFaces) if (f >= 0 && f < nI) deg[f] = 1;
        const size_t nIs = (size_t)std::max(nI, 1);
        std::vector<int> cd((size_t)W * nIs, 0);
        std::vector<double> kd((size_t)W * 3 * nIs, 0.0);
        for (int f = 0; f < nI; ++f) {
            const size_t o = perm[f];
            if (!deg[o]) fl[f] |= FF_LSQ;
            for (int j = 0; j < W; ++j) {
                cd[(size_t)j * nIs + f] = cells[(size_t)j * nIs + o];
                for (int q = 0; q < 3; ++q) kd[((size_t)j * 3 + q) * nIs + f] = coef[((size_t)j * 3 + q) * nIs + o];
            }
        }
        op.lsqW = W;
        op.lsqCells.upload(cd, g_stream); op.lsqCoef.upload(kd, g_stream);
    }
    op.vtx.upload(v4, g_stream); op.flags.upload(fl, g_stream); op.G.upload(Gp, g_stream); op.halfDist.upload(hd, g_stream);
}


// ============================================================================ C ABI
extern "C" {

int qgd_version(void) { return 100; }

const char* qgd_last_error(void) { return g_lastError.c_str(); }

int qgd_init(int device)
{
    return guarded([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw Error(QGD_ERR_CUDA, std::string("qgd_init: no CUDA device available (") + cudaGetErrorString(e) +
                                          "); libqgd_b200 has no CPU fallback");
        if (device < 0 || device >= n) throw Error(QGD_ERR_INVALID, "qgd_init: device index out of range");
        QGD_CUDA(cudaSetDevice(device));
        if (!g_stream) QGD_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
        if (!g_sideStream) {
            int prLo = 0, prHi = 0;
            QGD_CUDA(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
            QGD_CUDA(cudaStreamCreateWithPriority(&g_sideStream, cudaStreamNonBlocking, prHi));
            for (cudaEvent_t& e : g_evFork) QGD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (const char* v = getenv("QGD_BND_FORK")) g_bndFork = atoi(v);
        if (!g_ev0) { QGD_CUDA(cudaEventCreate(&g_ev0)); QGD_CUDA(cudaEventCreate(&g_ev1)); }
        g_initialised = true;
    });
}

int qgd_device_synchronize(void)
{
    return guarded([&] { requireInit(); QGD_CUDA(cudaStreamSynchronize(g_stream)); });
}

int qgd_timer_begin(void) { return guarded([&] { requireInit(); QGD_CUDA(cudaEventRecord(g_ev0, g_stream)); }); }
int qgd_timer_end(float* ms)
{
    return guarded([&] {
        requireInit();
        QGD_CUDA(cudaEventRecord(g_ev1, g_stream));
        QGD_CUDA(cudaEventSynchronize(g_ev1));
        QGD_CUDA(cudaEventElapsedTime(ms, g_ev0, g_ev1));
    });
}

// ---------------------------------------------------------------- mesh
int qgd_mesh_create(const qgd_mesh_desc* desc, qgd_mesh** out)
{
    return guarded([&] {
        requireInit();
        if (!desc || !out) throw Error(QGD_ERR_INVALID, "qgd_mesh_create: null argument");
        std::unique_ptr<qgd_mesh> m(new qgd_mesh());
        m->h.build(*desc);
        const HostMesh& h = m->h;
        const int nF = h.nFaces, nI = h.nInternal;
        // ---- device face order (env QGD_FACE_TILE: owners per tile, 0 keeps the polyMesh order)
        int tile = 32;
        if (const char* v = getenv("QGD_FACE_TILE")) tile = atoi(v);
        m->facePerm.resize(nF);
        for (int f = 0; f < nF; ++f) m->facePerm[f] = f;
        if (tile > 0 && nI > 0) {
            // rank of each internal face among the faces owned by its owner (faces are owner-sorted: upper-triangular order)
            std::vector<int> rank(nI, 0);
            for (int f = 1; f < nI; ++f) rank[f] = (h.owner[f] == h.owner[f - 1]) ? rank[f - 1] + 1 : 0;
            const int nOwn = h.nOwned;
            std::stable_sort(m->facePerm.begin(), m->facePerm.begin() + nI, [&](int a, int b) {
                const bool ha = h.owner[a] >= nOwn, hb = h.owner[b] >= nOwn;      // halo-owned faces last (never computed)
                if (ha != hb) return hb;
                const int ta = h.owner[a] / tile, tb = h.owner[b] / tile;
                if (ta != tb) return ta < tb;
                if (rank[a] != rank[b]) return rank[a] < rank[b];
                return h.owner[a] < h.owner[b];
            });
        }
        m->nIActive = 0;
        for (int f = 0; f < nI; ++f) if (h.owner[m->facePerm[f]] < h.nOwned) m->nIActive = f + 1;
        m->faceInv.resize(nF);
        for (int f = 0; f < nF; ++f) m->faceInv[m->facePerm[f]] = f;
        const std::vector<int>& perm = m->facePerm;
        auto permI = [&](const std::vector<int>& v, int n) { std::vector<int> o(n); for (int f = 0; f < n; ++f) o[f] = v[perm[f]]; return o; };
        auto permD = [&](const std::vector<double>& v) { std::vector<double> o(nF); for (int f = 0; f < nF; ++f) o[f] = v[perm[f]]; return o; };
        m->facePermDev.upload(perm, g_stream);
        m->owner.upload(permI(h.owner, nF), g_stream); m->neighbour.upload(permI(h.neighbour, nI), g_stream);
        // ---- cell -> faces ELL (+ tail), device face ids, polyMesh ascending order inside each row
        {
            int maxRow = 0;
            for (int c = 0; c < h.nCells; ++c) maxRow = std::max(maxRow, h.cfOff[c + 1] - h.cfOff[c]);
            int W = maxRow <= 4 ? 4 : (maxRow <= 6 ? 6 : 8);
            if (const char* v = getenv("QGD_ELL_MAXW")) W = std::min(W, std::max(4, atoi(v)) <= 4 ? 4 : (atoi(v) <= 6 ? 6 : 8));   // test hook: force CSR tails
            m->cfEllW = W;
            std::vector<int> ell((size_t)W * h.nCells, -1), tailOff(h.nCells + 1, 0), tail;
            for (int c = 0; c < h.nCells; ++c) {
                int j = 0;
                for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q, ++j) {
                    const int enc = (m->faceInv[h.cfEnc[q] >> 1] << 1) | (h.cfEnc[q] & 1);
                    if (j < W) ell[(size_t)j * h.nCells + c] = enc; else tail.push_back(enc);
                }
                tailOff[c + 1] = (int)tail.size();
            }
            if (tail.empty()) tail.push_back(-1);
            m->cfEll.upload(ell, g_stream); m->cfTailOff.upload(tailOff, g_stream); m->cfTailEnc.upload(tail, g_stream);
        }
        // ---- point -> cells ELL (+ tail)
        {
            int maxRow = 0;
            for (int p = 0; p < h.nPoints; ++p) maxRow = std::max(maxRow, h.pcOff[p + 1] - h.pcOff[p]);
            int W = maxRow <= 4 ? 4 : (maxRow <= 6 ? 6 : 8);
            if (const char* v = getenv("QGD_ELL_MAXW")) W = std::min(W, std::max(4, atoi(v)) <= 4 ? 4 : (atoi(v) <= 6 ? 6 : 8));   // test hook: force CSR tails
            m->pcEllW = W;
            std::vector<int> ell((size_t)W * h.nPoints, 0), cnt(h.nPoints, 0), tailOff(h.nPoints + 1, 0), tailC;
            std::vector<double> ellW((size_t)W * h.nPoints, 0.0), tailW;
            for (int p = 0; p < h.nPoints; ++p) {
                cnt[p] = h.pcOff[p + 1] - h.pcOff[p];
                const int first = cnt[p] ? h.pcCell[h.pcOff[p]] : 0;
                for (int j = 0; j < W; ++j) ell[(size_t)j * h.nPoints + p] = first;
                int j = 0;
                for (int q = h.pcOff[p]; q < h.pcOff[p + 1]; ++q, ++j) {
                    if (j < W) { ell[(size_t)j * h.nPoints + p] = h.pcCell[q]; ellW[(size_t)j * h.nPoints + p] = h.pcW[q]; }
                    else { tailC.push_back(h.pcCell[q]); tailW.push_back(h.pcW[q]); }
                }
                tailOff[p + 1] = (int)tailC.size();
            }
            if (tailC.empty()) { tailC.push_back(0); tailW.push_back(0.0); }
            m->pcEll.upload(ell, g_stream); m->pcEllWt.upload(ellW, g_stream); m->pcCount.upload(cnt, g_stream);
            m->pcTailOff.upload(tailOff, g_stream); m->pcTailCell.upload(tailC, g_stream); m->pcTailW.upload(tailW, g_stream);
        }
        m->patchPoints.upload(h.patchPoints, g_stream); m->ppOff.upload(h.ppOff, g_stream);
        m->ppFace.upload(h.ppFace, g_stream); m->ppW.upload(h.ppW, g_stream);
        std::vector<int> bk(h.nBnd);
        for (int b = 0; b < h.nBnd; ++b) bk[b] = h.patchKind[h.bfacePatch[b]];
        m->bfaceKind.upload(bk, g_stream);
        m->faceStride = (nF + 15) & ~15;
        std::vector<double> sfSoA(3 * (size_t)m->faceStride, 0.0);
        for (int f = 0; f < nF; ++f)
            for (int d = 0; d < 3; ++d) sfSoA[(size_t)d * m->faceStride + f] = h.Sf[3 * (size_t)perm[f] + d];
        m->Sf.upload(sfSoA, g_stream);
        m->magSf.upload(permD(h.magSf), g_stream); m->w.upload(permD(h.w), g_stream); m->dC.upload(permD(h.dC), g_stream);
        m->ndC.upload(permD(h.ndC), g_stream); m->V.upload(h.V, g_stream);
        m->hQGDf.upload(permD(h.hQGDf), g_stream); m->hQGD.upload(h.hQGD, g_stream);
        if (!h.wedgePts.empty()) { m->wedgePts.upload(h.wedgePts, g_stream); m->wedgeR.upload(h.wedgeR, g_stream); }
        *out = m.release();
    });
}

int qgd_mesh_set_degenerate_stencil_faces(qgd_mesh* m, const int* faces, int n)
{
    return guarded([&] {
        requireInit();
        if (!m || n < 0 || (n > 0 && !faces)) throw Error(QGD_ERR_INVALID, "qgd_mesh_set_degenerate_stencil_faces: bad argument");
        for (int i = 0; i < n; ++i)
            if (faces[i] < 0 || faces[i] >= m->h.nFaces) throw Error(QGD_ERR_INVALID, "qgd_mesh_set_degenerate_stencil_faces: face id out of range");
        m->h.forcedDegFaces.assign(faces, faces + n);
    });
}

int qgd_mesh_set_pcg_blocks(qgd_mesh* m, const int* cell_block)
{
    return guarded([&] {
        if (!m) throw Error(QGD_ERR_INVALID, "qgd_mesh_set_pcg_blocks: null mesh");
        if (!cell_block) { m->h.pcgBlock.clear(); return; }
        const int n = m->h.nCells;
        std::vector<int> b(cell_block, cell_block + n);
        for (int c = 0; c < n; ++c) {
            if (c >= m->h.nOwned) b[c] = -1;
            else if (b[c] < 0) throw Error(QGD_ERR_INVALID, "qgd_mesh_set_pcg_blocks: negative block id");
        }
        m->h.pcgBlock.swap(b);
    });
}

int qgd_mesh_make_pcg_blocks(qgd_mesh* m, int target_cells, int* n_blocks)
{
    return guarded([&] {
        if (!m || target_cells < 1) throw Error(QGD_ERR_INVALID, "qgd_mesh_make_pcg_blocks: bad argument");
        m->h.makePcgBlocks(target_cells);
        int nb = 0;
        for (int v : m->h.pcgBlock) nb = std::max(nb, v + 1);
        if (n_blocks) *n_blocks = nb;
    });
}

int qgd_mesh_get_pcg_blocks(qgd_mesh* m, int* cell_block)
{
    return guarded([&] {
        if (!m || !cell_block) throw Error(QGD_ERR_INVALID, "qgd_mesh_get_pcg_blocks: null argument");
        for (int c = 0; c < m->h.nCells; ++c) cell_block[c] = m->h.pcgBlock.empty() ? -1 : m->h.pcgBlock[c];
    });
}

int qgd_mesh_destroy(qgd_mesh* mesh) { return guarded([&] { delete mesh; }); }

int qgd_mesh_get(qgd_mesh* mesh, int what, double* out)
{
    return guarded([&] {
        if (!mesh || !out) throw Error(QGD_ERR_INVALID, "qgd_mesh_get: null argument");
        if (what == 0) std::copy(mesh->h.hQGDf.begin(), mesh->h.hQGDf.end(), out);
        else if (what == 1) { d2h(out, mesh->hQGD.p, mesh->hQGD.n); }
        else throw Error(QGD_ERR_INVALID, "qgd_mesh_get: unknown field id");
        QGD_CUDA(cudaStreamSynchronize(g_stream));
    });
}

// ---------------------------------------------------------------- fvsc
int qgd_fvsc_create(qgd_mesh* mesh, const char* scheme_name, qgd_fvsc** out)
{
    return guarded([&] {
        requireInit();
        if (!mesh || !scheme_name || !out) throw Error(QGD_ERR_INVALID, "qgd_fvsc_create: null argument");
        std::unique_ptr<qgd_fvsc> op(new qgd_fvsc());
        fvscBuild(*op, mesh, scheme_name);
        *out = op.release();
    });
}

int qgd_fvsc_destroy(qgd_fvsc* op) { return guarded([&] { delete op; }); }

static void fvscApply(qgd_fvsc* op, bool isGrad, int K, const double* cell, const double* bnd, const double* bsg,
                      const double* nbr, double* out)
{
    requireInit();
    if (!op || !cell || !out) throw Error(QGD_ERR_INVALID, "fvsc operator: null argument");
    const HostMesh& h = op->mesh->h;
    if (h.nBnd && (!bnd || !bsg)) throw Error(QGD_ERR_INVALID, "fvsc operator: boundary values and snGrad are required");
    if (isGrad && K != 1 && K != 3) throw Error(QGD_ERR_INVALID, "fvsc::grad: ncmpt must be 1 or 3");
    if (!isGrad && K != 3 && K != 9) throw Error(QGD_ERR_INVALID, "fvsc::div: ncmpt must be 3 or 9");
    const size_t outK = isGrad ? 3 * (size_t)K : (size_t)K / 3;
    h2d(op->dCell, cell, (size_t)h.nCells * K);
    h2d(op->dBnd, bnd, (size_t)h.nBnd * K);
    h2d(op->dBsg, bsg, (size_t)h.nBnd * K);
    if (nbr) h2d(op->dNbr, nbr, (size_t)h.nBnd * K);
    if (op->dPts.n < (size_t)h.nPoints * K) op->dPts.alloc((size_t)h.nPoints * K);
    if (op->dOut.n < (size_t)h.nFaces * outK) op->dOut.alloc((size_t)h.nFaces * outK);
    const FaceView fv = op->view();
    const bool needPoints = !op->reduced && h.nD > 1;
    if (needPoints) launchPointGather(g_stream, K, *op->mesh, op->dCell.p, op->dBnd.p, op->dPts.p);
    if (isGrad) launchFvscGrad(g_stream, K, fv, op->dCell.p, op->dPts.p, op->dBnd.p, op->dBsg.p, nbr ? op->dNbr.p : nullptr, op->dOut.p);
    else launchFvscDiv(g_stream, K, fv, op->dCell.p, op->dPts.p, op->dBnd.p, op->dBsg.p, nbr ? op->dNbr.p : nullptr, op->dOut.p);
    d2h(out, op->dOut.p, (size_t)h.nFaces * outK);
    QGD_CUDA(cudaStreamSynchronize(g_stream));
}

int qgd_fvsc_grad(qgd_fvsc* op, int ncmpt, const double* cell, const double* bnd, const double* bsg, const double* nbr, double* out)
{
    return guarded([&] { fvscApply(op, true, ncmpt, cell, bnd, bsg, nbr, out); });
}
int qgd_fvsc_div(qgd_fvsc* op, int ncmpt, const double* cell, const double* bnd, const double* bsg, const double* nbr, double* out)
{
    return guarded([&] { fvscApply(op, false, ncmpt, cell, bnd, bsg, nbr, out); });
}

// ---------------------------------------------------------------- QGDFoam
int qgd_qgdfoam_create(qgd_mesh* mesh, const qgd_qgdfoam_desc* d, qgd_solver** out)
{
    return guarded([&] {
        requireInit();
        if (!mesh || !d || !out || !d->fvsc_scheme || !d->qgd_coeffs_model) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_create: null argument");
        const std::string model = d->qgd_coeffs_model;
        if (!inTable(kCoeffsTable, model))     // QGDCoeffs.C:70-79
            throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown QGD coeffs evaluation approach type " + model +
                                                   "\n\nValid model types are:\n" + toc(kCoeffsTable));
        if (model != "constScPrModel1" && model != "constScPrModel1n" && model != "constScPrModel2" && model != "varScModel5" &&
            model != "varScModel6" && model != "varScModel7")
            throw Error(QGD_ERR_UNSUPPORTED, "QGDCoeffs model " + model + " is a QHDFoam model: it is not available in QGDFoam");
        if (model == "varScModel5") {
            if (mesh->h.nOwned != mesh->h.nCells)
                throw Error(QGD_ERR_UNSUPPORTED, "varScModel5 on extended sub-meshes (multi-GPU) is not available: fvc::smooth crosses processor "
                                                 "patches inside FaceCellWave");
            if (!(d->varsc5_smoothCoeff >= 0.0) || !(d->varsc5_maxAspectRatio > 0.0) || !(d->varsc_minSc <= d->varsc_maxSc))
                throw Error(QGD_ERR_INVALID, "varScModel5: smoothCoeff must be >= 0, maxAspectRatio > 0 and minSc <= maxSc");
        }
        int diffPrecond = 2;
        if (d->implicit_diffusion) {
            const std::string pc = d->diff_preconditioner ? d->diff_preconditioner : "DIC";
            if (pc == "DIC") diffPrecond = 2; else if (pc == "diagonal") diffPrecond = 1; else if (pc == "none") diffPrecond = 0;
            else throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown symmetric matrix preconditioner " + pc + "\n\nValid symmetric matrix preconditioners are:\n3\n(\nDIC\ndiagonal\nnone\n)\n");
            if (mesh->h.nOwned != mesh->h.nCells && diffPrecond == 2 && mesh->h.pcgBlock.empty())
                throw Error(QGD_ERR_UNSUPPORTED, "implicitDiffusion true on extended sub-meshes (multi-GPU): DIC is local to a block in a decomposed "
                                                 "run - set DIC blocks first (qgd_mesh_make_pcg_blocks) or use diagonal | none");
        }
        for (int pk : mesh->h.patchKind)
            if (pk == QGD_PATCH_PROCESSOR)
                throw Error(QGD_ERR_UNSUPPORTED, "processor patches: use the multi-GPU entry points");
        if (!(d->delta_t > 0.0)) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_create: deltaT must be positive");
        std::unique_ptr<qgd_solver> s(new qgd_solver());
        s->mesh = mesh;
        s->desc = *d;
        s->desc.diff_preconditioner = nullptr; s->desc.transport_model = nullptr; s->desc.thermo_model = nullptr;
        s->impl.precond = diffPrecond;
        s->fvsc.reset(new qgd_fvsc());
        fvscBuild(*s->fvsc, mesh, d->fvsc_scheme);
        Consts& k = s->k;
        k.R = d->R; k.Cp = d->Cp; k.Cv = d->Cp - d->R; k.Tref = d->Tref; k.Hsref = d->Hsref; k.mu = d->mu; k.Pr = d->Pr;
        k.ScQGD = d->ScQGD; k.PrQGD = d->PrQGD; k.gamma = d->Cp / (d->Cp - d->R);
        // thermoType instantiations of psiQGDThermos.C:65-111: transport const | sutherland | powerLaw, thermo hConst | eConst
        const std::string transport = d->transport_model ? d->transport_model : "const";
        const std::string thermoM = d->thermo_model ? d->thermo_model : "hConst";
        k.transport = 0; k.eConst = 0; k.mu0 = k.T0 = k.kExp = k.rPr = k.As = k.Ts = k.Esref = 0.0;
        if (transport == "powerLaw") {
            if (!(d->T0 > 0.0) || !(d->Pr > 0.0)) throw Error(QGD_ERR_INVALID, "powerLaw transport: T0 and Pr must be positive");
            k.transport = 1; k.mu0 = d->mu0; k.T0 = d->T0; k.kExp = d->k_exp; k.rPr = 1.0 / d->Pr;
        } else if (transport == "sutherland") { k.transport = 2; k.As = d->As; k.Ts = d->Ts; }
        else if (transport != "const")
            throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown transport type " + transport + "\n\nValid transport types of hePsiQGDThermo are:\n3\n(\nconst\npowerLaw\nsutherland\n)\n");
        if (thermoM == "eConst") {
            if (!(d->Cv > 0.0)) throw Error(QGD_ERR_INVALID, "eConst thermo: Cv must be positive");
            if (k.transport != 0) throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown psiQGDThermo type: eConst is instantiated with const transport only (psiQGDThermos.C:89-99)");
            k.eConst = 1; k.Cv = d->Cv; k.Cp = d->Cv + d->R; k.Esref = d->Esref; k.gamma = k.Cp / k.Cv;
        } else if (thermoM != "hConst")
            throw Error(QGD_ERR_UNKNOWN_MODEL, "Unknown thermo type " + thermoM + "\n\nValid thermo types of hePsiQGDThermo are:\n2\n(\neConst\nhConst\n)\n");
        k.alphaEffGamma = d->alpha_eff_gamma_factor; k.energyQuirk = d->energy_ddt_rhoE_quirk; k.reducedScheme = s->fvsc->reduced;
        k.model = (model == "constScPrModel1" || model == "varScModel6" || model == "varScModel7") ? 0 : (model == "constScPrModel1n" ? 1 : 2);
        // varScModel6.C:207-208 / varScModel7.C:173-174: tau as constScPrModel1; ScQGD from the sensor, boundary ScQGD = dict value (clamped by model 7)
        k.varSc = model == "varScModel6" ? 6 : (model == "varScModel7" ? 7 : 0);
        k.cSc1 = d->varsc_cSc1; k.minSc = d->varsc_minSc; k.maxSc = d->varsc_maxSc; k.ScB = d->ScQGD;
        if (k.varSc == 7) {
            if (k.minSc >= 0) k.ScB = std::max(k.ScB, k.minSc);
            if (k.maxSc >= 0) k.ScB = std::min(k.ScB, k.maxSc);
        }
        k.tauMode = 0; k.implicit = d->implicit_diffusion ? 1 : 0;
        const HostMesh& h = mesh->h;
        if (model == "varScModel5") {
            // the ordinary kernels see model 0 with tauQGDf = I(alphaQGD) hQGDf / I(c) (varScModel5.C:204-205); everything that depends
            // on ScQGD is rewritten by the model's own pass after each step (qgd_varsc5.h)
            k.model = 0; k.varSc = 0; k.tauMode = 2;
            s->v5.reset(new VarSc5Device());
            s->v5->create(h, k, d->varsc5_rC, d->varsc_minSc, d->varsc_maxSc, d->ScQGD, d->varsc5_smoothCoeff, d->varsc5_badQualitySc,
                          d->varsc5_maxAspectRatio, g_stream);
            std::vector<double> sc0(h.nCells, d->ScQGD);                    // ScQGD_.primitiveFieldRef() = ScQGD (varScModel5.C:76)
            s->scVar.upload(sc0, g_stream);
        }
        s->S.alloc(16 * (size_t)h.nCells); s->P.alloc(6 * (size_t)h.nPoints);
        s->bA.alloc(h.nBnd); s->bB.alloc(h.nBnd);
        s->aQGD.alloc(h.nCells);
        if (k.varSc) { s->scVar.alloc(h.nCells); s->scVar.zero(g_stream); }
        if (k.model != 0) { s->tauOut.alloc(h.nCells); s->tauOutB.alloc(h.nBnd + 1); s->tauOut.zero(g_stream); s->tauOutB.zero(g_stream); }
        s->psiB.alloc(h.nBnd); s->pGrad.alloc(h.nBnd); s->pNew.alloc(h.nBnd); s->phiw.alloc(h.nBnd);
        s->P.zero(g_stream);
        // default step form: two kernels (measured faster on B200 at 256^3); QGD_PIPELINE=1 opts in
        configurePipeline(s.get(), (getenv("QGD_PIPELINE") && atoi(getenv("QGD_PIPELINE")) && !d->adjust_time_step && !s->v5) ? 1 : 0, 0, -1, 0);
        StepScalars sc{};
        sc.dt = d->delta_t; sc.time = 0.0; sc.coNum = -1.0; sc.coMaxBits = 0ull;
        const double big = DBL_MAX;
        std::memcpy(&sc.tauMinBits, &big, sizeof(double));
        sc.maxCo = d->max_co; sc.maxDeltaT = d->max_delta_t; sc.cTau = d->c_tau; sc.adjust = d->adjust_time_step;
        s->sc.upload(std::vector<StepScalars>(1, sc), g_stream);
        if (const char* v = getenv("QGD_FACE_TMA")) setFaceTma(atoi(v));
        if (const char* v = getenv("QGD_FACE_L2HINT")) setFaceL2Hint(atoi(v));
        s->gridFaces = faceKernelGrid();
        *out = s.release();
    });
}

int qgd_qgdfoam_destroy(qgd_solver* s) { return guarded([&] { delete s; }); }

int qgd_qgdfoam_set_sources(qgd_solver* s, const double* rhoSu, const double* rhoUSu, const double* rhoESu)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_sources: null solver");
        if (!rhoSu && !rhoUSu && !rhoESu) { s->su.release(); return; }
        const size_t n = s->mesh->h.nCells;
        std::vector<double> h(5 * n, 0.0);
        for (size_t c = 0; c < n; ++c) {
            if (rhoSu) h[c] = rhoSu[c];
            if (rhoUSu) for (int j = 0; j < 3; ++j) h[(1 + j) * n + c] = rhoUSu[3 * c + j];
            if (rhoESu) h[4 * n + c] = rhoESu[c];
        }
        QGD_CUDA(cudaStreamSynchronize(g_stream));      // steps in flight still read the previous sources
        s->su.upload(h, g_stream);
    });
}

int qgd_qgdfoam_set_const_sc_cells(qgd_solver* s, const int* cells, int n)
{
    return guarded([&] {
        requireInit();
        if (!s || (n > 0 && !cells) || n < 0) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_const_sc_cells: bad argument");
        if (s->k.varSc != 7 && !s->v5) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_set_const_sc_cells: constScCellSet is read by varScModel7 and varScModel5 only");
        if (s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_set_const_sc_cells: call before qgd_qgdfoam_init_fields");
        const int nC = s->mesh->h.nCells;
        if (n == 0) { s->scConst.release(); return; }
        std::vector<unsigned char> mask(nC, 0);
        for (int i = 0; i < n; ++i) {
            if (cells[i] < 0 || cells[i] >= nC) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_const_sc_cells: cell id out of range");
            mask[cells[i]] = 1;
        }
        s->scConst.upload(mask, g_stream);
        QGD_CUDA(cudaStreamSynchronize(g_stream));
    });
}

int qgd_qgdfoam_set_bcs(qgd_solver* s, const int* bc_U, const int* bc_T, const int* bc_p, const double* val_U,
                        const double* val_T, const double* val_p)
{
    return guarded([&] {
        requireInit();
        if (!s || !bc_U || !bc_T || !bc_p) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_bcs: null argument");
        const HostMesh& h = s->mesh->h;
        std::vector<int> u(h.nBnd), t(h.nBnd), p(h.nBnd);
        s->anyQgdFlux = false;
        for (int b = 0; b < h.nBnd; ++b) {
            const int pi = h.bfacePatch[b];
            u[b] = bc_U[pi]; t[b] = bc_T[pi]; p[b] = bc_p[pi];
            if (h.patchKind[pi] == QGD_PATCH_EMPTY) continue;
            if (u[b] != QGD_BC_FIXED_VALUE && u[b] != QGD_BC_ZERO_GRADIENT && u[b] != QGD_BC_SLIP && u[b] != QGD_BC_WEDGE)
                throw Error(QGD_ERR_UNSUPPORTED, "U boundary condition outside the device-native set (fixedValue, zeroGradient, slip, wedge)");
            if (u[b] == QGD_BC_SLIP && s->k.implicit)
                throw Error(QGD_ERR_UNSUPPORTED, "slip / symmetryPlane velocity with implicitDiffusion true is not available (explicit branch only)");
            if (u[b] == QGD_BC_WEDGE) {
                if (h.patchKind[pi] != QGD_PATCH_WEDGE) throw Error(QGD_ERR_INVALID, "wedge velocity condition on a patch that is not a wedge patch");
                if (s->k.implicit) throw Error(QGD_ERR_UNSUPPORTED, "wedge patches with implicitDiffusion true are not available (explicit branch only)");
                if (h.nOwned != h.nCells) throw Error(QGD_ERR_UNSUPPORTED, "wedge patches on extended sub-meshes (multi-GPU) are not available");
                // `reduced` keeps nf*snGrad on wedge faces too (reducedFaceNormalStencil.C:69-108), with the wedge patch's own snGrad
                // (cellT . U_P - U_P) deltaCoeffs / 2, which the fused boundary kernels do not form; the operator-level calls take it from the caller
                if (s->fvsc->reduced && !s->fvsc->lsq) throw Error(QGD_ERR_UNSUPPORTED, "wedge patches with the fvsc scheme `reduced` are not available in the solver step (GaussVolPoint / leastSquares only)");
            }
            if (t[b] != QGD_BC_FIXED_VALUE && t[b] != QGD_BC_ZERO_GRADIENT)
                throw Error(QGD_ERR_UNSUPPORTED, "T boundary condition outside the device-native set (fixedValue, zeroGradient)");
            if (p[b] != QGD_BC_FIXED_VALUE && p[b] != QGD_BC_ZERO_GRADIENT && p[b] != QGD_BC_QGD_FLUX)
                throw Error(QGD_ERR_UNSUPPORTED, "p boundary condition outside the device-native set (fixedValue, zeroGradient, qgdFlux)");
            if ((u[b] == QGD_BC_FIXED_VALUE && !val_U) || (t[b] == QGD_BC_FIXED_VALUE && !val_T) || (p[b] == QGD_BC_FIXED_VALUE && !val_p))
                throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_bcs: fixedValue patch without values");
        }
        // Whether the qgdFlux mid-step sequence runs (k_bnd_pre + the halo exchange of p_b) must be the same decision on every rank
        // of a decomposed run: it follows from the PATCH table - a sub-mesh keeps every global patch, possibly with no local face
        // (a rank that held no qgdFlux face used to skip the exchange its neighbours were waiting in: N = 8 hang, round 2)
        for (int pi = 0; pi < h.nPatches; ++pi)
            if (h.patchKind[pi] != QGD_PATCH_EMPTY && bc_p[pi] == QGD_BC_QGD_FLUX) s->anyQgdFlux = true;
        s->bcU.upload(u, g_stream); s->bcT.upload(t, g_stream); s->bcP.upload(p, g_stream);
        std::vector<double> zero3(3 * (size_t)h.nBnd, 0.0), zero1(h.nBnd, 0.0);
        h2d(s->bvU, val_U ? val_U : zero3.data(), 3 * (size_t)h.nBnd);
        h2d(s->bvT, val_T ? val_T : zero1.data(), (size_t)h.nBnd);
        h2d(s->bvP, val_p ? val_p : zero1.data(), (size_t)h.nBnd);
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        s->bcsSet = true;
    });
}

int qgd_qgdfoam_init_fields(qgd_solver* s, const double* U, const double* T, const double* p, const double* alphaQGD)
{
    return guarded([&] {
        requireInit();
        if (!s || !U || !T || !p) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_init_fields: null argument");
        if (!s->bcsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_init_fields: call qgd_qgdfoam_set_bcs first");
        const HostMesh& h = s->mesh->h;
        const size_t n = h.nCells;
        if (s->stage.n < QGD_STATE_DOUBLES_PER_CELL * n) s->stage.alloc(QGD_STATE_DOUBLES_PER_CELL * n);
        QGD_CUDA(cudaMemcpyAsync(s->stage.p, U, 3 * n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QGD_CUDA(cudaMemcpyAsync(s->stage.p + 3 * n, T, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QGD_CUDA(cudaMemcpyAsync(s->stage.p + 4 * n, p, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        if (s->k.model == 1) {
            // constScPrModel1n forms tauQGDf = I(alphaQGD)*hQGDf/I(c) until "U" is registered (first step, constScPrModel1n.C:104-105)
            s->k.tauMode = 2;
            s->stepsDone = 0;
        }
        if (alphaQGD) QGD_CUDA(cudaMemcpyAsync(s->aQGD.p, alphaQGD, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        else { std::vector<double> a(n, 0.5); QGD_CUDA(cudaMemcpyAsync(s->aQGD.p, a.data(), n * sizeof(double), cudaMemcpyHostToDevice, g_stream)); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
        launchInit(g_stream, s->k, s->fvsc->view(), s->sview(), s->bview(), s->stage.p, s->stage.p + 3 * n, s->stage.p + 4 * n,
                   !h.wedgePts.empty());
        if (s->v5) varSc5Init(s);
        if (s->halo.active) s->launches += haloExchange(s);
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        s->fieldsSet = true;
    });
}

int qgd_qgdfoam_step(qgd_solver* s, int n_steps)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_step: call qgd_qgdfoam_init_fields first");
        runSteps(s, n_steps);
    });
}

int qgd_qgdfoam_step_host(qgd_solver* s, int n_steps, const qgd_state_host* in, const qgd_state_host* out)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step_host: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_step_host: call qgd_qgdfoam_init_fields first");
        if (s->v5) throw Error(QGD_ERR_UNSUPPORTED, "qgd_qgdfoam_step_host: varScModel5 keeps ScQGD between steps (relaxation), which the host state does not carry; use qgd_qgdfoam_step");
        const size_t n = s->mesh->h.nCells;
        if (s->stage.n < QGD_STATE_DOUBLES_PER_CELL * n) s->stage.alloc(QGD_STATE_DOUBLES_PER_CELL * n);
        double* st = s->stage.p;
        const int nb = (int)((n + 255) / 256);
        if (in) {
            if (!in->rho || !in->U || !in->e || !in->p || !in->T || !in->rhoU || !in->rhoE || !in->mu)
                throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step_host: incomplete input state");
            const double* src[8] = {in->rho, in->U, in->e, in->p, in->T, in->rhoU, in->rhoE, in->mu};
            const size_t off[8] = {0, 1, 4, 5, 6, 7, 10, 11}, len[8] = {1, 3, 1, 1, 1, 3, 1, 1};
            for (int i = 0; i < 8; ++i)
                QGD_CUDA(cudaMemcpyAsync(st + off[i] * n, src[i], len[i] * n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
            if (s->k.model == 1) s->k.tauMode = s->stepsDone == 0 ? 2 : 1;
            k_pack_state<<<nb, 256, 0, g_stream>>>(s->k, (int)n, s->S.p, s->aQGD.p, s->mesh->hQGD.p, st);
            s->launches++;
        }
        runSteps(s, n_steps);
        if (out) {
            k_unpack_state<<<nb, 256, 0, g_stream>>>((int)n, s->S.p, st);
            s->launches++;
            double* dst[8] = {out->rho, out->U, out->e, out->p, out->T, out->rhoU, out->rhoE, out->mu};
            const size_t off[8] = {0, 1, 4, 5, 6, 7, 10, 11}, len[8] = {1, 3, 1, 1, 1, 3, 1, 1};
            for (int i = 0; i < 8; ++i)
                if (dst[i]) QGD_CUDA(cudaMemcpyAsync(dst[i], st + off[i] * n, len[i] * n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        }
        QGD_CUDA(cudaStreamSynchronize(g_stream));
    });
}

int qgd_qgdfoam_step_fields_host(qgd_solver* s, int n_steps, const qgd_fields_host* in, const qgd_fields_host* out)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step_fields_host: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_step_fields_host: call qgd_qgdfoam_init_fields first");
        if (s->v5) throw Error(QGD_ERR_UNSUPPORTED, "qgd_qgdfoam_step_fields_host: varScModel5 keeps ScQGD between steps (relaxation), which U, T, p do not carry; use qgd_qgdfoam_step");
        const size_t n = s->mesh->h.nCells;
        if (s->stage.n < QGD_STATE_DOUBLES_PER_CELL * n) s->stage.alloc(QGD_STATE_DOUBLES_PER_CELL * n);
        double* st = s->stage.p;
        if (in) {
            if (!in->U || !in->T || !in->p) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step_fields_host: incomplete input fields (U, T, p)");
            QGD_CUDA(cudaMemcpyAsync(st, in->U, 3 * n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
            QGD_CUDA(cudaMemcpyAsync(st + 3 * n, in->T, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
            QGD_CUDA(cudaMemcpyAsync(st + 4 * n, in->p, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
            if (s->k.model == 1) { s->k.tauMode = 2; s->stepsDone = 0; }       // "U" is registered after the first correct() again
            launchInit(g_stream, s->k, s->fvsc->view(), s->sview(), s->bview(), st, st + 3 * n, st + 4 * n, !s->mesh->h.wedgePts.empty());
            s->launches += 2 + (s->k.varSc ? 1 : 0);
            if (s->halo.active) s->launches += haloExchange(s);
        }
        runSteps(s, n_steps);
        if (out) {
            if (!out->U || !out->T || !out->p) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_step_fields_host: incomplete output fields (U, T, p)");
            const bool cons = out->rho || out->rhoU || out->rhoE;
            k_unpack_fields<<<(int)((n + 255) / 256), 256, 0, g_stream>>>((int)n, s->S.p, st, cons ? 1 : 0);
            s->launches++;
            QGD_CUDA(cudaMemcpyAsync(out->U, st, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
            QGD_CUDA(cudaMemcpyAsync(out->T, st + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
            QGD_CUDA(cudaMemcpyAsync(out->p, st + 4 * n, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
            if (out->rho) QGD_CUDA(cudaMemcpyAsync(out->rho, st + 5 * n, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
            if (out->rhoU) QGD_CUDA(cudaMemcpyAsync(out->rhoU, st + 6 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
            if (out->rhoE) QGD_CUDA(cudaMemcpyAsync(out->rhoE, st + 9 * n, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        }
        QGD_CUDA(cudaStreamSynchronize(g_stream));
    });
}

const char* qgd_qgdfoam_face_kernel(qgd_solver* s, int* l2hint)
{
    if (!s) return "";
    const bool pipe = s->pipe.mode == 1 && !s->desc.adjust_time_step && !s->k.varSc;
    const char* name = pipe ? "k_face_cell_pipeline" : faceKernelName(s->fvsc->view());
    if (l2hint) *l2hint = std::string(name) == "k_face_flux_tma" ? faceKernelL2Hint() : 0;
    return name;
}

int qgd_qgdfoam_get(qgd_solver* s, int field, double* cells, double* bnd)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get: null solver");
        if (!s->fieldsSet) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_get: no fields yet");
        const HostMesh& h = s->mesh->h;
        auto fetch = [&](const RecA* dA, const RecB* dB, size_t n, double* outp) {
            if (!outp || !n) return;
            std::vector<RecA> a(n);
            std::vector<RecB> b(n);
            QGD_CUDA(cudaMemcpyAsync(a.data(), dA, n * sizeof(RecA), cudaMemcpyDeviceToHost, g_stream));
            QGD_CUDA(cudaMemcpyAsync(b.data(), dB, n * sizeof(RecB), cudaMemcpyDeviceToHost, g_stream));
            QGD_CUDA(cudaStreamSynchronize(g_stream));
            const double gam = s->k.gamma;
            for (size_t i = 0; i < n; ++i) {
                switch (field) {
                    case 0: outp[i] = a[i].rho; break;
                    case 1: outp[3 * i] = b[i].rhoUx; outp[3 * i + 1] = b[i].rhoUy; outp[3 * i + 2] = b[i].rhoUz; break;
                    case 2: outp[i] = b[i].rhoE; break;
                    case 3: outp[3 * i] = a[i].Ux; outp[3 * i + 1] = a[i].Uy; outp[3 * i + 2] = a[i].Uz; break;
                    case 4: outp[i] = a[i].e; break;
                    case 5: outp[i] = a[i].p; break;
                    case 6: outp[i] = a[i].T; break;
                    case 7: outp[i] = b[i].c; break;
                    case 8: outp[i] = b[i].mu; break;
                    case 9: outp[i] = s->k.alphaEffGamma ? b[i].alphaEff / gam : b[i].alphaEff; break;
                    case 10: outp[i] = b[i].aByC; break;   // scaled below for cells
                    case 11: outp[i] = a[i].H; break;
                    case 12: outp[i] = s->k.ScB; break;       // varScModel5: per-face values, copied below
                    default: throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get: unknown field id");
                }
            }
        };
        if (cells) {
            // SoA field index of each public field id (vectors: first component)
            static const int fieldOf[12] = {0, 8, 11, 1, 4, 5, 6, 12, 13, 14, 15, 7};
            if (field < 0 || field > 12) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get: unknown field id");
            const size_t n = h.nCells;
            const int k0 = field < 12 ? fieldOf[field] : 0;
            if (field == 12) {
                if (s->scVar.n) { d2h(cells, s->scVar.p, n); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
                else for (size_t c = 0; c < n; ++c) cells[c] = s->k.ScQGD;
            } else if (field == 10 && s->v5) {          // varScModel5.C:207: tauQGD = alphaQGD hQGD / c (the slot holds alphaQGD)
                std::vector<double> cs(n);
                d2h(cells, s->S.p + 15 * n, n); d2h(cs.data(), s->S.p + 12 * n, n);
                QGD_CUDA(cudaStreamSynchronize(g_stream));
                for (size_t c = 0; c < n; ++c) cells[c] = cells[c] * h.hQGD[c] / cs[c];
            } else if (field == 1 || field == 3) {
                std::vector<double> t(3 * n);
                d2h(t.data(), s->S.p + (size_t)k0 * n, 3 * n);
                QGD_CUDA(cudaStreamSynchronize(g_stream));
                for (size_t c = 0; c < n; ++c) for (int d = 0; d < 3; ++d) cells[3 * c + d] = t[d * n + c];
            } else {
                d2h(cells, s->S.p + (size_t)k0 * n, n);
                QGD_CUDA(cudaStreamSynchronize(g_stream));
                if (field == 9 && s->k.alphaEffGamma) for (size_t c = 0; c < n; ++c) cells[c] /= s->k.gamma;
                if (field == 10) {
                    if (s->tauOut.n) { d2h(cells, s->tauOut.p, n); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
                    else for (size_t c = 0; c < n; ++c) cells[c] *= h.hQGD[c];
                }
            }
        }
        fetch(s->bA.p, s->bB.p, h.nBnd, bnd);
        if (s->v5 && bnd && h.nBnd && (field == 10 || field == 12)) {
            if (field == 12) { d2h(bnd, s->v5->ScB.p, (size_t)h.nBnd); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
            else {
                std::vector<RecB> b(h.nBnd);
                QGD_CUDA(cudaMemcpyAsync(b.data(), s->bB.p, (size_t)h.nBnd * sizeof(RecB), cudaMemcpyDeviceToHost, g_stream));
                QGD_CUDA(cudaStreamSynchronize(g_stream));
                for (int i = 0; i < h.nBnd; ++i) bnd[i] = b[i].aByC * h.hQGDf[h.nInternal + i] / b[i].c;
            }
        } else if (field == 10 && bnd) {
            if (s->tauOutB.n) { d2h(bnd, s->tauOutB.p, (size_t)h.nBnd); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
            else for (int b = 0; b < h.nBnd; ++b) bnd[b] *= h.hQGDf[h.nInternal + b];
        }
    });
}

int qgd_qgdfoam_get_flux(qgd_solver* s, int which, double* out)
{
    return guarded([&] {
        requireInit();
        if (!s || !out) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get_flux: null argument");
        if (s->pipe.mode != 0)
            throw Error(QGD_ERR_STATE, "qgd_qgdfoam_get_flux: face fluxes are not kept in HBM by the pipelined step; call "
                                       "qgd_qgdfoam_set_pipeline(s, 0, ...) before stepping to keep them");
        if (which < 0 || which > 2) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get_flux: unknown flux id");
        const HostMesh& h = s->mesh->h;
        const size_t nF = h.nFaces;
        const std::vector<int>& perm = s->mesh->facePerm;
        const int k = (which == 1) ? 3 : 1;
        const size_t off = which == 0 ? 0 : (which == 1 ? 1 : 4);
        std::vector<double> t(k * nF + 1);
        d2h(t.data(), s->Fflux.p + off * nF, k * nF);              // two-kernel form: one [5][nF] array
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        for (size_t f = 0; f < nF; ++f)
            for (int d = 0; d < k; ++d) out[(size_t)k * perm[f] + d] = t[d * nF + f];
    });
}

int qgd_qgdfoam_get_scalars(qgd_solver* s, double* delta_t, double* courant, double* time)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get_scalars: null solver");
        StepScalars sc;
        QGD_CUDA(cudaMemcpyAsync(&sc, s->sc.p, sizeof(sc), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        if (delta_t) *delta_t = sc.dt;
        if (courant) *courant = sc.coNum;
        if (time) *time = sc.time;
    });
}

int qgd_qgdfoam_state_guard(qgd_solver* s, int* first_step)
{
    return guarded([&] {
        requireInit();
        if (!s || !first_step) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_state_guard: null argument");
        StepScalars sc;
        QGD_CUDA(cudaMemcpyAsync(&sc, s->sc.p, sizeof(sc), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        *first_step = sc.guardStep;
    });
}

long long qgd_qgdfoam_launch_count(qgd_solver* s) { return s ? s->launches : 0; }

int qgd_qgdfoam_diffusion_iterations(qgd_solver* s, int iters[4])
{
    return guarded([&] {
        requireInit();
        if (!s || !iters) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_diffusion_iterations: null argument");
        if (!s->k.implicit || !s->impl.built) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_diffusion_iterations: no implicit step yet");
        PcgResult r[4];
        static_assert(sizeof(PcgResult) == 4 * sizeof(double), "PcgResult staging layout");
        QGD_CUDA(cudaMemcpyAsync(r, s->stage.p, sizeof(r), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        for (int j = 0; j < 4; ++j) iters[j] = r[j].iters;
    });
}

int qgd_qgdfoam_set_pipeline(qgd_solver* s, int mode, int chunk_cells, int lag, int ring_slots)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_pipeline: null solver");
        if (mode != 0 && mode != 1) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_pipeline: mode must be 0 or 1");
        if (mode == 1 && s->v5) throw Error(QGD_ERR_UNSUPPORTED, "qgd_qgdfoam_set_pipeline: varScModel5 runs in the two-kernel step form");
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        configurePipeline(s, mode, chunk_cells, lag, ring_slots);
    });
}

long long qgd_qgdfoam_graph_steps(qgd_solver* s) { return s ? s->graphSteps : 0; }

int qgd_qgdfoam_get_pipeline(qgd_solver* s, int* mode, int* chunk_cells, int* lag, int* ring_slots, int* n_chunks, int* grid)
{
    return guarded([&] {
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_get_pipeline: null solver");
        if (mode) *mode = s->pipe.mode;
        if (chunk_cells) *chunk_cells = s->pipe.chunkCells;
        if (lag) *lag = s->pipe.lag;
        if (ring_slots) *ring_slots = s->pipe.mode ? s->pipe.ringSlots : 0;
        if (n_chunks) *n_chunks = s->pipe.mode ? s->pipe.nChunks : 0;
        if (grid) *grid = s->pipe.grid;
    });
}

int qgd_qgdfoam_profile(qgd_solver* s, int enable)
{
    return guarded([&] {
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_profile: null solver");
        s->profiling = enable != 0;
        s->eventsUsed = 0;
    });
}

int qgd_qgdfoam_kernel_times(qgd_solver* s, double* ms_points, double* ms_face, double* ms_cell, int* n_steps)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_kernel_times: null solver");
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        double t[3] = {0, 0, 0};
        const size_t steps = s->eventsUsed / 6;
        const bool pts = !s->k.reducedScheme;
        for (size_t i = 0; i < steps; ++i)
            for (int k = 0; k < 3; ++k) {
                if (k == 0 && !pts) continue;
                float ms = 0.f;
                QGD_CUDA(cudaEventElapsedTime(&ms, s->events[6 * i + 2 * k], s->events[6 * i + 2 * k + 1]));
                t[k] += ms;
            }
        if (ms_points) *ms_points = t[0];
        if (ms_face) *ms_face = t[1];
        if (ms_cell) *ms_cell = t[2];
        if (n_steps) *n_steps = (int)steps;
    });
}

int qgd_comm_unique_id(void* out128)
{
    return guarded([&] {
        if (!out128) throw Error(QGD_ERR_INVALID, "qgd_comm_unique_id: null buffer");
        loadNccl();
        ncclUniqueId id;
        QGD_NCCL(g_nccl.GetUniqueId(&id));
        static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
        std::memcpy(out128, &id, sizeof(id));
    });
}

int qgd_comm_init(int rank, int n_ranks, const void* id128)
{
    return guarded([&] {
        requireInit();
        if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw Error(QGD_ERR_INVALID, "qgd_comm_init: bad arguments");
        loadNccl();
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        QGD_NCCL(g_nccl.CommInitRank(&g_comm, n_ranks, id, rank));
        g_rank = rank; g_nranks = n_ranks;
        if (!g_commStream) {
            // highest priority: the exchange's CTAs are placed as soon as an SM has room, ahead of the point gather's
            int prLo = 0, prHi = 0;
            QGD_CUDA(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
            QGD_CUDA(cudaStreamCreateWithPriority(&g_commStream, cudaStreamNonBlocking, prHi));
            QGD_CUDA(cudaEventCreateWithFlags(&g_evStep, cudaEventDisableTiming));
            QGD_CUDA(cudaEventCreateWithFlags(&g_evHalo, cudaEventDisableTiming));
        }
    });
}

int qgd_comm_finalize(void)
{
    return guarded([&] {
        if (g_comm) { cudaStreamSynchronize(g_stream); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
        g_nranks = 1; g_rank = 0;
    });
}

int qgd_qgdfoam_set_halo(qgd_solver* s, int nn, const int* nbr_rank, const int* send_cell_off, const int* send_cells,
                         const int* recv_cell_off, const int* recv_cells, const int* send_bf_off, const int* send_bfaces,
                         const int* recv_bf_off, const int* recv_bfaces)
{
    return guarded([&] {
        requireInit();
        if (!s || nn < 0) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_halo: bad arguments");
        if (nn > 0 && !g_comm) throw Error(QGD_ERR_STATE, "qgd_qgdfoam_set_halo: call qgd_comm_init first");
        if (s->v5) throw Error(QGD_ERR_UNSUPPORTED, "qgd_qgdfoam_set_halo: varScModel5 is not available in decomposed runs");
        qgd_solver::Halo& h = s->halo;
        h.nbr.assign(nbr_rank, nbr_rank + nn);
        h.sendCellOff.assign(send_cell_off, send_cell_off + nn + 1); h.recvCellOff.assign(recv_cell_off, recv_cell_off + nn + 1);
        h.sendBfOff.assign(send_bf_off, send_bf_off + nn + 1); h.recvBfOff.assign(recv_bf_off, recv_bf_off + nn + 1);
        const HostMesh& hm = s->mesh->h;
        auto check = [&](const int* ids, int n, int lo, int hi, const char* what) {
            for (int i = 0; i < n; ++i) if (ids[i] < lo || ids[i] >= hi) throw Error(QGD_ERR_INVALID, std::string("qgd_qgdfoam_set_halo: ") + what + " out of range");
        };
        check(send_cells, h.sendCellOff[nn], 0, hm.nOwned, "send cell");
        check(recv_cells, h.recvCellOff[nn], hm.nOwned, hm.nCells, "recv cell");
        check(send_bfaces, h.sendBfOff[nn], 0, hm.nBnd, "send boundary face");
        check(recv_bfaces, h.recvBfOff[nn], 0, hm.nBnd, "recv boundary face");
        auto up = [&](DevBuf<int>& d, const int* p, int n) { std::vector<int> v(p, p + n); if (v.empty()) v.push_back(0); d.upload(v, g_stream); };
        up(h.sendCells, send_cells, h.sendCellOff[nn]); up(h.recvCells, recv_cells, h.recvCellOff[nn]);
        up(h.sendBf, send_bfaces, h.sendBfOff[nn]); up(h.recvBf, recv_bfaces, h.recvBfOff[nn]);
        h.sendBuf.alloc((size_t)kCellDoubles * h.sendCellOff[nn] + (size_t)kBfDoubles * h.sendBfOff[nn] + 1);
        h.recvBuf.alloc((size_t)kCellDoubles * h.recvCellOff[nn] + (size_t)kBfDoubles * h.recvBfOff[nn] + 1);
        h.midSend.alloc(h.sendBfOff[nn] + 1); h.midRecv.alloc(h.recvBfOff[nn] + 1);
        auto plan = [&](const std::vector<int>& cOff, const std::vector<int>& bOff, DevBuf<int>& item, DevBuf<int>& cell, DevBuf<int>& bf,
                        DevBuf<long long>& bufOff) {
            std::vector<int> it(nn + 1, 0);
            std::vector<long long> bo(nn + 1, 0);
            for (int k = 0; k <= nn; ++k) { it[k] = cOff[k] + bOff[k]; bo[k] = (long long)kCellDoubles * cOff[k] + (long long)kBfDoubles * bOff[k]; }
            item.upload(it, g_stream); cell.upload(cOff, g_stream); bf.upload(bOff, g_stream); bufOff.upload(bo, g_stream);
        };
        plan(h.sendCellOff, h.sendBfOff, h.planSendItem, h.planSendCell, h.planSendBf, h.planSendBuf);
        plan(h.recvCellOff, h.recvBfOff, h.planRecvItem, h.planRecvCell, h.planRecvBf, h.planRecvBuf);
        h.active = nn > 0;
        {   // points whose cells are all owned vs points that read a halo copy (patch points are handled by k_patch_points)
            std::vector<int> pin, pha;
            for (int p = 0; p < hm.nPoints; ++p) {
                if (hm.pcOff[p + 1] == hm.pcOff[p]) continue;
                bool halo = false;
                for (int q = hm.pcOff[p]; q < hm.pcOff[p + 1]; ++q) halo = halo || hm.pcCell[q] >= hm.nOwned;
                (halo ? pha : pin).push_back(p);
            }
            if (!pin.empty() && !pha.empty()) { h.ptsInterior.upload(pin, g_stream); h.ptsHalo.upload(pha, g_stream); }
        }
        // halo copies initialised locally carry wrong mesh-derived values (hQGD of an open halo cell): take the owners'
        if (h.active && s->fieldsSet) { s->launches += haloExchange(s); QGD_CUDA(cudaStreamSynchronize(g_stream)); }
    });
}

int qgd_qgdfoam_set_halo_faces(qgd_solver* s, int nn, const int* nbr_rank, const int* send_off, const int* send_cells,
                               const int* recv_off, const int* recv_cells)
{
    return guarded([&] {
        requireInit();
        if (!s) throw Error(QGD_ERR_INVALID, "qgd_qgdfoam_set_halo_faces: null solver");
        const HostMesh& h = s->mesh->h;
        s->haloFace.set(nn, nbr_rank, send_off, send_cells, recv_off, recv_cells, h.nCells, h.nOwned, 9, g_stream);
    });
}

int qgd_pcg_solve(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x, double tolerance,
                  double rel_tol, int max_iter, int precond, int* iters, double* initial_residual, double* final_residual)
{
    return guarded([&] {
        requireInit();
        if (!mesh || !diag || !b || !x || (mesh->h.nInternal > 0 && !upper)) throw Error(QGD_ERR_INVALID, "qgd_pcg_solve: null argument");
        if (precond < 0 || precond > 2) throw Error(QGD_ERR_INVALID, "qgd_pcg_solve: preconditioner must be 0 (none), 1 (diagonal) or 2 (DIC)");
        PcgMatrix A;
        A.build(mesh->h, diag, upper, precond, g_stream);
        const size_t n = mesh->h.nCells;
        QGD_CUDA(cudaMemcpyAsync(A.b.p, b, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QGD_CUDA(cudaMemcpyAsync(A.x.p, x, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        A.solve(tolerance, rel_tol, max_iter, g_stream);
        PcgResult r;
        QGD_CUDA(cudaMemcpyAsync(x, A.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaMemcpyAsync(&r, A.out.p, sizeof(r), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        if (iters) *iters = r.iters;
        if (initial_residual) *initial_residual = r.res0;
        if (final_residual) *final_residual = r.res;
    });
}

// Decomposed PCG: every rank passes the matrix of its extended sub-mesh (rows of the owned cells are used, the columns of halo
// cells are read from halo entries), the face-neighbour exchange lists (decompose.SubDomain.send_face_cells / recv_face_cells) and
// its part of b and x.  Per iteration: one exchange of the search direction over NCCL send/recv and three all-reduced scalars.
// Oracle: or_pcg_solve_blocks (preconditioner none | diagonal are decomposition-independent); CPU prototype: tests/gloo_pcg_worker.py.
int qgd_pcg_solve_multi(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x, double tolerance,
                        double rel_tol, int max_iter, int precond, int n_neighbours, const int* nbr_rank, const int* send_off,
                        const int* send_cells, const int* recv_off, const int* recv_cells, int* iters, double* initial_residual,
                        double* final_residual)
{
    return guarded([&] {
        requireInit();
        if (!g_comm) throw Error(QGD_ERR_STATE, "qgd_pcg_solve_multi: call qgd_comm_init first");
        if (!mesh || !diag || !b || !x || (mesh->h.nInternal > 0 && !upper) || n_neighbours < 0 ||
            (n_neighbours > 0 && (!nbr_rank || !send_off || !send_cells || !recv_off || !recv_cells)))
            throw Error(QGD_ERR_INVALID, "qgd_pcg_solve_multi: null argument");
        if (precond < 0 || precond > 2 || (precond == 2 && mesh->h.pcgBlock.empty()))
            throw Error(QGD_ERR_UNSUPPORTED, "qgd_pcg_solve_multi: preconditioner must be 0 (none), 1 (diagonal) or 2 (DIC, with DIC blocks set on the mesh)");
        const HostMesh& h = mesh->h;
        const size_t n = h.nCells;
        const int nOwned = h.nOwned, nn = n_neighbours;
        std::vector<double> d(diag, diag + n);
        for (size_t c = nOwned; c < n; ++c) d[c] = 1.0;            // halo rows are never solved; keep 1/diag finite
        PcgMatrix A;
        A.build(h, d.data(), upper, precond, g_stream);
        QGD_CUDA(cudaMemcpyAsync(A.b.p, b, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QGD_CUDA(cudaMemcpyAsync(A.x.p, x, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        const int nS = nn ? send_off[nn] : 0, nR = nn ? recv_off[nn] : 0;
        DevBuf<int> sIds, rIds;
        DevBuf<double> sBuf, rBuf;
        sIds.upload(std::vector<int>(send_cells, send_cells + std::max(nS, 0)), g_stream);
        rIds.upload(std::vector<int>(recv_cells, recv_cells + std::max(nR, 0)), g_stream);
        sBuf.alloc(std::max(nS, 1)); rBuf.alloc(std::max(nR, 1));
        PcgHooks hooks;
        hooks.exchange = [&](double* vec, cudaStream_t st) {
            if (nS) k_gather1<<<(nS + 255) / 256, 256, 0, st>>>(nS, sIds.p, vec, sBuf.p);
            QGD_NCCL(g_nccl.GroupStart());
            for (int k = 0; k < nn; ++k) {
                const int ns = send_off[k + 1] - send_off[k], nr = recv_off[k + 1] - recv_off[k];
                if (ns) QGD_NCCL(g_nccl.Send(sBuf.p + send_off[k], ns, ncclDouble, nbr_rank[k], g_comm, st));
                if (nr) QGD_NCCL(g_nccl.Recv(rBuf.p + recv_off[k], nr, ncclDouble, nbr_rank[k], g_comm, st));
            }
            QGD_NCCL(g_nccl.GroupEnd());
            if (nR) k_scatter1<<<(nR + 255) / 256, 256, 0, st>>>(nR, rIds.p, rBuf.p, vec);
        };
        hooks.allreduceSum = [&](double* dev, int count, cudaStream_t st) {
            QGD_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, g_comm, st));
        };
        StepwisePcg sw;
        sw.alloc(A, nOwned);
        PcgResult r;
        sw.solve(A, A.b.p, A.x.p, tolerance, rel_tol, max_iter, precond, g_stream, &hooks, &r);
        hooks.exchange(A.x.p, g_stream);                             // the solution's halo entries for the caller
        QGD_CUDA(cudaMemcpyAsync(x, A.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        if (iters) *iters = r.iters;
        if (initial_residual) *initial_residual = r.res0;
        if (final_residual) *final_residual = r.res;
    });
}

// the stepwise (multi-kernel) form of the same solver on one GPU: validates the kernels a decomposed run will use
int qgd_pcg_solve_stepwise(qgd_mesh* mesh, const double* diag, const double* upper, const double* b, double* x, double tolerance,
                           double rel_tol, int max_iter, int precond, int* iters, double* initial_residual, double* final_residual)
{
    return guarded([&] {
        requireInit();
        if (!mesh || !diag || !b || !x || (mesh->h.nInternal > 0 && !upper)) throw Error(QGD_ERR_INVALID, "qgd_pcg_solve_stepwise: null argument");
        if (precond < 0 || precond > 2 || (precond == 2 && mesh->h.pcgBlock.empty()))
            throw Error(QGD_ERR_UNSUPPORTED, "qgd_pcg_solve_stepwise: preconditioner must be 0 (none), 1 (diagonal) or 2 (DIC, with DIC blocks set on the mesh)");
        PcgMatrix A;
        A.build(mesh->h, diag, upper, precond, g_stream);
        const size_t n = mesh->h.nCells;
        QGD_CUDA(cudaMemcpyAsync(A.b.p, b, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QGD_CUDA(cudaMemcpyAsync(A.x.p, x, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        StepwisePcg sw;
        sw.alloc(A, (int)n);
        PcgResult r;
        sw.solve(A, A.b.p, A.x.p, tolerance, rel_tol, max_iter, precond, g_stream, nullptr, &r);
        QGD_CUDA(cudaMemcpyAsync(x, A.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        QGD_CUDA(cudaStreamSynchronize(g_stream));
        if (iters) *iters = r.iters;
        if (initial_residual) *initial_residual = r.res0;
        if (final_residual) *final_residual = r.res;
    });
}

} // extern "C"
