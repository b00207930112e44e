// Stepwise PCG: the same algorithm as k_pcg (OpenFOAM v2312 PCG.C + DIC (blocks) / diagonal / no preconditioner, normFactor residual
// norm; SURVEY.md App. A.5, call site QHDpEqn.H:45) cut into one kernel per phase, so that a halo exchange of the search
// direction and all-reduces of the dot products can be placed between the phases: the form a decomposed (multi-GPU) run needs,
// where every rank owns the rows [0, nRows) of its extended sub-mesh matrix and reads neighbour values from halo entries
// (oracle: or_pcg_solve_blocks; CPU prototype with real message passing: tests/gloo_pcg_worker.py).
//
// Nothing is synchronised with the host inside an iteration: all scalars (wArA, wApA, residual, normFactor, the
// convergence flag) live on the device, kernels of iterations issued after convergence return at once, and the host looks
// at the flag once per chunk of iterations.  Without hooks the solver runs on one GPU (qgd_pcg_solve_stepwise, used to
// validate the kernels against the oracle before any communication is involved).
//
// Used by QHDFoam (pressure) and the implicit branch of QGDFoam (U, e) on extended sub-meshes; preconditioner none | diagonal |
// block-local DIC (launchDicBlocks, qgd_pcg.cu).  Parity with the oracle: 1 GPU (tests/test_gpu_extra.py), 2 and 8 GPUs
// (tests/multi_gpu_pcg_worker.py, multi_gpu_qhd_worker.py, multi_gpu_worker.py implicit cases).
#include <algorithm>

#include "qgd_pcg.cuh"

namespace qgd {

namespace {

constexpr int kB = 256;

// red[] slots
enum { R_WARA = 0, R_WAPA = 1, R_SUMR = 2, R_WARA_OLD = 3, R_NORM = 4, R_RES0 = 5, R_RES = 6, R_XSUM = 7, R_NROWS = 8, R_NF = 9, R_COUNT = 10 };
// state[] slots
enum { S_DONE = 0, S_ITERS = 1, S_COUNT = 2 };

struct SwView {
    int n, nRows, W;
    const int* enc; const double* coef;
    const int* tailOff; const int* tailEnc; const double* tailCoef;
    const double* diag; const double* rD;
    const double* b; double* x;
    double* r; double* w; double* p;
    double* partials;            // [2][gridDim.x]
    double* red; int* state;
    double tol, relTol; int maxIter, precond;
};

// block sum of two values -> partials[0][block], partials[1][block] (fixed order inside the block)
__device__ __forceinline__ void blockPartials(double a, double b, double* partials)
{
    __shared__ double sa[kB / 32], sb[kB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sa[wid] = a; sb[wid] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
#pragma unroll
        for (int k = 0; k < kB / 32; ++k) { x += sa[k]; y += sb[k]; }
        partials[blockIdx.x] = x;
        partials[gridDim.x + blockIdx.x] = y;
    }
}

__device__ __forceinline__ double rowDot(const SwView& v, int c, const double* __restrict__ vec)
{
    double y = 0.0;
    for (int j = 0; j < v.W; ++j) y += __ldg(&v.coef[(size_t)j * v.n + c]) * vec[__ldg(&v.enc[(size_t)j * v.n + c]) >> 1];
    for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) y += __ldg(&v.tailCoef[q]) * vec[__ldg(&v.tailEnc[q]) >> 1];
    return y;
}

// wA = A x ; rA = b - wA ; partial sum of x   (x must carry valid halo entries)
__global__ void __launch_bounds__(kB) k_sw_init1(SwView v)
{
    double sx = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) {
        const double xc = v.x[c];
        const double y = __ldg(&v.diag[c]) * xc + rowDot(v, c, v.x);
        v.w[c] = y;
        v.r[c] = __ldg(&v.b[c]) - y;
        v.p[c] = 0.0;
        sx += xc;
    }
    blockPartials(sx, 0.0, v.partials);
}

// normFactor and initial residual partials (xRef = red[R_XSUM] / red[R_NROWS], both already global)
__global__ void __launch_bounds__(kB) k_sw_init2(SwView v)
{
    const double xRef = v.red[R_XSUM] / v.red[R_NROWS];
    double nf = 0.0, sr = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) {
        double sumA = __ldg(&v.diag[c]);
        for (int j = 0; j < v.W; ++j) sumA += __ldg(&v.coef[(size_t)j * v.n + c]);
        for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) sumA += __ldg(&v.tailCoef[q]);
        const double t = sumA * xRef;
        nf += fabs(v.w[c] - t) + fabs(__ldg(&v.b[c]) - t);
        sr += fabs(v.r[c]);
    }
    blockPartials(nf, sr, v.partials);
}

// wA = M^-1 rA ; partial wA.rA
__global__ void __launch_bounds__(kB) k_sw_precond(SwView v)
{
    if (v.state[S_DONE]) return;
    double zr = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) {
        const double rc = v.r[c];
        const double zc = v.precond ? __ldg(&v.rD[c]) * rc : rc;
        v.w[c] = zc;
        zr += zc * rc;
    }
    blockPartials(zr, 0.0, v.partials);
}

// pA = wA + beta pA on the owned rows (the halo entries follow by exchange)
__global__ void __launch_bounds__(kB) k_sw_pupdate(SwView v)
{
    if (v.state[S_DONE]) return;
    const double beta = v.state[S_ITERS] == 0 ? 0.0 : v.red[R_WARA] / v.red[R_WARA_OLD];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) v.p[c] = v.w[c] + beta * v.p[c];
}

// wA = A pA ; partial wA.pA
__global__ void __launch_bounds__(kB) k_sw_spmv(SwView v)
{
    if (v.state[S_DONE]) return;
    double wp = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) {
        const double pc = v.p[c];
        const double y = __ldg(&v.diag[c]) * pc + rowDot(v, c, v.p);
        v.w[c] = y;
        wp += y * pc;
    }
    blockPartials(wp, 0.0, v.partials);
}

// x += alpha pA ; rA -= alpha wA ; partial sum |rA|
__global__ void __launch_bounds__(kB) k_sw_update(SwView v)
{
    if (v.state[S_DONE]) return;
    const double alpha = v.red[R_WARA] / v.red[R_WAPA];
    double sr = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < v.nRows; c += gridDim.x * blockDim.x) {
        v.x[c] += alpha * v.p[c];
        const double rc = v.r[c] - alpha * v.w[c];
        v.r[c] = rc;
        sr += fabs(rc);
    }
    blockPartials(sr, 0.0, v.partials);
}

// one block: partials -> red[] (fixed order), before the all-reduce.  what: 0 init1 (xsum, nRows), 1 init2 (nf, sumr),
// 2 precond (wArA; the previous value moves to R_WARA_OLD), 3 spmv (wApA), 4 update (sumr)
__global__ void __launch_bounds__(kB) k_sw_collect(SwView v, int G, int what)
{
    if (what >= 2 && v.state[S_DONE]) return;
    __shared__ double sa[kB], sb[kB];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < G; i += kB) { a += v.partials[i]; b += v.partials[G + i]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int k = 0; k < kB; ++k) { x += sa[k]; y += sb[k]; }
        if (what == 0) { v.red[R_XSUM] = x; v.red[R_NROWS] = (double)v.nRows; }
        else if (what == 1) { v.red[R_NF] = x; v.red[R_SUMR] = y; }
        else if (what == 2) { v.red[R_WARA_OLD] = v.red[R_WARA]; v.red[R_WARA] = x; }
        else if (what == 3) v.red[R_WAPA] = x;
        else v.red[R_SUMR] = x;
    }
}

// one thread: the scalar logic after the all-reduce.  what: 1 after init2, 3 after spmv (singularity check), 4 after update
__global__ void k_sw_decide(SwView v, int what)
{
    if (what == 1) {
        const double nf = v.red[R_NF] + 1e-20;
        const double res0 = v.red[R_SUMR] / nf;
        v.red[R_NORM] = nf; v.red[R_RES0] = res0; v.red[R_RES] = res0;
        v.state[S_ITERS] = 0;
        v.state[S_DONE] = (res0 < v.tol || (v.relTol > 1e-20 && res0 < v.relTol * res0) || v.maxIter <= 0) ? 1 : 0;
        return;
    }
    if (v.state[S_DONE]) return;
    if (what == 3) {
        if (fabs(v.red[R_WAPA]) / v.red[R_NORM] < 1e-300) v.state[S_DONE] = 1;      // solverPerformance::checkSingularity
        return;
    }
    const double res = v.red[R_SUMR] / v.red[R_NORM];
    v.red[R_RES] = res;
    const int it = ++v.state[S_ITERS];
    if (it >= v.maxIter || res < v.tol || (v.relTol > 1e-20 && res < v.relTol * v.red[R_RES0])) v.state[S_DONE] = 1;
}

} // namespace

void StepwisePcg::alloc(const PcgMatrix& A, int rows)
{
    n = A.n; nRows = rows;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    grid = std::max(1, std::min(sms * 8, (nRows + kB - 1) / kB));
    // the block-local DIC works on whole-mesh vectors (halo cells are singleton blocks): r and w span all n cells then
    r.alloc(A.nBlocks ? n : nRows); w.alloc(A.nBlocks ? n : nRows); p.alloc(n);
    if (A.nBlocks) { r.zero(); w.zero(); }
    partials.alloc(2 * (size_t)std::max(grid, dicBlocksGrid(A))); red.alloc(R_COUNT); state.alloc(S_COUNT);
}

// b: nRows device doubles ; x: n device doubles (owned rows first, halo entries valid on entry) ; precond 0 | 1
int StepwisePcg::solve(const PcgMatrix& A, const double* b, double* x, double tol, double relTol, int maxIter, int precond,
                       cudaStream_t st, const PcgHooks* hooks, PcgResult* result)
{
    if (precond < 0 || precond > 2) throw Error(QGD_ERR_UNSUPPORTED, "stepwise PCG: preconditioner must be none, diagonal or DIC");
    if (precond == 2 && A.nBlocks == 0)
        throw Error(QGD_ERR_UNSUPPORTED, "stepwise PCG: DIC needs DIC blocks on the mesh (qgd_mesh_make_pcg_blocks / qgd_mesh_set_pcg_blocks): in a "
                                         "decomposed run the preconditioner is local to a block");
    if (precond != 0 && A.precond != precond) throw Error(QGD_ERR_STATE, "stepwise PCG: the matrix was built with another preconditioner");
    SwView v;
    v.n = n; v.nRows = nRows; v.W = A.W; v.enc = A.enc.p; v.coef = A.coef.p; v.tailOff = A.tailOff.p; v.tailEnc = A.tailEnc.p; v.tailCoef = A.tailCoef.p;
    v.diag = A.diag.p; v.rD = A.rD.p; v.b = b; v.x = x; v.r = r.p; v.w = w.p; v.p = p.p; v.partials = partials.p; v.red = red.p; v.state = state.p;
    v.tol = tol; v.relTol = relTol; v.maxIter = maxIter; v.precond = precond;
    int launches = 0;
    auto reduce = [&](int what, int first, int count) {      // partials -> red[first .. first+count) -> global sum
        k_sw_collect<<<1, kB, 0, st>>>(v, grid, what); ++launches;
        if (hooks && hooks->allreduceSum) hooks->allreduceSum(red.p + first, count, st);
    };
    QGD_CUDA(cudaMemsetAsync(red.p, 0, R_COUNT * sizeof(double), st));
    QGD_CUDA(cudaMemsetAsync(state.p, 0, S_COUNT * sizeof(int), st));
    QGD_CUDA(cudaMemsetAsync(p.p, 0, (size_t)n * sizeof(double), st));
    if (hooks && hooks->exchange) hooks->exchange(x, st);
    k_sw_init1<<<grid, kB, 0, st>>>(v); ++launches;
    reduce(0, R_XSUM, 2);                                     // R_XSUM, R_NROWS are adjacent
    k_sw_init2<<<grid, kB, 0, st>>>(v); ++launches;
    k_sw_collect<<<1, kB, 0, st>>>(v, grid, 1); ++launches;
    if (hooks && hooks->allreduceSum) { hooks->allreduceSum(red.p + R_NF, 1, st); hooks->allreduceSum(red.p + R_SUMR, 1, st); }
    k_sw_decide<<<1, 1, 0, st>>>(v, 1); ++launches;
    const int chunk = 16;
    int hostState[S_COUNT] = {0, 0};
    for (int it0 = 0; it0 < maxIter; it0 += chunk) {
        for (int k = 0; k < chunk && it0 + k < maxIter; ++k) {
            if (precond == 2) {
                launchDicBlocks(A, r.p, w.p, partials.p, state.p + S_DONE, st); ++launches;
                k_sw_collect<<<1, kB, 0, st>>>(v, dicBlocksGrid(A), 2); ++launches;
                if (hooks && hooks->allreduceSum) hooks->allreduceSum(red.p + R_WARA, 1, st);
            } else {
                k_sw_precond<<<grid, kB, 0, st>>>(v); ++launches;
                reduce(2, R_WARA, 1);
            }
            k_sw_pupdate<<<grid, kB, 0, st>>>(v); ++launches;
            if (hooks && hooks->exchange) hooks->exchange(p.p, st);
            k_sw_spmv<<<grid, kB, 0, st>>>(v); ++launches;
            reduce(3, R_WAPA, 1);
            k_sw_decide<<<1, 1, 0, st>>>(v, 3); ++launches;
            k_sw_update<<<grid, kB, 0, st>>>(v); ++launches;
            reduce(4, R_SUMR, 1);
            k_sw_decide<<<1, 1, 0, st>>>(v, 4); ++launches;
        }
        QGD_CUDA(cudaMemcpyAsync(hostState, state.p, sizeof(hostState), cudaMemcpyDeviceToHost, st));
        QGD_CUDA(cudaStreamSynchronize(st));
        if (hostState[S_DONE]) break;
    }
    QGD_CUDA(cudaGetLastError());
    if (result) {
        double h[R_COUNT];
        QGD_CUDA(cudaMemcpyAsync(h, red.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        QGD_CUDA(cudaMemcpyAsync(hostState, state.p, sizeof(hostState), cudaMemcpyDeviceToHost, st));
        QGD_CUDA(cudaStreamSynchronize(st));
        result->iters = hostState[S_ITERS]; result->res0 = h[R_RES0]; result->res = h[R_RES]; result->normFactor = h[R_NORM];
    }
    return launches;
}

} // namespace qgd
