// Device-resident PCG for symmetric LDU matrices (see qgd_pcg.cuh).  Algorithm = OpenFOAM v2312 PCG.C / DICPreconditioner.C /
// diagonalPreconditioner.C / lduMatrix::solver::normFactor as restated in SURVEY.md App. A.5 (call site QHDpEqn.H:45):
//   wA = A x ; rA = b - wA ; normFactor = sum(|wA - xRef*sumA| + |b - xRef*sumA|) + 1e-20 ; residual = sum|rA|/normFactor
//   loop: wA = M^-1 rA ; wArA = wA.rA ; pA = wA + (wArA/wArAold) pA ; wA = A pA ; alpha = wArA/(wA.pA) ; x += alpha pA ;
//         rA -= alpha wA
// B200 design: one cooperative persistent kernel, 2 grid-wide reductions per iteration (Jacobi), SpMV as an atomic-free
// row gather over the cell->face ELL, p-update fused into the SpMV by recomputing the neighbours' p on the fly,
// reductions = warp shuffles -> per-block partial -> every block sums the partials in a fixed order (deterministic).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "qgd_pcg.cuh"

namespace cg = cooperative_groups;

namespace qgd {

namespace {

constexpr int kPcgBlock = 256;

// grid-wide sum of two values; identical result in every thread of the grid
__device__ double2 gridSum2(cg::grid_group& grid, double a, double b, double* partials, int& slot)
{
    __shared__ double sa[kPcgBlock / 32], sb[kPcgBlock / 32];
    __shared__ double2 tot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sa[wid] = a; sb[wid] = b; }
    __syncthreads();
    const int G = gridDim.x;
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
#pragma unroll
        for (int k = 0; k < kPcgBlock / 32; ++k) { x += sa[k]; y += sb[k]; }
        __stcg(&partials[(size_t)(slot * 2 + 0) * G + blockIdx.x], x);
        __stcg(&partials[(size_t)(slot * 2 + 1) * G + blockIdx.x], y);
    }
    grid.sync();
    if (threadIdx.x < 32) {
        double x = 0.0, y = 0.0;
        for (int i = lane; i < G; i += 32) {
            x += __ldcg(&partials[(size_t)(slot * 2 + 0) * G + i]);
            y += __ldcg(&partials[(size_t)(slot * 2 + 1) * G + i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { x += __shfl_down_sync(0xffffffffu, x, o); y += __shfl_down_sync(0xffffffffu, y, o); }
        if (lane == 0) tot = make_double2(x, y);
    }
    __syncthreads();
    const double2 r = tot;
    __syncthreads();
    slot ^= 1;
    return r;
}

// DIC: z = rD r ; forward sweep (ascending levels) ; backward sweep (descending levels) ; returns sum(z*r) contributions
template <int WT>
__device__ void dicSweeps(cg::grid_group& grid, const PcgView& v, int tid, int nth)
{
    const int n = v.n, W = WT ? WT : v.W;
    for (int c = tid; c < n; c += nth) v.z[c] = __ldg(&v.rD[c]) * v.r[c];
    grid.sync();
    for (int l = 1; l < v.nLevels; ++l) {          // wA[u] -= rD[u]*upper*wA[l], faces ascending
        for (int i = __ldg(&v.lvlOff[l]) + tid; i < __ldg(&v.lvlOff[l + 1]); i += nth) {
            const int c = __ldg(&v.lvlCells[i]);
            const double rd = __ldg(&v.rD[c]);
            double zc = v.z[c];
            for (int j = 0; j < W; ++j) {
                const int e = __ldg(&v.enc[(size_t)j * n + c]);
                if (e & 1) zc -= rd * __ldg(&v.coef[(size_t)j * n + c]) * v.z[e >> 1];
            }
            for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) {
                const int e = __ldg(&v.tailEnc[q]);
                if (e & 1) zc -= rd * __ldg(&v.tailCoef[q]) * v.z[e >> 1];
            }
            v.z[c] = zc;
        }
        grid.sync();
    }
    for (int l = v.nLevels - 2; l >= 0; --l) {     // wA[l] -= rD[l]*upper*wA[u], faces descending
        for (int i = __ldg(&v.lvlOff[l]) + tid; i < __ldg(&v.lvlOff[l + 1]); i += nth) {
            const int c = __ldg(&v.lvlCells[i]);
            const double rd = __ldg(&v.rD[c]);
            double zc = v.z[c];
            for (int q = __ldg(&v.tailOff[c + 1]) - 1; q >= __ldg(&v.tailOff[c]); --q) {
                const int e = __ldg(&v.tailEnc[q]);
                if (!(e & 1)) zc -= rd * __ldg(&v.tailCoef[q]) * v.z[e >> 1];
            }
            for (int j = W - 1; j >= 0; --j) {
                const int e = __ldg(&v.enc[(size_t)j * n + c]);
                if (!(e & 1)) zc -= rd * __ldg(&v.coef[(size_t)j * n + c]) * v.z[e >> 1];
            }
            v.z[c] = zc;
        }
        grid.sync();
    }
}

// ---- block-local DIC.  One work unit per block at a time: the block's rows (in-block entries, block order), its level offsets
// and its part of the vector are staged in shared memory, the forward / backward sweeps run level by level with a unit-local
// barrier only - no grid-wide synchronisation inside the preconditioner.  The unit is a WARP when a block fits a warp's share of
// the shared memory (PcgView::dicWarp: 8 blocks in flight per CTA, __syncwarp between levels - the levels of a small tile hold at
// most a few dozen cells, so a CTA-wide barrier would idle 7 of 8 warps), else the whole CTA.  Operation order per cell =
// ascending (forward) / descending (backward) face order of its in-block entries, i.e. the sequential sweeps of DICPreconditioner
// on the block's own lduMatrix.
// smem layout of one unit: z[B] | rD[B] | coef[Wb][B] | enc[Wb][B] | lvl[maxLevels+2]   (B = v.maxBlockCells)
struct BlockSmem { double* z; double* rD; double* coef; int* enc; int* lvl; };
__device__ __forceinline__ BlockSmem blockSmem(const PcgView& v, unsigned char* raw)
{
    BlockSmem s;
    const size_t B = v.maxBlockCells;
    s.z = reinterpret_cast<double*>(raw);
    s.rD = s.z + B;
    s.coef = s.rD + B;
    s.enc = reinterpret_cast<int*>(s.coef + (size_t)v.Wb * B);
    s.lvl = s.enc + (size_t)v.Wb * B;
    return s;
}
template <bool WARP> struct Unit {
    __device__ static int lanes() { return WARP ? 32 : (int)blockDim.x; }
    __device__ static int lane() { return WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
    __device__ static int id() { return WARP ? (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : (int)blockIdx.x; }
    __device__ static int count() { return WARP ? (int)(gridDim.x * (blockDim.x >> 5)) : (int)gridDim.x; }
    __device__ static void sync() { if (WARP) __syncwarp(); else __syncthreads(); }
};
// stage rows + level offsets of block [p0, p0+nb); src: the per-cell array that seeds rD (the reciprocal diagonal or, for the
// factorisation, the matrix diagonal)
template <bool WARP>
__device__ __forceinline__ int blockStage(const PcgView& v, const BlockSmem& sm, int b, int& p0, int& nb, const double* src)
{
    const size_t n = v.n, B = v.maxBlockCells;
    p0 = __ldg(&v.bOff[b]); nb = __ldg(&v.bOff[b + 1]) - p0;
    const int l0 = __ldg(&v.bLvlStart[b]), nl = __ldg(&v.bLvlStart[b + 1]) - l0 - 1;
    for (int i = Unit<WARP>::lane(); i <= nl; i += Unit<WARP>::lanes()) sm.lvl[i] = __ldg(&v.bLvlOff[l0 + i]) - p0;
    for (int i = Unit<WARP>::lane(); i < nb; i += Unit<WARP>::lanes()) {
        sm.rD[i] = src[__ldg(&v.bCells[p0 + i])];
        for (int j = 0; j < v.Wb; ++j) {
            sm.enc[j * B + i] = __ldg(&v.bEnc[(size_t)j * n + p0 + i]);
            sm.coef[j * B + i] = __ldg(&v.bCoef[(size_t)j * n + p0 + i]);
        }
    }
    return nl;
}
// z = M^-1 r on every block; returns this thread's part of sum(z*r)
template <bool WARP>
__device__ double dicBlocksT(const PcgView& v, unsigned char* raw)
{
    const BlockSmem sm = blockSmem(v, raw + (WARP ? (size_t)(threadIdx.x >> 5) * v.dicUnitSmem : 0));
    const size_t B = v.maxBlockCells;
    const int L = Unit<WARP>::lanes(), t = Unit<WARP>::lane();
    double zr = 0.0;
    for (int b = Unit<WARP>::id(); b < v.nBlocks; b += Unit<WARP>::count()) {
        int p0, nb;
        Unit<WARP>::sync();                                // the previous block's shared data is no longer read
        const int nl = blockStage<WARP>(v, sm, b, p0, nb, v.rD);
        for (int i = t; i < nb; i += L) sm.z[i] = sm.rD[i] * v.r[__ldg(&v.bCells[p0 + i])];
        Unit<WARP>::sync();
        for (int l = 1; l < nl; ++l) {                     // wA[u] -= rD[u]*upper*wA[l], faces ascending
            for (int i = sm.lvl[l] + t; i < sm.lvl[l + 1]; i += L) {
                const double rd = sm.rD[i];
                double zc = sm.z[i];
                for (int j = 0; j < v.Wb; ++j) {
                    const int en = sm.enc[j * B + i];
                    if (en >= 0 && (en & 1)) zc -= rd * sm.coef[j * B + i] * sm.z[en >> 1];
                }
                sm.z[i] = zc;
            }
            Unit<WARP>::sync();
        }
        for (int l = nl - 2; l >= 0; --l) {                // wA[l] -= rD[l]*upper*wA[u], faces descending
            for (int i = sm.lvl[l] + t; i < sm.lvl[l + 1]; i += L) {
                const double rd = sm.rD[i];
                double zc = sm.z[i];
                for (int j = v.Wb - 1; j >= 0; --j) {
                    const int en = sm.enc[j * B + i];
                    if (en >= 0 && !(en & 1)) zc -= rd * sm.coef[j * B + i] * sm.z[en >> 1];
                }
                sm.z[i] = zc;
            }
            Unit<WARP>::sync();
        }
        for (int i = t; i < nb; i += L) {
            const int c = __ldg(&v.bCells[p0 + i]);
            const double zc = sm.z[i];
            v.z[c] = zc;
            zr += zc * v.r[c];
        }
    }
    if (!WARP) __syncthreads();
    return zr;
}
__device__ __forceinline__ double dicBlocks(const PcgView& v, unsigned char* raw)
{
    return v.dicWarp ? dicBlocksT<true>(v, raw) : dicBlocksT<false>(v, raw);
}

// DIC::calcReciprocalD on every block: rD = diag ; faces ascending (in-block): rD[u] -= upper^2/rD[l] ; rD = 1/rD
template <bool WARP>
__device__ void dicFactorT(const PcgView& v, unsigned char* raw, double* rDout)
{
    const BlockSmem sm = blockSmem(v, raw + (WARP ? (size_t)(threadIdx.x >> 5) * v.dicUnitSmem : 0));
    const size_t B = v.maxBlockCells;
    const int L = Unit<WARP>::lanes(), t = Unit<WARP>::lane();
    for (int b = Unit<WARP>::id(); b < v.nBlocks; b += Unit<WARP>::count()) {
        int p0, nb;
        Unit<WARP>::sync();
        const int nl = blockStage<WARP>(v, sm, b, p0, nb, v.diag);
        Unit<WARP>::sync();
        for (int l = 1; l < nl; ++l) {
            for (int i = sm.lvl[l] + t; i < sm.lvl[l + 1]; i += L) {
                double rc = sm.rD[i];
                for (int j = 0; j < v.Wb; ++j) {
                    const int en = sm.enc[j * B + i];
                    if (en >= 0 && (en & 1)) { const double cf = sm.coef[j * B + i]; rc -= cf * cf / sm.rD[en >> 1]; }
                }
                sm.rD[i] = rc;
            }
            Unit<WARP>::sync();
        }
        for (int i = t; i < nb; i += L) rDout[__ldg(&v.bCells[p0 + i])] = 1.0 / sm.rD[i];
    }
}
__global__ void __launch_bounds__(kPcgBlock) k_dic_factor_blocks(PcgView v, double* rDout)
{
    extern __shared__ __align__(16) unsigned char rawF[];
    if (v.dicWarp) dicFactorT<true>(v, rawF, rDout); else dicFactorT<false>(v, rawF, rDout);
}

// one preconditioner application as a stand-alone kernel (stepwise / decomposed solver): z = M^-1 r, per-CTA partial of z.r
__global__ void __launch_bounds__(kPcgBlock) k_dic_apply_blocks(PcgView v, double* partial, const int* done)
{
    extern __shared__ __align__(16) unsigned char rawA[];
    __shared__ double red[kPcgBlock / 32];
    if (done && *done) return;
    double zr = dicBlocks(v, rawA);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zr += __shfl_down_sync(0xffffffffu, zr, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = zr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kPcgBlock / 32; ++k) t += red[k];
        partial[blockIdx.x] = t;
        partial[gridDim.x + blockIdx.x] = 0.0;
    }
}

template <int WT>
__global__ void __launch_bounds__(kPcgBlock) k_pcg(PcgView v)
{
    extern __shared__ __align__(16) unsigned char rawP[];
    cg::grid_group grid = cg::this_grid();
    const int n = v.n, W = WT ? WT : v.W;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    int slot = 0;
    // ---- wA = A x ; rA = b - wA ; xRef = average(x)
    double sx = 0.0;
    for (int c = tid; c < n; c += nth) {
        const double xc = v.x[c];
        double y = __ldg(&v.diag[c]) * xc;
#pragma unroll
        for (int j = 0; j < W; ++j) y += __ldg(&v.coef[(size_t)j * n + c]) * v.x[__ldg(&v.enc[(size_t)j * n + c]) >> 1];
        for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) y += __ldg(&v.tailCoef[q]) * v.x[__ldg(&v.tailEnc[q]) >> 1];
        v.w[c] = y;
        v.r[c] = __ldg(&v.b[c]) - y;
        v.p0[c] = 0.0;
        sx += xc;
    }
    double2 s = gridSum2(grid, sx, 0.0, v.partials, slot);
    const double xRef = s.x / (double)n;
    // ---- normFactor and initial residual
    double nf = 0.0, sr = 0.0;
    for (int c = tid; c < n; c += nth) {
        double sumA = __ldg(&v.diag[c]);
#pragma unroll
        for (int j = 0; j < W; ++j) sumA += __ldg(&v.coef[(size_t)j * n + c]);
        for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) sumA += __ldg(&v.tailCoef[q]);
        const double t = sumA * xRef;
        nf += fabs(v.w[c] - t) + fabs(__ldg(&v.b[c]) - t);
        const double rc = v.r[c];
        sr += fabs(rc);
        if (v.precond < 2) v.z[c] = v.precond ? __ldg(&v.rD[c]) * rc : rc;
    }
    s = gridSum2(grid, nf, sr, v.partials, slot);
    const double normFactor = s.x + 1e-20;
    const double res0 = s.y / normFactor;
    double res = res0;
    int it = 0;
    auto converged = [&]() { return res < v.tol || (v.relTol > 1e-20 && res < v.relTol * res0); };
    if (!converged() && v.maxIter > 0) {
        double zr = 0.0;
        if (v.precond == 2 && v.nBlocks > 0) zr = dicBlocks(v, rawP);
        else {
            if (v.precond == 2) dicSweeps<WT>(grid, v, tid, nth);
            for (int c = tid; c < n; c += nth) zr += v.z[c] * v.r[c];
        }
        s = gridSum2(grid, zr, 0.0, v.partials, slot);
        double wArA = s.x, wArAold = wArA;
        double* po = v.p0;
        double* pn = v.p1;
        while (true) {
            const double beta = (it == 0) ? 0.0 : wArA / wArAold;
            // ---- pA = wA + beta pA (own cell stored, neighbours recomputed) ; wA = A pA ; wApA
            double wp = 0.0;
            for (int c = tid; c < n; c += nth) {
                const double pc = v.z[c] + beta * po[c];
                double y = __ldg(&v.diag[c]) * pc;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const int o = __ldg(&v.enc[(size_t)j * n + c]) >> 1;
                    y += __ldg(&v.coef[(size_t)j * n + c]) * (v.z[o] + beta * po[o]);
                }
                for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) {
                    const int o = __ldg(&v.tailEnc[q]) >> 1;
                    y += __ldg(&v.tailCoef[q]) * (v.z[o] + beta * po[o]);
                }
                pn[c] = pc;
                v.w[c] = y;
                wp += y * pc;
            }
            s = gridSum2(grid, wp, 0.0, v.partials, slot);
            const double wApA = s.x;
            if (fabs(wApA) / normFactor < 1e-300) break;          // solverPerformance::checkSingularity
            const double alpha = wArA / wApA;
            // ---- x += alpha pA ; rA -= alpha wA ; residual ; (Jacobi / none) next wA = M^-1 rA and wArA
            sr = 0.0; zr = 0.0;
            for (int c = tid; c < n; c += nth) {
                v.x[c] += alpha * pn[c];
                const double rc = v.r[c] - alpha * v.w[c];
                v.r[c] = rc;
                sr += fabs(rc);
                if (v.precond < 2) {
                    const double zc = v.precond ? __ldg(&v.rD[c]) * rc : rc;
                    v.z[c] = zc;
                    zr += zc * rc;
                }
            }
            s = gridSum2(grid, sr, zr, v.partials, slot);
            res = s.x / normFactor;
            ++it;
            double* t = po; po = pn; pn = t;
            if (it >= v.maxIter || converged()) break;
            wArAold = wArA;
            wArA = s.y;
            if (v.precond == 2) {
                zr = 0.0;
                if (v.nBlocks > 0) zr = dicBlocks(v, rawP);
                else {
                    dicSweeps<WT>(grid, v, tid, nth);
                    for (int c = tid; c < n; c += nth) zr += v.z[c] * v.r[c];
                }
                s = gridSum2(grid, zr, 0.0, v.partials, slot);
                wArA = s.x;
            }
        }
    }
    if (tid == 0) { v.out->iters = it; v.out->res0 = res0; v.out->res = res; v.out->normFactor = normFactor; }
}

// DIC::calcReciprocalD: rD = diag ; for faces ascending: rD[u] -= upper^2/rD[l] ; rD = 1/rD   (level-scheduled, exact order)
__global__ void __launch_bounds__(kPcgBlock) k_dic_factor(PcgView v, double* rDraw)
{
    cg::grid_group grid = cg::this_grid();
    const int n = v.n, W = v.W;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int l = 0; l < v.nLevels; ++l) {
        for (int i = __ldg(&v.lvlOff[l]) + tid; i < __ldg(&v.lvlOff[l + 1]); i += nth) {
            const int c = __ldg(&v.lvlCells[i]);
            double rc = __ldg(&v.diag[c]);
            for (int j = 0; j < W; ++j) {
                const int e = __ldg(&v.enc[(size_t)j * n + c]);
                const double a = __ldg(&v.coef[(size_t)j * n + c]);
                if (e & 1) rc -= a * a / rDraw[e >> 1];
            }
            for (int q = __ldg(&v.tailOff[c]); q < __ldg(&v.tailOff[c + 1]); ++q) {
                const int e = __ldg(&v.tailEnc[q]);
                const double a = __ldg(&v.tailCoef[q]);
                if (e & 1) rc -= a * a / rDraw[e >> 1];
            }
            rDraw[c] = rc;
        }
        grid.sync();
    }
    for (int c = tid; c < n; c += nth) rDraw[c] = 1.0 / rDraw[c];
}

__global__ void k_recip(int n, const double* __restrict__ d, double* __restrict__ o)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) o[c] = 1.0 / d[c];
}

template <class K> int coopGrid(K kernel, size_t smem = 0)
{
    int dev = 0, sms = 148, perSM = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kPcgBlock, smem);
    if (perSM < 1) perSM = 1;
    if (perSM > 4) perSM = 4;
    return sms * perSM;
}

} // namespace

__global__ void k_fill_coef(int n, int W, const int* __restrict__ encFace, const double* __restrict__ faceCoef, double* __restrict__ coef,
                            int nTail, const int* __restrict__ tailFace, double* __restrict__ tailCoef)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n)
        for (int j = 0; j < W; ++j) {
            const int f = encFace[(size_t)j * n + c];
            coef[(size_t)j * n + c] = f >= 0 ? -faceCoef[f] : 0.0;
        }
    if (c < nTail) { const int f = tailFace[c]; tailCoef[c] = f >= 0 ? -faceCoef[f] : 0.0; }
}

int dicBlocksGrid(const PcgMatrix& A)
{
    const int units = A.dicWarp ? (A.nBlocks + kPcgBlock / 32 - 1) / (kPcgBlock / 32) : A.nBlocks;
    return std::max(1, std::min(units, 148 * (A.dicWarp ? 2 : 8)));
}
// z = M^-1 r with the block-local DIC of A as a stand-alone launch (stepwise / decomposed solver); partials: 2*dicBlocksGrid(A) doubles
void launchDicBlocks(const PcgMatrix& A, const double* r, double* z, double* partials, const int* done, cudaStream_t st)
{
    PcgView v = A.view(0, 0, 0);
    v.r = const_cast<double*>(r); v.z = z;
    const size_t smem = A.dicSmemBytes();
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_dic_apply_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_dic_apply_blocks<<<dicBlocksGrid(A), kPcgBlock, smem, st>>>(v, partials, done);
}

static void dicFactor(PcgMatrix& A, cudaStream_t st)
{
    PcgView v = A.view(0, 0, 0);
    double* raw = A.rD.p;
    if (A.nBlocks > 0) {
        const size_t smem = A.dicSmemBytes();
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_dic_factor_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_dic_factor_blocks<<<dicBlocksGrid(A), kPcgBlock, smem, st>>>(v, raw);
        return;
    }
    void* args[] = {&v, &raw};
    const int g = std::max(1, std::min(coopGrid(k_dic_factor), (A.n + kPcgBlock - 1) / kPcgBlock));
    QGD_CUDA(cudaLaunchCooperativeKernel((void*)k_dic_factor, dim3(g), dim3(kPcgBlock), args, 0, st));
}

void PcgMatrix::refresh(const double* faceCoef, const double* diagDev, cudaStream_t st)
{
    if (faceCoef && !encFace.n) throw Error(QGD_ERR_STATE, "PcgMatrix::refresh: matrix was built without face ids");
    const int nTail = (int)tailFace.n;
    if (faceCoef)      // nullptr: only the diagonal changed (e.g. a new deltaT)
        k_fill_coef<<<(std::max(n, nTail) + 255) / 256, 256, 0, st>>>(n, W, encFace.p, faceCoef, coef.p, nTail, tailFace.p, tailCoef.p);
    if (faceCoef && nBlocks > 0)       // the block-ordered copy of the in-block coefficients
        k_fill_coef<<<(n + 255) / 256, 256, 0, st>>>(n, Wb, bEncFace.p, faceCoef, bCoef.p, 0, nullptr, nullptr);
    QGD_CUDA(cudaMemcpyAsync(diag.p, diagDev, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (precond == 1) k_recip<<<(n + 255) / 256, 256, 0, st>>>(n, diag.p, rD.p);
    else if (precond == 2) dicFactor(*this, st);
    QGD_CUDA(cudaGetLastError());
}

void PcgMatrix::build(const HostMesh& h, const double* hdiag, const double* upper, int pc, cudaStream_t st, const std::vector<int>* faceInv)
{
    n = h.nCells;
    precond = pc;
    const int nI = h.nInternal;
    // rows in ascending polyMesh face order (HostMesh::cfEnc), internal faces only
    int maxRow = 0;
    std::vector<int> rowLen(n, 0);
    for (int c = 0; c < n; ++c) {
        int k = 0;
        for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q) if ((h.cfEnc[q] >> 1) < nI) ++k;
        rowLen[c] = k;
        maxRow = std::max(maxRow, k);
    }
    W = maxRow <= 4 ? 4 : (maxRow <= 6 ? 6 : 8);
    if (const char* v = getenv("QGD_ELL_MAXW")) W = std::min(W, atoi(v) <= 4 ? 4 : (atoi(v) <= 6 ? 6 : 8));   // test hook: force CSR tails
    std::vector<int> e((size_t)W * n), tOff(n + 1, 0), tEnc, level(n, 0), eFace((size_t)W * n, -1), tFace;
    std::vector<double> a((size_t)W * n, 0.0), tCoef;
    for (int c = 0; c < n; ++c) {
        for (int j = 0; j < W; ++j) e[(size_t)j * n + c] = c << 1;
        int j = 0;
        for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q) {
            const int f = h.cfEnc[q] >> 1;
            if (f >= nI) continue;
            const int lowerSide = h.cfEnc[q] & 1;                  // this cell is the face's neighbour (upper) cell
            const int o = lowerSide ? h.owner[f] : h.neighbour[f];
            const int enc1 = (o << 1) | lowerSide;
            if (lowerSide) level[c] = std::max(level[c], level[o] + 1);      // owner < neighbour: level[o] is final
            const int devFace = faceInv ? (*faceInv)[f] : f;
            if (j < W) { e[(size_t)j * n + c] = enc1; a[(size_t)j * n + c] = upper[f]; eFace[(size_t)j * n + c] = devFace; }
            else { tEnc.push_back(enc1); tCoef.push_back(upper[f]); tFace.push_back(devFace); }
            ++j;
        }
        tOff[c + 1] = (int)tEnc.size();
    }
    if (tEnc.empty()) { tEnc.push_back(0); tCoef.push_back(0.0); tFace.push_back(-1); }
    if (faceInv) { encFace.upload(eFace, st); tailFace.upload(tFace, st); }
    for (int f = 0; f < nI; ++f)
        if (h.owner[f] >= h.neighbour[f]) throw Error(QGD_ERR_INVALID, "PCG: mesh is not in upper-triangular order (owner < neighbour)");
    nLevels = 0;
    for (int c = 0; c < n; ++c) nLevels = std::max(nLevels, level[c] + 1);
    std::vector<int> lOff(nLevels + 1, 0), lCells(n);
    for (int c = 0; c < n; ++c) lOff[level[c] + 1]++;
    for (int l = 0; l < nLevels; ++l) lOff[l + 1] += lOff[l];
    {
        std::vector<int> pos(lOff.begin(), lOff.end() - 1);
        for (int c = 0; c < n; ++c) lCells[pos[level[c]]++] = c;
    }
    enc.upload(e, st); coef.upload(a, st); tailOff.upload(tOff, st); tailEnc.upload(tEnc, st); tailCoef.upload(tCoef, st);
    lvlOff.upload(lOff, st); lvlCells.upload(lCells, st);
    // ---- block-local DIC (HostMesh::pcgBlock): cells ordered (block, level inside the block, id); rows restricted to in-block entries
    nBlocks = 0; Wb = 0; maxBlockCells = 0;
    if (precond == 2 && !h.pcgBlock.empty()) {
        const std::vector<int>& blk = h.pcgBlock;
        int nb = 0;
        for (int c = 0; c < n; ++c) nb = std::max(nb, blk[c] + 1);
        // cells without a block (halo copies of a sub-mesh) form singleton blocks: their rows are never solved
        std::vector<int> bid(n);
        for (int c = 0; c < n; ++c) bid[c] = blk[c] >= 0 ? blk[c] : nb++;
        std::vector<int> lvl(n, 0), rowLenB(n, 0);
        for (int c = 0; c < n; ++c)
            for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q) {
                const int f = h.cfEnc[q] >> 1;
                if (f >= nI) continue;
                const int lowerSide = h.cfEnc[q] & 1;
                const int o = lowerSide ? h.owner[f] : h.neighbour[f];
                if (bid[o] != bid[c]) continue;
                ++rowLenB[c];
                if (lowerSide) lvl[c] = std::max(lvl[c], lvl[o] + 1);
            }
        for (int c = 0; c < n; ++c) Wb = std::max(Wb, rowLenB[c]);
        Wb = std::max(Wb, 1);
        std::vector<int> order(n);
        for (int c = 0; c < n; ++c) order[c] = c;
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            if (bid[x] != bid[y]) return bid[x] < bid[y];
            if (lvl[x] != lvl[y]) return lvl[x] < lvl[y];
            return x < y;
        });
        std::vector<int> pos(n), off(nb + 1, 0), lvlStart(nb + 1, 0), lvlOffB;
        for (int p = 0; p < n; ++p) { pos[order[p]] = p; off[bid[order[p]] + 1]++; }
        for (int b = 0; b < nb; ++b) { off[b + 1] += off[b]; maxBlockCells = std::max(maxBlockCells, off[b + 1] - off[b]); }
        for (int b = 0; b < nb; ++b) {
            lvlStart[b] = (int)lvlOffB.size();
            int cur = -1;
            for (int p = off[b]; p < off[b + 1]; ++p)
                while (cur < lvl[order[p]]) { lvlOffB.push_back(p); ++cur; }
            lvlOffB.push_back(off[b + 1]);
        }
        lvlStart[nb] = (int)lvlOffB.size();
        std::vector<int> be((size_t)Wb * n, -1), bf((size_t)Wb * n, -1);
        std::vector<double> bc((size_t)Wb * n, 0.0);
        for (int c = 0; c < n; ++c) {
            int j = 0;
            const int p = pos[c];
            for (int q = h.cfOff[c]; q < h.cfOff[c + 1]; ++q) {
                const int f = h.cfEnc[q] >> 1;
                if (f >= nI) continue;
                const int lowerSide = h.cfEnc[q] & 1;
                const int o = lowerSide ? h.owner[f] : h.neighbour[f];
                if (bid[o] != bid[c]) continue;
                be[(size_t)j * n + p] = ((pos[o] - off[bid[c]]) << 1) | lowerSide;
                bc[(size_t)j * n + p] = upper[f];
                bf[(size_t)j * n + p] = faceInv ? (*faceInv)[f] : f;
                ++j;
            }
        }
        nBlocks = nb;
        maxBlockLevels = 0;
        for (int b = 0; b < nb; ++b) maxBlockLevels = std::max(maxBlockLevels, lvlStart[b + 1] - lvlStart[b] - 1);
        dicWarp = dicUnitSmem() <= 12 * 1024;          // a warp per block, 8 blocks in flight per CTA
        bOff.upload(off, st); bLvlStart.upload(lvlStart, st); bLvlOff.upload(lvlOffB, st); bCells.upload(order, st);
        bEnc.upload(be, st); bCoef.upload(bc, st);
        if (faceInv) bEncFace.upload(bf, st);
        if (dicSmemBytes() > 200 * 1024)
            throw Error(QGD_ERR_INVALID, "PCG: a DIC block of " + std::to_string(maxBlockCells) + " cells with " + std::to_string(Wb) +
                                             " in-block neighbours does not fit the shared memory of one CTA; use smaller blocks");
    }
    diag.upload(std::vector<double>(hdiag, hdiag + n), st);
    rD.alloc(n); b.alloc(n); x.alloc(n); r.alloc(n); w.alloc(n); z.alloc(n); p0.alloc(n); p1.alloc(n);
    out.alloc(1);
    const size_t smem = dicSmemBytes();
    gridBlocks = std::min(std::min(coopGrid(k_pcg<4>, smem), coopGrid(k_pcg<6>, smem)), coopGrid(k_pcg<8>, smem));
    gridBlocks = std::max(1, std::min(gridBlocks, (n + kPcgBlock - 1) / kPcgBlock));
    partials.alloc(4 * (size_t)gridBlocks);
    if (precond == 1) k_recip<<<(n + 255) / 256, 256, 0, st>>>(n, diag.p, rD.p);
    else if (precond == 2) dicFactor(*this, st);
    QGD_CUDA(cudaGetLastError());
    QGD_CUDA(cudaStreamSynchronize(st));
}

PcgView PcgMatrix::view(double tol, double relTol, int maxIter) const
{
    PcgView v;
    v.n = n; v.W = W; v.enc = enc.p; v.coef = coef.p; v.tailOff = tailOff.p; v.tailEnc = tailEnc.p; v.tailCoef = tailCoef.p;
    v.diag = diag.p; v.rD = rD.p; v.b = bExternal ? bExternal : b.p; v.x = xExternal ? xExternal : x.p; v.r = r.p; v.w = w.p; v.z = z.p; v.p0 = p0.p; v.p1 = p1.p;
    v.partials = partials.p; v.nLevels = nLevels; v.lvlOff = lvlOff.p; v.lvlCells = lvlCells.p;
    v.nBlocks = nBlocks; v.Wb = Wb; v.maxBlockCells = maxBlockCells; v.dicWarp = dicWarp ? 1 : 0; v.dicUnitSmem = (int)dicUnitSmem();
    v.bOff = bOff.p; v.bLvlStart = bLvlStart.p; v.bLvlOff = bLvlOff.p; v.bCells = bCells.p; v.bEnc = bEnc.p; v.bCoef = bCoef.p;
    v.tol = tol; v.relTol = relTol; v.maxIter = maxIter; v.precond = precond; v.out = out.p;
    return v;
}

int PcgMatrix::solve(double tol, double relTol, int maxIter, cudaStream_t st)
{
    PcgView v = view(tol, relTol, maxIter);
    void* args[] = {&v};
    void* fn = (W == 4) ? (void*)k_pcg<4> : (W == 6 ? (void*)k_pcg<6> : (void*)k_pcg<8>);
    QGD_CUDA(cudaLaunchCooperativeKernel(fn, dim3(gridBlocks), dim3(kPcgBlock), args, dicSmemBytes(), st));
    return 1;
}

} // namespace qgd
