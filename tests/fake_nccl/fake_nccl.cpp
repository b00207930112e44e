// TEST INFRASTRUCTURE ONLY - never shipped, never linked into the product.
//
// A stand-in for libnccl.so.2 that moves the messages of libqgd_b200's decomposed runs between PROCESSES THAT SHARE ONE GPU
// (or, with FAKE_NCCL_HOST=1, between host buffers), so that the N-rank code path of the library - pack / unpack kernels, the
// exchange lists, the global decisions that steer collectives (the N = 8 hang of round 2), the all-reduced time-step control -
// can be exercised on a box with a single GPU.  The library binds NCCL at run time (dlopen("libnccl.so.2")); the loopback tests
// put this directory first on LD_LIBRARY_PATH.  NCCL itself is not emulated beyond the nine entry points the library binds:
//   ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclSend / ncclRecv / ncclAllReduce / ncclGroupStart / ncclGroupEnd /
//   ncclGetErrorString
// Transport: one named pipe per ordered pair of ranks under a rendezvous directory named in the unique id.  Every call is
// executed synchronously: the stream is synchronised, device buffers are staged through host memory (cudaMemcpy), sends run on
// their own threads so that matched send / recv pairs can never wait on each other in a cycle; a send without its matching
// receive (a rank that skipped an exchange) blocks - which is exactly the failure the tests are there to catch (timeout).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" {

typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7,
               ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct FakeComm { std::string dir; int n = 1, rank = 0; std::vector<int> wfd, rfd; };
typedef FakeComm* ncclComm_t;

}

namespace {

bool hostMode() { const char* v = getenv("FAKE_NCCL_HOST"); return v && atoi(v) != 0; }

struct Op { int kind; const void* src; void* dst; size_t bytes; int peer; ncclRedOp_t red; cudaStream_t st; FakeComm* c; };   // 0 send, 1 recv, 2 allreduce(double)
thread_local int g_depth = 0;
thread_local std::vector<Op> g_queue;

std::string fifo(const FakeComm* c, int from, int to) { return c->dir + "/f_" + std::to_string(from) + "_" + std::to_string(to); }

bool writeAll(int fd, const void* p, size_t n)
{
    const char* b = static_cast<const char*>(p);
    while (n) { ssize_t k = write(fd, b, n); if (k < 0) { if (errno == EINTR) continue; return false; } b += k; n -= (size_t)k; }
    return true;
}
bool readAll(int fd, void* p, size_t n)
{
    char* b = static_cast<char*>(p);
    while (n) { ssize_t k = read(fd, b, n); if (k < 0) { if (errno == EINTR) continue; return false; } if (k == 0) return false; b += k; n -= (size_t)k; }
    return true;
}
int wfd(FakeComm* c, int to) { if (c->wfd[to] < 0) c->wfd[to] = open(fifo(c, c->rank, to).c_str(), O_WRONLY); return c->wfd[to]; }
int rfd(FakeComm* c, int from) { if (c->rfd[from] < 0) c->rfd[from] = open(fifo(c, from, c->rank).c_str(), O_RDONLY); return c->rfd[from]; }

bool sendMsg(FakeComm* c, int to, const void* host, size_t bytes)
{
    const int fd = wfd(c, to);
    unsigned long long hdr = bytes;
    return fd >= 0 && writeAll(fd, &hdr, sizeof(hdr)) && writeAll(fd, host, bytes);
}
bool recvMsg(FakeComm* c, int from, void* host, size_t bytes)
{
    const int fd = rfd(c, from);
    unsigned long long hdr = 0;
    if (fd < 0 || !readAll(fd, &hdr, sizeof(hdr))) return false;
    if (hdr != bytes) { fprintf(stderr, "fake_nccl: rank %d expected %zu bytes from %d, got %llu\n", c->rank, bytes, from, hdr); return false; }
    return readAll(fd, host, bytes);
}

ncclResult_t toHost(void* h, const void* d, size_t n)
{
    if (hostMode()) { std::memcpy(h, d, n); return ncclSuccess; }
    return cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost) == cudaSuccess ? ncclSuccess : ncclUnhandledCudaError;
}
ncclResult_t toDev(void* d, const void* h, size_t n)
{
    if (hostMode()) { std::memcpy(d, h, n); return ncclSuccess; }
    return cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) == cudaSuccess ? ncclSuccess : ncclUnhandledCudaError;
}

// executes a batch: all point-to-point sends concurrently (one thread each), the receives in order on this thread; all-reduces
// (gather to rank 0 in rank order, reduce, broadcast) one after the other
ncclResult_t run(std::vector<Op>& ops)
{
    if (!hostMode())
        for (const Op& o : ops) if (cudaStreamSynchronize(o.st) != cudaSuccess) return ncclUnhandledCudaError;
    std::vector<std::vector<char>> sbuf(ops.size()), rbuf(ops.size());
    std::vector<std::thread> th;
    std::vector<int> ok(ops.size(), 1);
    int maxPeer = -1;
    for (size_t i = 0; i < ops.size(); ++i) {
        Op& o = ops[i];
        if (o.kind != 0) continue;
        sbuf[i].resize(o.bytes);
        if (toHost(sbuf[i].data(), o.src, o.bytes) != ncclSuccess) return ncclUnhandledCudaError;
        if (o.peer > maxPeer) maxPeer = o.peer;
    }
    // one sender thread per peer: the messages to one peer keep their order on the pipe, different peers never wait on each other
    for (int peer = 0; peer <= maxPeer; ++peer) {
        bool any = false;
        for (const Op& o : ops) any = any || (o.kind == 0 && o.peer == peer);
        if (!any) continue;
        th.emplace_back([&, peer] {
            for (size_t i = 0; i < ops.size(); ++i)
                if (ops[i].kind == 0 && ops[i].peer == peer) ok[i] = sendMsg(ops[i].c, peer, sbuf[i].data(), ops[i].bytes) ? 1 : 0;
        });
    }
    ncclResult_t res = ncclSuccess;
    for (size_t i = 0; i < ops.size(); ++i) {
        Op& o = ops[i];
        if (o.kind != 1) continue;
        rbuf[i].resize(o.bytes);
        if (!recvMsg(o.c, o.peer, rbuf[i].data(), o.bytes)) { res = ncclSystemError; break; }
        if (toDev(o.dst, rbuf[i].data(), o.bytes) != ncclSuccess) { res = ncclUnhandledCudaError; break; }
    }
    for (std::thread& t : th) t.join();
    for (int v : ok) if (!v) res = ncclSystemError;
    if (res != ncclSuccess) return res;
    for (Op& o : ops) {
        if (o.kind != 2) continue;
        FakeComm* c = o.c;
        const size_t n = o.bytes / sizeof(double);
        std::vector<double> mine(n), acc(n), tmp(n);
        if (toHost(mine.data(), o.src, o.bytes) != ncclSuccess) return ncclUnhandledCudaError;
        if (c->rank == 0) {
            acc = mine;
            for (int r = 1; r < c->n; ++r) {
                if (!recvMsg(c, r, tmp.data(), o.bytes)) return ncclSystemError;
                for (size_t k = 0; k < n; ++k)
                    acc[k] = o.red == ncclSum ? acc[k] + tmp[k] : (o.red == ncclMax ? (tmp[k] > acc[k] ? tmp[k] : acc[k]) : (o.red == ncclMin ? (tmp[k] < acc[k] ? tmp[k] : acc[k]) : acc[k] * tmp[k]));
            }
            for (int r = 1; r < c->n; ++r) if (!sendMsg(c, r, acc.data(), o.bytes)) return ncclSystemError;
        } else {
            if (!sendMsg(c, 0, mine.data(), o.bytes) || !recvMsg(c, 0, acc.data(), o.bytes)) return ncclSystemError;
        }
        if (toDev(o.dst, acc.data(), o.bytes) != ncclSuccess) return ncclUnhandledCudaError;
    }
    return ncclSuccess;
}

// a cudaMemcpy from pageable host memory may return before the DMA has landed and the library's streams are non-blocking: make
// every staged copy visible before the caller enqueues the kernels that read it
ncclResult_t runAndSettle(std::vector<Op>& ops)
{
    const ncclResult_t r = run(ops);
    if (r == ncclSuccess && !hostMode() && cudaDeviceSynchronize() != cudaSuccess) return ncclUnhandledCudaError;
    return r;
}

ncclResult_t submit(const Op& o)
{
    g_queue.push_back(o);
    if (g_depth > 0) return ncclSuccess;
    std::vector<Op> ops;
    ops.swap(g_queue);
    return runAndSettle(ops);
}

} // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId* id)
{
    std::memset(id, 0, sizeof(*id));
    const char* base = getenv("FAKE_NCCL_DIR");
    std::string d = std::string(base ? base : "/tmp") + "/fakenccl_" + std::to_string((long)getpid()) + "_" + std::to_string((long)time(nullptr));
    if (d.size() + 1 > sizeof(id->internal)) return ncclInvalidArgument;
    if (mkdir(d.c_str(), 0700) != 0 && errno != EEXIST) return ncclSystemError;
    std::memcpy(id->internal, d.c_str(), d.size() + 1);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank)
{
    id.internal[sizeof(id.internal) - 1] = 0;
    FakeComm* c = new FakeComm();
    c->dir = id.internal; c->n = nranks; c->rank = rank;
    c->wfd.assign(nranks, -1); c->rfd.assign(nranks, -1);
    for (int to = 0; to < nranks; ++to)
        if (to != rank && mkfifo(fifo(c, rank, to).c_str(), 0600) != 0 && errno != EEXIST) { delete c; return ncclSystemError; }
    // wait until every rank has created its outgoing pipes
    for (int from = 0; from < nranks; ++from)
        for (int to = 0; to < nranks; ++to) {
            if (from == to) continue;
            struct stat sb;
            int tries = 0;
            while (stat(fifo(c, from, to).c_str(), &sb) != 0) { usleep(2000); if (++tries > 150000) { delete c; return ncclSystemError; } }
        }
    *comm = c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c)
{
    if (!c) return ncclSuccess;
    for (int fd : c->wfd) if (fd >= 0) close(fd);
    for (int fd : c->rfd) if (fd >= 0) close(fd);
    delete c;
    return ncclSuccess;
}

static size_t typeSize(ncclDataType_t t) { return t == ncclFloat64 || t == ncclInt64 || t == ncclUint64 ? 8 : (t == ncclInt8 || t == ncclUint8 ? 1 : (t == ncclFloat16 ? 2 : 4)); }

ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{
    if (!c || peer < 0 || peer >= c->n || peer == c->rank) return ncclInvalidArgument;
    return submit(Op{0, buf, nullptr, count * typeSize(t), peer, ncclSum, st, c});
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{
    if (!c || peer < 0 || peer >= c->n || peer == c->rank) return ncclInvalidArgument;
    return submit(Op{1, nullptr, buf, count * typeSize(t), peer, ncclSum, st, c});
}
ncclResult_t ncclAllReduce(const void* sendbuf, void* recvbuf, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, cudaStream_t st)
{
    if (!c || t != ncclFloat64) return ncclInvalidArgument;          // the library reduces doubles only
    return submit(Op{2, sendbuf, recvbuf, count * sizeof(double), -1, op, st, c});
}
ncclResult_t ncclGroupStart() { ++g_depth; return ncclSuccess; }
ncclResult_t ncclGroupEnd()
{
    if (g_depth <= 0) return ncclInvalidArgument;
    if (--g_depth > 0) return ncclSuccess;
    std::vector<Op> ops;
    ops.swap(g_queue);
    return ops.empty() ? ncclSuccess : runAndSettle(ops);
}
const char* ncclGetErrorString(ncclResult_t r)
{
    switch (r) {
        case ncclSuccess: return "no error";
        case ncclUnhandledCudaError: return "fake_nccl: CUDA error";
        case ncclSystemError: return "fake_nccl: transport error (pipe closed or message size mismatch)";
        case ncclInvalidArgument: return "fake_nccl: invalid argument";
        default: return "fake_nccl: internal error";
    }
}

} // extern "C"
