"""Worker of tests/test_fake_nccl_cpu.py: N processes drive tests/fake_nccl/libnccl.so.2 with HOST buffers (FAKE_NCCL_HOST=1): grouped
send / recv between every pair (two messages per pair in one group, sizes from a few bytes to several MB), ungrouped exchanges,
all-reduces (max, min, sum) - the patterns libqgd_b200 issues."""
import ctypes as C
import os
import sys
import time

import numpy as np

rank, world, rdv = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_nccl", "libnccl.so.2"))


class UID(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UID, C.c_int]
L.ncclSend.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
L.ncclRecv.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
L.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
L.ncclCommDestroy.argtypes = [C.c_void_p]
F64, SUM, MAX, MIN = 8, 0, 2, 3
uid = UID()
idf = os.path.join(rdv, "id.bin")
if rank == 0:
    os.environ["FAKE_NCCL_DIR"] = rdv
    assert L.ncclGetUniqueId(C.byref(uid)) == 0
    with open(idf + ".tmp", "wb") as f:
        f.write(bytes(uid))
    os.rename(idf + ".tmp", idf)
else:
    while not os.path.exists(idf):
        time.sleep(0.01)
    C.memmove(C.byref(uid), open(idf, "rb").read(), 128)
comm = C.c_void_p()
assert L.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0


def payload(src, dst, k, n):
    return np.arange(n, dtype=np.float64) * 0.5 + 1000.0 * src + 10.0 * dst + k


for rep, n in enumerate((1, 7, 1000, 400000)):                 # 8 B ... 3.2 MB, beyond the 64 KB pipe buffer
    sends = {(p, k): payload(rank, p, k + 2 * rep, n + k) for p in range(world) if p != rank for k in range(2)}
    recvs = {(p, k): np.zeros(n + k) for p in range(world) if p != rank for k in range(2)}
    assert L.ncclGroupStart() == 0
    for (p, k), a in sends.items():
        assert L.ncclSend(a.ctypes.data, a.size, F64, p, comm, None) == 0
    for (p, k), a in recvs.items():
        assert L.ncclRecv(a.ctypes.data, a.size, F64, p, comm, None) == 0
    assert L.ncclGroupEnd() == 0
    for (p, k), a in recvs.items():
        assert np.array_equal(a, payload(p, rank, k + 2 * rep, n + k)), (rank, p, k, n)
# ungrouped ring: everyone sends right and receives from the left (send first on even ranks, receive first on odd ones)
right, left = (rank + 1) % world, (rank - 1) % world
a, b = payload(rank, right, 99, 5000), np.zeros(5000)
if world > 1:
    if rank % 2 == 0:
        assert L.ncclSend(a.ctypes.data, a.size, F64, right, comm, None) == 0
        assert L.ncclRecv(b.ctypes.data, b.size, F64, left, comm, None) == 0
    else:
        assert L.ncclRecv(b.ctypes.data, b.size, F64, left, comm, None) == 0
        assert L.ncclSend(a.ctypes.data, a.size, F64, right, comm, None) == 0
    assert np.array_equal(b, payload(left, rank, 99, 5000))
# all-reduces, grouped like the Courant max / tau min of the step, and a sum in place
x, y = np.array([float(rank) + 0.25]), np.array([10.0 - rank])
assert L.ncclGroupStart() == 0
assert L.ncclAllReduce(x.ctypes.data, x.ctypes.data, 1, F64, MAX, comm, None) == 0
assert L.ncclAllReduce(y.ctypes.data, y.ctypes.data, 1, F64, MIN, comm, None) == 0
assert L.ncclGroupEnd() == 0
assert x[0] == world - 1 + 0.25 and y[0] == 10.0 - (world - 1)
z = np.array([1.0 * (rank + 1), 0.5, -2.0 * rank])
assert L.ncclAllReduce(z.ctypes.data, z.ctypes.data, 3, F64, SUM, comm, None) == 0
assert np.allclose(z, [world * (world + 1) / 2, 0.5 * world, -2.0 * world * (world - 1) / 2])
# a message of the wrong size is an error, not silent corruption
if world > 1:
    if rank == 0:
        q = np.zeros(3)
        assert L.ncclSend(q.ctypes.data, 3, F64, 1, comm, None) == 0
    if rank == 1:
        q = np.zeros(5)
        assert L.ncclRecv(q.ctypes.data, 5, F64, 0, comm, None) != 0
L.ncclCommDestroy(comm)
print("FAKE_NCCL_OK", rank, flush=True)
