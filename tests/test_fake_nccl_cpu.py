"""The loopback transport of the single-GPU N-rank tests (tests/fake_nccl: a stand-in for the nine NCCL entry points libqgd_b200 binds
at run time) checked on CPU with host buffers: grouped and ungrouped point-to-point traffic between 2, 4 and 8 processes, messages
larger than a pipe buffer, ordered delivery per pair, all-reduces, size-mismatch detection."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_nccl")


def build_fake_nccl():
    out, src = os.path.join(FAKE, "libnccl.so.2"), os.path.join(FAKE, "fake_nccl.cpp")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", "/usr/local/cuda/include", "-o", out, src,
                               "-L/usr/local/cuda/lib64", "-lcudart"])
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fake_nccl_transport_between_processes(world, tmp_path):
    build_fake_nccl()
    env = dict(os.environ, FAKE_NCCL_HOST="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "fake_nccl_host_worker.py"), str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=240)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"FAKE_NCCL_OK {r}" in o, o[-2000:]
