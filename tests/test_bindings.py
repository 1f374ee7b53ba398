"""The two reference-shaped layers above the C-ABI: the C++ class (include/ddstore_b200.hpp) and the Cython
module `pyddstore` (ddstore_b200/cython). CPU part: they build / import and refuse to run without a GPU.
GPU part: the reference's own demos re-expressed on them."""
import os
import subprocess
import sys
import uuid

import numpy as np
import pytest

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CYDIR = os.path.join(ROOT, "ddstore_b200", "cython")
G = load_golden()


def _build_cpp_demo(tmp_path):
    exe = str(tmp_path / "demo_ddstore")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "demo_ddstore.cpp"),
           "-L", os.path.join(ROOT, "ddstore_b200"), "-lddstore_b200", f"-Wl,-rpath,{os.path.join(ROOT, 'ddstore_b200')}",
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _pyddstore():
    if CYDIR not in sys.path:
        sys.path.insert(0, CYDIR)
    import pyddstore
    return pyddstore


def test_cpp_header_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build_cpp_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "0", "1", "k"], capture_output=True, text=True)
    assert r.returncode == 1 and "No usable CUDA device" in r.stdout


def test_cython_module_surface():
    import torch
    m = _pyddstore()
    for name in ("add", "get", "epoch_begin", "epoch_end", "free", "init", "update", "get_batch"):
        assert hasattr(m.PyDDStore, name)  # src/pyddstore.pyx:65-131
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="No usable CUDA device"):
            m.PyDDStore()


@pytest.mark.gpu
def test_cpp_demo_two_processes(tmp_path):
    exe = _build_cpp_demo(tmp_path)
    key = "cpp" + uuid.uuid4().hex[:10]
    procs = [subprocess.Popen([exe, str(r), "2", key, "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        want = G["demo_cxx"][r]["got"]  # test/demo.cxx at P=2: rank0 {13,14}, rank1 {3,4}
        assert f"{r}: start {G['demo_cxx'][r]['start']} got {want[0]:g} {want[1]:g} 0 0" in o, o
        assert f"{r}: itemsize 8 disp 2 lenlist_last 4" in o
        assert o.count("invalid_argument: Invalid data type") == 1
        assert o.count("invalid_argument: Invalid start on target") == 1
        assert o.count("invalid_argument: Invalid count on target") == 1
        assert "logic_error: Fence already activated" in o and "logic_error: Fence is not activated" in o
        assert f"{r}: batch 64 bytes first 13 last 2" in o


@pytest.mark.gpu
def test_cython_pyddstore_demo_py_flow():
    """test/demo.py:29-56 on the Cython module: shard == rank+1, random single-row gets inside fences, mean check"""
    import threading
    m = _pyddstore()
    from ddstore_b200 import ShmComm
    P, num, dim, nbatch = 2, 4096, 64, 32
    key = "cy" + uuid.uuid4().hex[:10]
    errs = []

    def run(rank):
        try:
            comm = ShmComm(key, rank, P)
            ddstore = m.PyDDStore(comm, method=0, device=0)
            arr = np.ones((num, dim), dtype=np.float64) * (rank + 1)
            ddstore.add("var", arr)
            rng = np.random.default_rng(rank)
            idx_list, buff_list = [], []
            for _ in range(nbatch):
                ddstore.epoch_begin()
                idx = int(rng.integers(num * P))
                buff = np.zeros((1, dim), dtype=np.float64)
                ddstore.get("var", buff, idx)
                ddstore.epoch_end()
                idx_list.append(idx)
                buff_list.append(buff)
            for i, idx in enumerate(idx_list):
                expected = idx // num + 1
                assert np.mean(buff_list[i]) == expected, (np.mean(buff_list[i]), expected)
            out = np.zeros((nbatch, dim), np.float64)
            n = ddstore.get_batch("var", idx_list, out=out)
            assert n == out.nbytes and out.tobytes() == np.concatenate(buff_list).tobytes()
            with pytest.raises(ValueError, match="Invalid count on target"):
                ddstore.get("var", np.zeros((1, dim)), num * P)
            with pytest.raises(KeyError):
                ddstore.get("nope", np.zeros((1, dim)), 0)
            ddstore.init("z", 8, 3, 4)
            ddstore.update("z", np.arange(6, dtype=np.float32).reshape(2, 3), 2)
            got = np.zeros((8, 3), np.float32)
            ddstore.get("z", got, rank * 8)
            assert got[2:4].tobytes() == np.arange(6, dtype=np.float32).tobytes() and got[:2].sum() == 0
            ddstore.free()
            del ddstore
            comm.close()
        except BaseException:  # noqa: BLE001
            import traceback
            errs.append(traceback.format_exc())

    th = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert not errs, "\n".join(errs)
