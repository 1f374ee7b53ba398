"""Multi-GPU parity (-m gpu, skipped on a single-GPU box): the same byte-exact checks with the shards on
DIFFERENT GPUs, so the gather kernel's TMA loads really cross NVLink -- (a) thread-ranks of one process
(cudaDeviceEnablePeerAccess path) and (b) one process per GPU (CUDA IPC path, the deployment shape)."""
import os
import subprocess
import sys
import uuid

import numpy as np
import pytest

from oracle import oracle as O
from tests.gpu_helpers import run_world
from tests.helpers import random_valid_requests, random_world

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("dtype,disp", [(np.float32, 1024), (np.float32, 1), (np.uint8, 7), (np.int64, 2)])
def test_thread_ranks_on_distinct_gpus(coracle, dtype, disp):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    P = min(n, 4)
    rng = np.random.default_rng(77 + disp)
    nrows, shards = random_world(rng, P, dtype, disp, max_rows=3000, allow_empty=False)
    ll = O.np_lenlist(nrows)
    starts, counts = random_valid_requests(rng, ll, 2000, max_count=50)
    exp, exp_offs, bad, _ = coracle.get_batch(shards, starts, counts)
    assert bad == -1

    def body(store, r):
        store.add("v", shards[r])
        out = np.zeros(exp.size, np.uint8)
        assert store.get_batch("v", starts, counts, out=out) == exp.size
        assert out.tobytes() == exp.tobytes()
        return True

    assert all(run_world(P, body, devices=list(range(P))))


IPC_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from ddstore_b200 import PyDDStore, ShmComm
from oracle.oracle import np_synth_rows
rank, P, key = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
torch.cuda.set_device(rank)
comm = ShmComm(key, rank, P)
store = PyDDStore(comm, device=rank)
per, disp = 50000, 1024
store.init("x", per + 17 * rank, disp, 4)
store.synth_fill("x", 0xDD5)
store.init("lab", per + 17 * rank, 1, 4)       # tiny rows: 4-byte requests
store.synth_fill("lab", 0xDD6)
ll = store.query("x")["lenlist"]
total = ll[-1]
rng = np.random.default_rng(99 + rank)
idx = rng.integers(0, total, size=20000)
out = torch.empty(len(idx) * disp, dtype=torch.float32, device=f"cuda:{{rank}}")
store.epoch_begin()
n = store.get_batch("x", idx, out=out, count=1)
lab = np.zeros(len(idx), np.float32)
store.get_batch("lab", idx, out=lab, count=1)
store.epoch_end()
got = out.cpu().numpy().reshape(len(idx), disp)
for b0 in range(0, len(idx), 4000):
    exp = np.concatenate([np_synth_rows(0xDD5, int(i), 1, disp, np.float32) for i in idx[b0:b0 + 4000:97]])
    assert got[b0:b0 + 4000:97].tobytes() == exp.tobytes()
owners = np.searchsorted(np.array(ll), idx, side="right")
assert len(set(owners.tolist())) == P, "batch did not touch every owner"
expl = np.concatenate([np_synth_rows(0xDD6, int(i), 1, 1, np.float32) for i in idx[:500]]).reshape(-1)
assert lab[:500].tobytes() == expl.tobytes()
# variable-length, unaligned: rows of the float32 variable re-read as runs of a few rows
st = rng.integers(0, total - 8, size=3000); ct = rng.integers(0, 4, size=3000)
ok = [(int(s), int(c)) for s, c in zip(st, ct) if np.searchsorted(np.array(ll), s, side="right") == np.searchsorted(np.array(ll), s + max(c, 1) - 1, side="right")]
st = np.array([o[0] for o in ok]); ct = np.array([o[1] for o in ok])
o2 = np.zeros(int(ct.sum()) * disp * 4, np.uint8)
store.get_batch("x", st, ct, out=o2)
pos = 0
for s, c in ok[:200]:
    if c:
        e = np_synth_rows(0xDD5, s, c, disp, np.float32).tobytes()
        assert o2[pos:pos + len(e)].tobytes() == e
    pos += c * disp * 4
store.free(); store.close(); comm.close()
print("ipc-ok", rank)
"""


def test_one_process_per_gpu_cuda_ipc(tmp_path):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    P = min(n, 8)
    script = tmp_path / "ipc_world.py"
    script.write_text(IPC_SCRIPT.format(root=ROOT))
    key = "ipc" + uuid.uuid4().hex[:10]
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(P), key], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(P)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ipc-ok {r}" in o, f"rank {r}:\n{o[-3000:]}"
