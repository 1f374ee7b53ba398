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


def test_collective_push_fetch_matches_pull(coracle):
    """dds_get_batch_push (every rank publishes its start rows, every OWNER pushes its rows into the requesters'
    windows over NVLink): byte for byte the rows the one-sided pull delivers, for several steps (the windows' two
    buffers alternate), different batch sizes per rank, multi-row requests, 4-byte-phase rows; an invalid request is
    reported to the REQUESTER with its index and the reference's text; thread-ranks sharing a GPU are refused."""
    import torch
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    P = min(n, 4)
    rng = np.random.default_rng(4242)
    for dtype, disp, count in ((np.float32, 1024, 1), (np.float32, 3, 5), (np.uint8, 7, 2)):
        nrows, shards = random_world(rng, P, dtype, disp, max_rows=6000, allow_empty=False)
        ll = O.np_lenlist(nrows)
        steps = 5
        reqs = [[random_valid_requests(rng, ll, 1500 + 301 * r + 17 * t, max_count=count)[0] for t in range(steps)] for r in range(P)]
        for r in range(P):  # keep every request of `count` rows inside its owner
            for t in range(steps):
                s = reqs[r][t]
                own = np.searchsorted(ll, s, side="right")
                s[:] = np.minimum(s, ll[own] - count)
                lo = np.concatenate([[0], ll])[own]
                s[:] = np.maximum(s, lo)
        row = disp * np.dtype(dtype).itemsize

        def body(store, r):
            dev = torch.device("cuda", r)
            torch.cuda.set_device(r)
            store.add("v", shards[r])
            store.push_setup(4000, 4000 * count * row)
            st = torch.cuda.Stream(device=dev)
            for t in range(steps):
                ids = torch.from_numpy(reqs[r][t]).to(dev)
                torch.cuda.synchronize(dev)
                got = store.get_batch_push("v", ids, count=count, stream=st.cuda_stream)
                store.wait()
                exp, _, bad, _ = coracle.get_batch(shards, reqs[r][t], np.full(len(reqs[r][t]), count))
                assert bad == -1 and got.cpu().numpy().tobytes() == exp.tobytes(), (r, t)
                pull = torch.zeros(exp.size, dtype=torch.uint8, device=dev)
                store.get_batch("v", ids, out=pull, count=count)
                assert torch.equal(pull, got)
            # rank 1 asks for a row that does not exist: only rank 1 sees the error, at the right index
            ids = torch.from_numpy(reqs[r][0]).to(dev).clone()
            if r == 1:
                ids[77] = int(ll[-1])
            torch.cuda.synchronize(dev)
            store.get_batch_push("v", ids, count=count, stream=st.cuda_stream)
            if r == 1:
                with pytest.raises(ValueError, match="Invalid count on target"):
                    store.wait()
                assert store.last_bad_index == 77
            else:
                store.wait()
            ids = torch.from_numpy(reqs[r][1]).to(dev)
            torch.cuda.synchronize(dev)
            got = store.get_batch_push("v", ids, count=count, stream=st.cuda_stream)  # the step after an error works
            store.wait()
            exp, _, _, _ = coracle.get_batch(shards, reqs[r][1], np.full(len(reqs[r][1]), count))
            assert got.cpu().numpy().tobytes() == exp.tobytes()
            return True

        assert all(run_world(P, body, devices=list(range(P))))

    def shared(store, r):
        store.add("v", np.zeros((10, 4), np.float32))
        with pytest.raises(ValueError, match="GPU of its own"):
            store.push_setup(16, 4096)
        return True

    assert all(run_world(2, shared, devices=[0, 0]))
