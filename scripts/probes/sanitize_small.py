"""small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every drain path, both plan
variants, packed groups, fixed + variable + by-sample entries, checked against the oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore
from oracle.oracle import COracle
from tests.helpers import random_valid_requests
co = COracle()
rng = np.random.default_rng(3)
store = PyDDStore(device=0)
for dtype, disp in ((np.uint8, 7), (np.float32, 1), (np.float32, 1024), (np.int64, 2)):
    shard = rng.integers(0, 256, size=(3000 * disp * np.dtype(dtype).itemsize,), dtype=np.uint8).view(dtype).reshape(3000, disp)
    name = f"v{disp}_{np.dtype(dtype).name}"
    store.add(name, shard)
    for B in (1, 50, 1500):
        st, ct = random_valid_requests(rng, [3000], B, max_count=20)
        exp, offs, bad, _ = co.get_batch([shard], st, ct)
        out = np.zeros(max(exp.size, 1), np.uint8)
        assert store.get_batch(name, st, ct, out=out) == exp.size and out[:exp.size].tobytes() == exp.tobytes()
        ok = [int(s) for s in st if s + 2 <= 3000]
        e2, _, _, _ = co.get_batch([shard], ok, [2] * len(ok))
        o2 = torch.zeros(max(e2.size, 16), dtype=torch.uint8, device="cuda")
        store.get_batch(name, torch.tensor(ok).cuda(), out=o2, count=2)
        assert o2[:e2.size].cpu().numpy().tobytes() == e2.tobytes()
# sample index, multi-array launch, overlapped queues (fixed and variable counts)
L = rng.integers(0, 30, size=400)
ss = np.concatenate([[0], np.cumsum(L)])
feat = rng.integers(0, 2**32, size=(int(ss[-1]), 4), dtype=np.uint32).view(np.float32)
edge = rng.integers(-9, 9, size=(int(2 * ss[-1]), 2), dtype=np.int64)
store.add("feat", feat); store.add("edge", edge)
store.set_sample_index("feat", ss[:-1], L); store.set_sample_index("edge", 2 * ss[:-1], 2 * L)
ids = rng.integers(0, 400, size=333)
ef = co.get_batch([feat], ss[ids], L[ids])[0]; ee = co.get_batch([edge], 2 * ss[ids], 2 * L[ids])[0]
of = torch.zeros(ef.size + 16, dtype=torch.uint8, device="cuda"); oe = torch.zeros(ee.size + 16, dtype=torch.uint8, device="cuda")
tot = store.get_samples_multi(["feat", "edge"], ids, [of, oe])
assert tot == [ef.size, ee.size] and of[:ef.size].cpu().numpy().tobytes() == ef.tobytes() and oe[:ee.size].cpu().numpy().tobytes() == ee.tobytes()
side = torch.cuda.Stream()
d_ids = torch.from_numpy(ids).cuda()
bufs = [torch.zeros_like(of), torch.zeros_like(of)]
torch.cuda.synchronize()
for k in range(6):
    store.get_samples("feat", d_ids, bufs[k & 1], stream=side.cuda_stream, wait=False, overlap=True)
assert store.wait() == ef.size and bufs[1][:ef.size].cpu().numpy().tobytes() == ef.tobytes()
fx = torch.from_numpy(rng.integers(0, 2990, size=500)).cuda()
ofx = [torch.zeros(500 * 2 * 28, dtype=torch.uint8, device="cuda") for _ in range(2)]
for k in range(6):
    store.get_batch("v7_uint8", fx, out=ofx[k & 1][:500 * 2 * 7], count=2, stream=side.cuda_stream, wait=False, overlap=True)
store.wait()
# round 2: large variable batch through the plan kernel (serialised and as an overlapped queue with scratch slots), the
# single-request kernels (doorbell round trips, host and device destinations), the on-device verifier
big = rng.integers(0, 2**32, size=(60000, 3), dtype=np.uint32).view(np.float32)
store.add("big", big)
st, ct = random_valid_requests(rng, [60000], 9000, max_count=12)
exp, offs, _, _ = co.get_batch([big], st, ct)
d_st, d_ct = torch.from_numpy(st).cuda(), torch.from_numpy(ct).cuda()
ob = [torch.zeros(exp.size + 16, dtype=torch.uint8, device="cuda") for _ in range(2)]
oo = [torch.zeros(9001, dtype=torch.int64, device="cuda") for _ in range(2)]
assert store.get_batch("big", d_st, d_ct, out=ob[0], offsets=oo[0]) == exp.size
assert ob[0][:exp.size].cpu().numpy().tobytes() == exp.tobytes() and oo[0].cpu().tolist() == offs.tolist()
torch.cuda.synchronize()
for k in range(7):
    store.get_batch("big", d_st, d_ct, out=ob[k & 1], offsets=oo[k & 1], stream=side.cuda_stream, wait=False, overlap=True)
assert store.wait() == exp.size
for k in (0, 1):
    assert ob[k][:exp.size].cpu().numpy().tobytes() == exp.tobytes() and oo[k].cpu().tolist() == offs.tolist()
for start, cnt in ((0, 1), (17, 3), (59999, 1), (1234, 0)):
    h = np.zeros((cnt, 3), np.float32); store.get("big", h, start); assert h.tobytes() == big[start:start + cnt].tobytes()
    d = torch.zeros((cnt, 3), dtype=torch.float32, device="cuda"); store.get("big", d, start)
    assert d.cpu().numpy().tobytes() == big[start:start + cnt].tobytes()
store.init("syn", 5000, 8, 4); store.synth_fill("syn", 99)
sid = torch.from_numpy(rng.integers(0, 5000, size=777)).cuda()
osyn = torch.zeros(777 * 32, dtype=torch.uint8, device="cuda")
store.get_batch("syn", sid, out=osyn, count=1)
assert store.synth_verify("syn", osyn, sid, seed=99)[:2] == (0, 777)
store.free(); store.close()
print("sanitize-ok")
