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
store.free(); store.close()
print("sanitize-ok")
