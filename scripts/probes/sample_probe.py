"""probe: explicit (starts,counts) vs by-sample-id batches, B=4096, for an ncu launch list"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore
store = PyDDStore(device=0)
nsamp = 200000
L = np.random.default_rng(42).integers(100, 10001, size=nsamp)
ss = np.concatenate([[0], np.cumsum(L)])
store.init("x", int(ss[-1]), 1, 4)
store.synth_fill("x", 1)
d_start, d_len = torch.from_numpy(ss[:-1].copy()).cuda(), torch.from_numpy(L).cuda()
store.set_sample_index("x", d_start, d_len)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ids = torch.from_numpy(np.random.default_rng(1).integers(0, nsamp, size=B)).cuda()
st, ct = d_start[ids].contiguous(), d_len[ids].contiguous()
out = torch.empty(int(ct.sum().item()) * 4, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(4):
    store.get_batch("x", st, ct, out=out)
for _ in range(4):
    store.get_samples("x", ids, out)
store.free(); store.close()
