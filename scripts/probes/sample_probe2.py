"""probe: host-side cost per async call and device time per step, explicit vs by-sample"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore
store = PyDDStore(device=0)
nsamp = 1000000
L = np.random.default_rng(42).integers(100, 10001, size=nsamp)
ss = np.concatenate([[0], np.cumsum(L)])
store.init("x", int(ss[-1]), 1, 4)
store.synth_fill("x", 1)
d_start, d_len = torch.from_numpy(ss[:-1].copy()).cuda(), torch.from_numpy(L).cuda()
store.set_sample_index("x", d_start, d_len)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
for B in (4096, 16384):
    ids = torch.from_numpy(np.random.default_rng(1).integers(0, nsamp, size=B)).cuda()
    st, ct = d_start[ids].contiguous(), d_len[ids].contiguous()
    out = torch.empty(int(ct.sum().item()) * 4, dtype=torch.uint8, device="cuda")
    offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
    fns = {"explicit": lambda: store.get_batch("x", st, ct, out=out, offsets=offs, stream=side.cuda_stream, wait=False),
           "by-sample": lambda: store.get_samples("x", ids, out, offsets=offs, stream=side.cuda_stream, wait=False)}
    for rep in range(2):
        for name, fn in fns.items():
            for _ in range(3): fn()
            store.wait(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(20): fn()
            t1 = time.perf_counter(); e1.record()
            store.wait(); torch.cuda.synchronize()
            print(f"B={B} {name:10s} host submit {1e6*(t1-t0)/20:7.1f} us/call   device {1e3*e0.elapsed_time(e1)/20:7.1f} us/step", flush=True)
store.free(); store.close()
