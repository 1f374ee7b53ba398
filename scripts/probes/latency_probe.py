"""probe: per-call latency of the synchronous entry points (the legacy per-sample loader pattern)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore
store = PyDDStore(device=0)
store.init("x", 1_000_000, 1024, 4); store.synth_fill("x", 1)
store.init("lab", 1_000_000, 1, 4); store.synth_fill("lab", 2)
rng = np.random.default_rng(0)
def timeit(label, fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{label:58s} {dt*1e6:8.1f} us/call", flush=True)
hb = np.zeros((1, 1024), np.float32); hl = np.zeros((1, 1), np.float32)
pin = torch.zeros((1, 1024), dtype=torch.float32).pin_memory(); pinn = pin.numpy()
db = torch.zeros((1, 1024), dtype=torch.float32, device="cuda")
idx = rng.integers(0, 1_000_000, size=4096)
k = [0]
def nxt():
    k[0] = (k[0] + 1) % 4096; return int(idx[k[0]])
timeit("get(4 KiB row -> pageable ndarray)", lambda: store.get("x", hb, nxt()))
timeit("get(4 KiB row -> pinned ndarray)", lambda: store.get("x", pinn, nxt()))
timeit("get(4 KiB row -> CUDA tensor)", lambda: store.get("x", db, nxt()))
timeit("get(4 B label -> pageable ndarray)", lambda: store.get("lab", hl, nxt()))
for B in (32, 256):
    ob = torch.zeros((B, 1024), dtype=torch.float32, device="cuda")
    ids = idx[:B].copy()
    timeit(f"get_batch(B={B} host ids -> CUDA tensor), sync", lambda: store.get_batch("x", ids, out=ob, count=1))
    dids = torch.from_numpy(ids).cuda()
    timeit(f"get_batch(B={B} device ids -> CUDA tensor), sync", lambda: store.get_batch("x", dids, out=ob, count=1))
    oh = torch.zeros((B, 1024), dtype=torch.float32).pin_memory().numpy()
    timeit(f"get_batch(B={B} host ids -> pinned host), sync", lambda: store.get_batch("x", ids, out=oh, count=1))
store.free(); store.close()
