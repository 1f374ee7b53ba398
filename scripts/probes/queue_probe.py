"""probe: timeline of two consecutive batches of an overlapped variable-count queue (DDS_DEBUG_TIMING=1)"""
import ctypes, os, sys
import numpy as np
os.environ.setdefault("DDS_DEBUG_TIMING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore
L = ctypes.CDLL(os.path.join(ROOT, "ddstore_b200", "libddstore_b200.so"))
L.ddsk_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
store = PyDDStore(device=0)
nsamp = 200_000
Ls = np.random.default_rng(42).integers(100, 10001, size=nsamp)
sstart = np.concatenate([[0], np.cumsum(Ls)])
store.init("x", int(sstart[-1]), 1, 4); store.synth_fill("x", 1)
d_start, d_len = torch.from_numpy(sstart[:-1].copy()).to(dev), torch.from_numpy(Ls).to(dev)
store.set_sample_index("x", d_start, d_len)
rng = np.random.default_rng(0)
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side); st = side.cuda_stream
REG = 4096 + 8
for B in (4096, 16384):
    ids = torch.from_numpy(rng.integers(0, nsamp, size=B)).to(dev)
    s_, c_ = d_start[ids].contiguous(), d_len[ids].contiguous()
    nbytes = int(c_.sum().item()) * 4
    outs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    offs = [torch.empty(B + 1, dtype=torch.int64, device=dev) for _ in range(2)]
    for mode in ("explicit", "by-sample"):
        for rep in range(2):
            for i in range(9):
                if mode == "explicit":
                    store.get_batch("x", s_, c_, out=outs[i & 1], offsets=offs[i & 1], stream=st, wait=False, overlap=True)
                else:
                    store.get_samples("x", ids, outs[i & 1], offsets=offs[i & 1], stream=st, wait=False, overlap=True)
            store.wait(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (2 * REG))()
        L.ddsk_debug_timing(buf, 2 * REG)
        a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64).reshape(2, REG)
        g = [a[r, :4096].reshape(1024, 4) for r in range(2)]
        g = [x[x[:, 0] > 0] for x in g]
        last = 0 if g[0][:, 3].max() > g[1][:, 3].max() else 1   # region of the LAST launch (q); the other is q-1
        prev = 1 - last
        t0 = g[prev][:, 0].min()
        f = lambda v: "%7.1f .. %7.1f (med %7.1f)" % ((v.min() - t0) / 1e3, (v.max() - t0) / 1e3, (np.median(v) - t0) / 1e3)
        print(f"== {mode} B={B} ({nbytes/1e6:.0f} MB, floor {2*nbytes/6571e3:.1f} us); times in us since the first CTA of gather q-1 started")
        print("   gather q-1: CTA entry   ", f(g[prev][:, 0]))
        print("   gather q-1: last warp   ", f(g[prev][:, 3]))
        p = a[last, 4096:4100]
        print("   plan q: lookup last CTA start %.1f end %.1f | scan last CTA start %.1f end %.1f" % tuple((p - t0) / 1e3))
        print("   gather q  : CTA entry   ", f(g[last][:, 0]))
        print("   gather q  : plan known  ", f(g[last][:, 1]))
        print("   gather q  : first data  ", f(g[last][:, 2]))
        print("   gather q  : last warp   ", f(g[last][:, 3]), flush=True)
store.free(); store.close()
