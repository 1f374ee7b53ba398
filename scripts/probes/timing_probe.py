"""probe: where the time of one gather launch goes (DDS_DEBUG_TIMING=1: per-CTA globaltimer stamps)"""
import ctypes, os, sys, time
import numpy as np
os.environ.setdefault("DDS_DEBUG_TIMING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ddstore_b200 import PyDDStore, _capi
L = ctypes.CDLL(os.path.join(ROOT, "ddstore_b200", "libddstore_b200.so"))
L.ddsk_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
store = PyDDStore(device=0)
nsamp = 200_000
Ls = np.random.default_rng(42).integers(100, 10001, size=nsamp)
sstart = np.concatenate([[0], np.cumsum(Ls)])
store.init("x", int(sstart[-1]), 1, 4); store.synth_fill("x", 1)
store.init("f", 1_000_000, 1024, 4); store.synth_fill("f", 2)
d_start, d_len = torch.from_numpy(sstart[:-1].copy()).to(dev), torch.from_numpy(Ls).to(dev)
store.set_sample_index("x", d_start, d_len)
rng = np.random.default_rng(0)
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side); st = side.cuda_stream

def report(label, nbytes):
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (1024 * 4))()
    n = L.ddsk_debug_timing(buf, 1024)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 4)[:n].astype(np.int64)
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    r = (a - t0) / 1e3
    tot = r[:, 3].max()
    q = lambda v: "min %6.1f med %6.1f max %6.1f" % (v.min(), np.median(v), v.max())
    print(f"{label}: {len(a)} CTAs, kernel span {tot:.1f} us ({nbytes/1e6:.0f} MB -> {2*nbytes/tot/1e3:.0f} GB/s traffic, floor {2*nbytes/6571e3:.1f} us)")
    print("   entry     ", q(r[:, 0]))
    print("   plan done ", q(r[:, 1]), " (+%.1f med after entry)" % np.median(r[:, 1] - r[:, 0]))
    print("   first data", q(r[:, 2]), " (+%.1f med after plan)" % np.median(r[:, 2] - r[:, 1]))
    print("   last warp ", q(r[:, 3]), flush=True)

def timed(fn, k=20):
    for _ in range(3): fn()
    store.wait(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): fn()
    e1.record(); tc = time.perf_counter() - t0
    store.wait(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3, tc / k * 1e6

for B in (4096, 16384):
    ids = torch.from_numpy(rng.integers(0, nsamp, size=B)).to(dev)
    s_, c_ = d_start[ids].contiguous(), d_len[ids].contiguous()
    nbytes = int(c_.sum().item()) * 4
    out = torch.empty(nbytes, dtype=torch.uint8, device=dev); out2 = torch.empty_like(out)
    offs = torch.empty(B + 1, dtype=torch.int64, device=dev)
    gpu_us, cpu_us = timed(lambda: store.get_batch("x", s_, c_, out=out, offsets=offs, stream=st, wait=False))
    print(f"== cfg3 explicit B={B}: {gpu_us:.1f} us/step on the GPU, {cpu_us:.1f} us/step of CPU enqueue")
    store.get_batch("x", s_, c_, out=out, offsets=offs, stream=st); report(f"cfg3 explicit B={B}", nbytes)
    gpu_us, cpu_us = timed(lambda: store.get_samples("x", ids, out, offsets=offs, stream=st, wait=False))
    print(f"== cfg3 by-sample B={B}: {gpu_us:.1f} us/step on the GPU, {cpu_us:.1f} us/step of CPU enqueue")
    store.get_samples("x", ids, out, offsets=offs, stream=st); report(f"cfg3 by-sample B={B}", nbytes)
    # fixed-count batch of the same number of bytes (4 KiB rows)
    Bf = nbytes // 4096
    idf = torch.from_numpy(rng.integers(0, 1_000_000, size=Bf)).to(dev)
    of = torch.empty(Bf * 4096, dtype=torch.uint8, device=dev)
    gpu_us, cpu_us = timed(lambda: store.get_batch("f", idf, out=of, count=1, stream=st, wait=False))
    print(f"== fixed 4 KiB rows, same bytes (B={Bf}): {gpu_us:.1f} us/step on the GPU, {cpu_us:.1f} us/step of CPU enqueue")
    store.get_batch("f", idf, out=of, count=1, stream=st); report(f"fixed B={Bf}", Bf * 4096)
store.free(); store.close()
