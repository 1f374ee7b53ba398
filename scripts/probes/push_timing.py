"""probe: per-CTA timeline of one collective push-fetch step (torchrun, 2+ ranks, DDS_DEBUG_TIMING=1)"""
import ctypes, os, sys
import numpy as np
os.environ.setdefault("DDS_DEBUG_TIMING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from ddstore_b200 import PyDDStore, TorchDistComm
L = ctypes.CDLL(os.path.join(ROOT, "ddstore_b200", "libddstore_b200.so"))
L.ddsk_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
rank, local, N = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("gloo", init_method="env://")
store = PyDDStore(TorchDistComm(), device=local)
rows, B = 2_000_000, 65536
store.init("x", rows, 1024, 4); store.synth_fill("x", 0xDD5)
store.push_setup(B, B * 4096)
rng = np.random.default_rng(1234 + rank)
idx = [torch.from_numpy(rng.integers(0, rows * N, size=B)).to(dev) for _ in range(4)]
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side); st = side.cuda_stream
for mode in ("push", "pull"):
    out = torch.empty(B * 4096, dtype=torch.uint8, device=dev)
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            if mode == "push":
                store.get_batch_push("x", idx[i % 4], stream=st)
            else:
                store.get_batch("x", idx[i % 4], out=out, count=1, stream=st, wait=False)
        e1.record(); store.wait(); torch.cuda.synchronize()
    REG = 4096 + 8
    buf = (ctypes.c_ulonglong * (2 * REG))()
    L.ddsk_debug_timing(buf, 2 * REG)
    a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64).reshape(2, REG)[0, :4096].reshape(1024, 4)
    a = a[a[:, 0] > 0]; t0 = a[:, 0].min(); r = (a - t0) / 1e3
    q = lambda v: "min %6.1f med %6.1f max %6.1f" % (v.min(), np.median(v), v.max())
    if rank == 0:
        print(f"== {mode}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per step; last launch of rank 0, us since its first CTA entered:")
        print("   entry          ", q(r[:, 0]))
        print("   prologue done  ", q(r[:, 1]), "(push: ready of every rank seen)")
        print("   first data     ", q(r[:, 2]))
        print("   last warp done ", q(r[:, 3]), flush=True)
dist.barrier()
store.free(); store.close(); dist.destroy_process_group()
