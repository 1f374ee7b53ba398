"""probe: NVLink counters of the fixed-stride gather at N=2 (run under torchrun; rank 0 is wrapped by ncu, see
scripts/r2/nvlink_ncu_wrap.sh). Rank 0 launches M batches of config 2 (B=65536 x 4 KiB); rank 1 either idles (mode B:
one requester) or keeps fetching its own batches until rank 0 is done (mode A: both directions busy)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from ddstore_b200 import PyDDStore, TorchDistComm
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
mode = sys.argv[1] if len(sys.argv) > 1 else "A"
flag = sys.argv[2] if len(sys.argv) > 2 else "/tmp/nvlink_ncu_done"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo", init_method="env://")   # host-side only: no NCCL kernels under the profiler
store = PyDDStore(TorchDistComm(), device=local)
N, B, rows = 2, 65536, 2_000_000
store.init("x", rows, 1024, 4); store.synth_fill("x", 0xDD5)
rng = np.random.default_rng(1234 + rank)
remote_only = mode.endswith("r")
lo, hi = (0, rows * N) if not remote_only else ((1 - rank) * rows, (2 - rank) * rows)
idx = [torch.from_numpy(rng.integers(lo, hi, size=B)).to(dev) for _ in range(4)]
outs = [torch.empty(B * 4096, dtype=torch.uint8, device=dev) for _ in range(2)]
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side); st = side.cuda_stream
for i in range(3):
    store.get_batch("x", idx[i % 4], out=outs[i & 1], count=1, stream=st, wait=False)
store.wait(); dist.barrier()
if rank == 0:
    for i in range(12):
        store.get_batch("x", idx[i % 4], out=outs[i & 1], count=1, stream=st, wait=False)
    store.wait()
    open(flag, "w").write("done")
else:
    n = 0
    while not os.path.exists(flag):
        if mode.startswith("A"):
            for i in range(8):
                store.get_batch("x", idx[i % 4], out=outs[i & 1], count=1, stream=st, wait=False, overlap=True)
            store.wait(); n += 8
        else:
            time.sleep(0.01)
    print(f"rank 1: {n} concurrent batches while rank 0 was profiled", flush=True)
dist.barrier()
store.free(); store.close()
dist.destroy_process_group()
