"""scripts/probes/mix_probe.py -- measurement probe: gather throughput at 2 GPUs as a function of the remote
fraction of the batch, with thread-ranks (same-process peer access) vs process-ranks (CUDA IPC), one rank or both
ranks fetching. Usage: python mix_probe.py threads|procs"""
import os
import sys
import threading
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

PER, DISP, B = 2_000_000, 1024, 65536


def body(rank, P, key, barrier, results, geom_env=None):
    import torch
    from ddstore_b200 import PyDDStore, ShmComm
    torch.cuda.set_device(rank)
    comm = ShmComm(key, rank, P)
    store = PyDDStore(comm, device=rank)
    store.init("x", PER, DISP, 4)
    store.synth_fill("x", 0xDD5)
    dev = torch.device("cuda", rank)
    out = torch.empty(B * DISP * 4, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)
    rng = np.random.default_rng(5 + rank)
    for both in (False, True):
        for frac in (0.0, 0.5, 1.0):
            nrem = int(B * frac)
            other = 1 - rank
            idx = np.concatenate([rng.integers(other * PER, (other + 1) * PER, size=nrem),
                                  rng.integers(rank * PER, (rank + 1) * PER, size=B - nrem)])
            rng.shuffle(idx)
            d_idx = torch.from_numpy(idx).to(dev)
            torch.cuda.synchronize()
            comm.Barrier()
            active = both or rank == 0
            ms = 0.0
            if active:
                with torch.cuda.stream(side):
                    for _ in range(3):
                        store.get_batch("x", d_idx, out=out, count=1, stream=side.cuda_stream, wait=False)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        store.get_batch("x", d_idx, out=out, count=1, stream=side.cuda_stream, wait=False)
                    e1.record()
                    store.wait()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 10
            comm.Barrier()
            if active:
                print(f"rank {rank} both={both} remote_frac={frac:4.2f}: {B * DISP * 4 / ms / 1e6:8.1f} GB/s payload "
                      f"({B * DISP * 4 * frac / ms / 1e6:7.1f} over NVLink)", flush=True)
    store.free()
    store.close()
    comm.close()


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "threads"
    key = "mx" + uuid.uuid4().hex[:10]
    if mode == "threads":
        th = [threading.Thread(target=body, args=(r, 2, key, None, None)) for r in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
    elif mode == "procs":
        import subprocess
        ps = [subprocess.Popen([sys.executable, __file__, "child", str(r), key]) for r in range(2)]
        [p.wait() for p in ps]
    else:
        body(int(sys.argv[2]), 2, sys.argv[3], None, None)


if __name__ == "__main__":
    main()
