#!/usr/bin/env python
"""scripts/bench_configs.py -- throughput of the other BASELINE.json configs (3: variable length, 4: multi-array,
5: size sweep, 1: demo shape) through the same C-ABI call as bench.py, on 1 GPU or under torchrun on N GPUs.
Indices and output are device resident; K async launches between CUDA events; max over ranks.
Writes one JSON line per case to stdout (rank 0)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cases", default="demo,cfg3,cfg4,cfg5")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the stores (1.0 = sizes in the table below)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ddstore_b200 import PyDDStore, SelfComm, TorchDistComm

    N = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if N > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
        comm = TorchDistComm()
    else:
        comm = SelfComm()
    store = PyDDStore(comm, device=local)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    K, W = args.steps, args.warmup
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nbytes, label, extra):
        for _ in range(W):
            fn()
        store.wait()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        store.wait()
        barrier()
        ms = e0.elapsed_time(e1) / K
        if N > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        gbs = nbytes * N / ms / 1e6
        if rank == 0:
            rec = {"case": label, "n_gpus": N, "payload_GBps": round(gbs, 1), "ms_per_step": round(ms, 4),
                   "bytes_per_step_per_gpu": int(nbytes)}
            if N == 1:
                rec["hbm_frac_of_measured_copy_peak"] = round(2 * gbs / peaks["hbm_gbs"], 3)
            rec.update(extra)
            print(json.dumps(rec), flush=True)

    def share(total):
        per = total // N
        return per if rank < N - 1 else total - per * (N - 1)

    rng = np.random.default_rng(1234 + rank)
    cases = args.cases.split(",")

    # ---- config 1 shape (test/demo.py): rows of 64 float64 = 512 B, single-row requests
    if "demo" in cases:
        total = int(8 * 1024 * 1024 * args.scale)
        store.init("demo", share(total), 64, 8)
        store.synth_fill("demo", 0xDD5)
        for B in (4096, 262144):
            idx = torch.from_numpy(rng.integers(0, total, size=B)).to(dev)
            out = torch.empty(B * 512, dtype=torch.uint8, device=dev)
            timed(lambda: store.get_batch("demo", idx, out=out, count=1, stream=side.cuda_stream, wait=False), B * 512,
                  f"demo-shape 512 B rows, B={B}", {"requests_per_s": None})
        store.free()
        store = PyDDStore(comm, device=local)

    # ---- config 3: variable-length float32 samples, 100..10000 elements, disp=1
    if "cfg3" in cases:
        nsamp = int(1_000_000 * args.scale) * N
        L = np.random.default_rng(42).integers(100, 10001, size=nsamp)
        sstart = np.concatenate([[0], np.cumsum(L)])
        per = nsamp // N
        lo, hi = rank * per, (rank + 1) * per if rank < N - 1 else nsamp
        store.init("x", int(sstart[hi] - sstart[lo]), 1, 4)
        store.synth_fill("x", 0xDD5)
        d_start, d_len = torch.from_numpy(sstart[:-1].copy()).to(dev), torch.from_numpy(L).to(dev)
        for B in (4096, 16384):
            ids = torch.from_numpy(rng.integers(0, nsamp, size=B)).to(dev)
            st, ct = d_start[ids].contiguous(), d_len[ids].contiguous()
            nbytes = int(ct.sum().item()) * 4
            out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            offs = torch.empty(B + 1, dtype=torch.int64, device=dev)
            timed(lambda: store.get_batch("x", st, ct, out=out, offsets=offs, stream=side.cuda_stream, wait=False), nbytes,
                  f"cfg3 variable 100-10000 f32 (4-byte aligned rows), B={B}", {"samples": nsamp})
            out_c, offs_c = torch.empty_like(out), torch.empty_like(offs)
            flipc = [0]

            def queued_explicit():
                flipc[0] ^= 1
                store.get_batch("x", st, ct, out=out_c if flipc[0] else out, offsets=offs_c if flipc[0] else offs,
                                stream=side.cuda_stream, wait=False, overlap=True)

            timed(queued_explicit, nbytes, f"cfg3 explicit, DDS_OVERLAP double-buffered queue, B={B}", {"samples": nsamp})
            if B == 4096:
                store.set_sample_index("x", d_start, d_len)
            timed(lambda: store.get_samples("x", ids, out, offsets=offs, stream=side.cuda_stream, wait=False), nbytes,
                  f"cfg3 by sample id (device index), B={B}", {"samples": nsamp})
            out_b, offs_b = torch.empty_like(out), torch.empty_like(offs)
            flip = [0]

            def queued():
                flip[0] ^= 1
                store.get_samples("x", ids, out_b if flip[0] else out, offsets=offs_b if flip[0] else offs,
                                  stream=side.cuda_stream, wait=False, overlap=True)

            timed(queued, nbytes, f"cfg3 by sample id, DDS_OVERLAP double-buffered queue, B={B}", {"samples": nsamp})
        store.free()
        store = PyDDStore(comm, device=local)

    # ---- config 4: node_feat f32 [n,16] + edge_index i64 [8n,2]
    if "cfg4" in cases:
        nsamp = int(250_000 * args.scale) * N
        n = np.random.default_rng(43).integers(8, 513, size=nsamp)
        e = 8 * n
        ns, es = np.concatenate([[0], np.cumsum(n)]), np.concatenate([[0], np.cumsum(e)])
        per = nsamp // N
        lo, hi = rank * per, (rank + 1) * per if rank < N - 1 else nsamp
        store.init("node_feat", int(ns[hi] - ns[lo]), 16, 4)
        store.init("edge_index", int(es[hi] - es[lo]), 2, 8)
        store.synth_fill("node_feat", 0xDD5)
        store.synth_fill("edge_index", 0xDD6)
        dns, dn = torch.from_numpy(ns[:-1].copy()).to(dev), torch.from_numpy(n).to(dev)
        des, de = torch.from_numpy(es[:-1].copy()).to(dev), torch.from_numpy(e).to(dev)
        for B in (4096,):
            ids = torch.from_numpy(rng.integers(0, nsamp, size=B)).to(dev)
            s1, c1, s2, c2 = dns[ids].contiguous(), dn[ids].contiguous(), des[ids].contiguous(), de[ids].contiguous()
            b1, b2 = int(c1.sum().item()) * 64, int(c2.sum().item()) * 16
            o1 = torch.empty(b1, dtype=torch.uint8, device=dev)
            o2 = torch.empty(b2, dtype=torch.uint8, device=dev)

            def both():
                store.get_batch("node_feat", s1, c1, out=o1, stream=side.cuda_stream, wait=False)
                store.get_batch("edge_index", s2, c2, out=o2, stream=side.cuda_stream, wait=False)

            timed(both, b1 + b2, f"cfg4 node_feat f32[n,16] + edge_index i64[8n,2], B={B} (2 launches)", {"samples": nsamp})
            store.set_sample_index("node_feat", dns, dn)
            store.set_sample_index("edge_index", des, de)
            timed(lambda: store.get_samples_multi(["node_feat", "edge_index"], ids, [o1, o2], stream=side.cuda_stream,
                                                  wait=False),
                  b1 + b2, f"cfg4 both arrays by sample id in ONE launch, B={B}", {"samples": nsamp})
            p1, p2 = torch.empty_like(o1), torch.empty_like(o2)
            flip4 = [0]

            def queued4():
                flip4[0] ^= 1
                store.get_samples_multi(["node_feat", "edge_index"], ids, [p1, p2] if flip4[0] else [o1, o2],
                                        stream=side.cuda_stream, wait=False, overlap=True)

            timed(queued4, b1 + b2, f"cfg4 ONE launch, DDS_OVERLAP double-buffered queue, B={B}", {"samples": nsamp})
        store.free()
        store = PyDDStore(comm, device=local)

    # ---- config 5: size sweep, fixed-length float32 rows of R bytes, one row per request
    if "cfg5" in cases:
        for R in (1 << 10, 4 << 10, 16 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20):
            shard = max(1 << 30, 64 * R)
            rows = shard // R
            name = f"s{R}"
            store.init(name, rows, R // 4, 4)
            store.synth_fill(name, 0xDD5)
            B = max(1, (256 << 20) // R)
            idx = torch.from_numpy(rng.integers(0, rows * N, size=B)).to(dev)
            out = torch.empty(B * R, dtype=torch.uint8, device=dev)
            timed(lambda: store.get_batch(name, idx, out=out, count=1, stream=side.cuda_stream, wait=False), B * R,
                  f"cfg5 R={R} B, B={B}", {})
            store.free()
            store = PyDDStore(comm, device=local)
    store.close()
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
