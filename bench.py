#!/usr/bin/env python
"""bench.py -- batch-fetch throughput of the get() hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU MPI_Get path on host cores

Workload (BASELINE.json configs[1]): 10,000,000 samples x 1024 float32 (4096 B rows, 40.96 GB) sharded by
contiguous blocks over the N GPUs; one "step" = every rank fetches one batch of B=65536 uniform-random
samples (268 MB) into a packed device buffer. Payload is synthetic (splitmix64 of the element index,
generated on device); indices are seeded per rank and per step.

Printed JSON (one line, rank 0):
  value      aggregate GB/s over all ranks, indices and output resident in HBM, K back-to-back steps timed with
             CUDA events between a barrier+synchronize on both sides, max over ranks
  e2e        same metric through the host-facing call: pinned HOST index arrays in, pinned HOST buffer out,
             H2D + kernel + D2H all inside the timed region
  roofline   dominant kernel (dds_gather_kernel): algorithmic bytes = 2 x payload (one HBM read + one HBM write
             per byte at N=1), per-launch duration from CUDA events on the launching stream, peak from
             MEASURED_PEAKS.json
  cpu_baseline  oracle/_ref (the unmodified reference compiled against the MPI thread-rank shim) doing the same
             per-sample get() loop on the host cores, on a bounded sample of the workload
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOTAL_SAMPLES = 10_000_000
DISP = 1024
ROW_BYTES = DISP * 4
SEED = 0xDD5
METRIC = "batch_fetch_GBps"
UNIT = "GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per rank per step")
    ap.add_argument("--samples", type=int, default=TOTAL_SAMPLES, help="total samples (default: the full config)")
    ap.add_argument("--cpu-samples", type=int, default=1_000_000, help="rows in the CPU baseline's bounded sample")
    ap.add_argument("--cpu-batch", type=int, default=32768, help="get() calls per rank-thread per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def np_synth_rows(seed, first_global_row, nrows, disp, dtype):
    """host recomputation of the device payload generator: low bytes of splitmix64(seed ^ index)"""
    dtype = np.dtype(dtype)
    g = (np.arange(nrows, dtype=np.uint64)[:, None] + np.uint64(first_global_row)) * np.uint64(disp) \
        + np.arange(disp, dtype=np.uint64)[None, :]
    x = (g ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    raw = x.view(np.uint8).reshape(nrows, disp, 8)[:, :, :dtype.itemsize]
    return np.ascontiguousarray(raw).view(dtype).reshape(nrows, disp)


def bind_to_gpu_numa(gpu_index):
    """pin this rank to the CPUs next to its GPU (NVML's ideal affinity) so pinned host buffers and the PCIe
    copies of the e2e path stay on the GPU's NUMA node"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:  # noqa: BLE001 -- best effort
        pass


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def cpu_reference_run(cpu_samples, cpu_batch, steps, warmup, nthreads=None):
    """The reference's own get() path on the host cores: the UNMODIFIED DDStore (method 0) compiled against the
    MPI thread-rank shim when oracle/_ref is built, else the oracle's C port. One rank-thread per core,
    each doing `cpu_batch` blocking single-row get() calls per step into a packed host buffer -- the loader loop
    of examples/vae/distdataset.py:79-89. Returns (GB/s aggregate, info dict)."""
    from oracle import oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if nthreads is None:
        # "all the host threads it can use": more rank-threads than the memory system can feed only adds contention,
        # and threads spread over both sockets pay for remote memory, so time a short run for every candidate
        # (cpu set, thread count) and keep the fastest
        base = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        cpu_sets = [("all", base)]
        for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
            try:
                cpus = set()
                for part in open(os.path.join(nd, "cpulist")).read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus |= set(range(int(lo), int(hi or lo) + 1))
                if base is not None and cpus & base and (cpus & base) != base:
                    cpu_sets.append((os.path.basename(nd), cpus & base))
            except Exception:  # noqa: BLE001
                pass
        best = None
        for tag, cset in cpu_sets:
            if cset is not None:
                os.sched_setaffinity(0, cset)
            ncs = len(cset) if cset is not None else cores
            for cand in sorted({c for c in (8, 16, 32, 64, 128, ncs) if c <= ncs} | {min(ncs, 8)}):
                g, _ = cpu_reference_run(min(cpu_samples, 250_000), cpu_batch // 4, 2, 1, nthreads=cand)
                if best is None or g > best[0]:
                    best = (g, cand, cset, tag)
        nthreads = best[1]
        if best[2] is not None:
            os.sched_setaffinity(0, best[2])
        cpu_reference_run.last_cpu_set = best[3]
    P = max(1, min(256, nthreads))
    per = cpu_samples // P
    co = O.COracle()
    shards = [co.synth_rows(SEED, r * per, per, DISP, np.float32) for r in range(P)]
    total = per * P
    rngs = [np.random.default_rng(1234 + r) for r in range(P)]
    outs = [np.empty(cpu_batch * ROW_BYTES, np.uint8) for _ in range(P)]
    counts = [np.ones(cpu_batch, np.int64) for _ in range(P)]
    times = []
    if O.have_ref():
        kind = "reference"
        w = O.RefWorld(P)
        w.add("x", shards)
        for it in range(warmup + steps):
            starts = [rng.integers(0, total, size=cpu_batch) for rng in rngs]
            ns = w.get_loop_all("x", starts, counts, outs)
            if ns < 0:
                raise RuntimeError("reference get() failed: " + w.err())
            if it >= warmup:
                times.append(ns * 1e-9)
        # parity spot check of the timed path against the generator
        exp = co.synth_rows(SEED, int(starts[0][-1]), 1, DISP, np.float32)
        assert outs[0][-ROW_BYTES:].tobytes() == exp.tobytes()
        w.close()
    else:
        kind = "port"
        import ctypes as C
        bases = (C.c_void_p * P)(*[s.ctypes.data for s in shards])
        ll = O.np_lenlist([per] * P)
        LP = C.POINTER(C.c_long)

        def one(r, starts):
            co.L.orc_get_batch(bases, ll.ctypes.data_as(LP), P, DISP, 4, 4, starts.ctypes.data_as(LP),
                               counts[r].ctypes.data_as(LP), cpu_batch, outs[r].ctypes.data, None, None)

        for it in range(warmup + steps):
            starts = [np.ascontiguousarray(rng.integers(0, total, size=cpu_batch)) for rng in rngs]
            th = [threading.Thread(target=one, args=(r, starts[r])) for r in range(P)]
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    step_bytes = P * cpu_batch * ROW_BYTES
    t = float(np.sum(times))
    gbs = step_bytes * len(times) / t / 1e9
    info = {"value": gbs, "unit": UNIT, "cores": P, "kind": kind,
            "sample": f"{P} rank-threads (fastest of the calibrated thread counts / cpu sets) x {cpu_batch} single-row "
                      f"get() per step x {len(times)} steps on a {total}-row ({total * ROW_BYTES / 1e9:.2f} GB) slice "
                      f"of the workload, host buffers",
            "samples_per_s": P * cpu_batch * len(times) / t, "host_cpus": cores,
            "cpu_set": getattr(cpu_reference_run, "last_cpu_set", "inherited"),
            "ms_per_step": 1e3 * t / len(times)}
    return gbs, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    # keep the whole run within a few minutes whatever K is asked for
    gbs, info = cpu_reference_run(args.cpu_samples, args.cpu_batch, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: fixed-length 1024-float32 samples, uniform-random batch fetch "
                                   "(bounded CPU sample of the 10M-sample store)", "row_bytes": ROW_BYTES},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "samples_per_s": info["samples_per_s"],
            "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ddstore_b200 import PyDDStore, SelfComm, TorchDistComm, _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    if world != N:
        raise SystemExit(f"--gpus {N} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {N}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa(local)
    if N > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
        comm = TorchDistComm()
    else:
        comm = SelfComm()

    K, W, B = args.steps, max(args.warmup, 3), args.batch
    total = args.samples
    per = total // N
    nrows = per if rank < N - 1 else total - per * (N - 1)

    store = PyDDStore(comm, device=local)
    store.init("x", nrows, DISP, 4)
    store.synth_fill("x", SEED)
    lenlist = store.query("x")["lenlist"]
    assert lenlist[-1] == total

    rng = np.random.default_rng(1234 + rank)
    nsets = min(K + W, 16)
    idx_host = [torch.from_numpy(rng.integers(0, total, size=B)).pin_memory() for _ in range(nsets)]
    idx_dev = [t.to(dev) for t in idx_host]
    out_dev = torch.empty(B * ROW_BYTES, dtype=torch.uint8, device=dev)
    step_bytes = B * ROW_BYTES
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)  # the launching stream: kernels AND the timing events live on it
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness spot check of exactly the timed call (generator recomputed on the host)
    store.get_batch("x", idx_dev[0], out=out_dev, count=1, stream=stream)
    got = out_dev.view(torch.float32).view(B, DISP)
    for j in (0, B // 2, B - 1):
        exp = np_synth_rows(SEED, int(idx_host[0][j]), 1, DISP, np.float32)
        assert got[j].cpu().numpy().tobytes() == exp.tobytes(), "bench: fetched row differs from the generator"

    # ---- value: device-resident indices and output, K back-to-back async launches. The batches are independent
    # (static device-resident index sets, two alternating output buffers -- a double-buffered prefetch queue), so
    # they are queued with overlap=True: the head of batch k+1 fills the SMs the tail of batch k vacates.
    out_dev2 = torch.empty_like(out_dev)
    outs = (out_dev, out_dev2)
    for i in range(W):
        store.get_batch("x", idx_dev[i % nsets], out=outs[i & 1], count=1, stream=stream, wait=False, overlap=True)
    store.wait()
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _capi.lib().dds_kernel_launches()
    barrier()
    t_all0.record()
    for i in range(K):  # EXACTLY K steps, nothing else on the stream
        store.get_batch("x", idx_dev[(W + i) % nsets], out=outs[i & 1], count=1, stream=stream, wait=False, overlap=True)
    t_all1.record()
    store.wait()
    barrier()
    launches = _capi.lib().dds_kernel_launches() - launches0
    # the last two batches of the timed region, against the generator
    for i in (K - 2, K - 1):
        if i < 0:
            continue
        g2 = outs[i & 1].view(torch.float32).view(B, DISP)
        ih = idx_host[(W + i) % nsets]
        for j in (0, B // 3, B - 1):
            exp = np_synth_rows(SEED, int(ih[j]), 1, DISP, np.float32)
            assert g2[j].cpu().numpy().tobytes() == exp.tobytes(), "bench: timed batch differs from the generator"
    ms_total = t_all0.elapsed_time(t_all1)
    # second pass, same K steps, one CUDA-event pair around every launch: the kernel's own duration for the roofline
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        ev[i][0].record()
        store.get_batch("x", idx_dev[(W + i) % nsets], out=out_dev, count=1, stream=stream, wait=False)
        ev[i][1].record()
    store.wait()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    pair_list = [a.elapsed_time(b) for a, b in ev]
    per_launch_ms = float(np.mean(pair_list))
    pair_pcts = [float(np.percentile(pair_list, q)) for q in (10, 50, 90)]
    if N > 1:
        t = torch.tensor([ms_total, per_launch_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, per_launch_ms = float(t[0]), float(t[1])
    value = N * K * step_bytes / (ms_total * 1e-3) / 1e9

    # ---- e2e: pinned host indices in, pinned host buffer out, through the same call
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty(B * ROW_BYTES, dtype=torch.uint8).pin_memory()
        out_np = out_host.numpy()
        idx_np = [t.numpy() for t in idx_host]
        Ke = max(3, min(K, 10))
        for i in range(3):
            store.get_batch("x", idx_np[i % nsets], out=out_np, count=1)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            store.get_batch("x", idx_np[(W + i) % nsets], out=out_np, count=1)
        torch.cuda.synchronize()
        t_e = time.perf_counter() - t0
        exp = np_synth_rows(SEED, int(idx_np[(W + Ke - 1) % nsets][B - 1]), 1, DISP, np.float32)
        assert out_np[-ROW_BYTES:].tobytes() == exp.tobytes(), "bench e2e: last row differs from the generator"
        if N > 1:
            t = torch.tensor([t_e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e = float(t[0])
        e2e = {"value": N * Ke * step_bytes / t_e / 1e9, "unit": UNIT, "h2d_bytes_per_step": B * 8,
               "d2h_bytes_per_step": step_bytes, "steps": Ke, "ms_per_step": 1e3 * t_e / Ke,
               "path": "dds_get_batch(host int64 starts -> pinned host buffer): H2D idx + gather kernel + D2H payload"}

    # ---- roofline of the dominant kernel
    peaks, peak_src = measured_peaks()
    import ctypes
    geom = [ctypes.c_int() for _ in range(5)]
    _capi.lib().dds_gather_geometry(*[ctypes.byref(g) for g in geom])
    if N == 1:
        bound, alg_bytes, peak = "hbm", 2 * step_bytes, float(peaks["hbm_gbs"])
        note = ("algorithmic bytes per launch = 2 x payload (each byte read once from HBM, written once to HBM); "
                "per_launch_ms = timed region / K with the launches overlapping head-to-tail (DDS_OVERLAP), "
                "per_launch_event_pair_ms = the same kernel serialised, one CUDA-event pair per launch")
    else:
        # per GPU: payload r; HBM moves 2r; NVLink-in carries r(N-1)/N at <= 770 GB/s measured per direction
        bound, alg_bytes = "nvlink", step_bytes * (N - 1) / N
        peak = 770.0
        note = "algorithmic NVLink-in bytes per launch = payload x (N-1)/N (uniform-random owners); peak = measured 770 GB/s/dir"
    # The gather kernel is the only kernel of a step, so the timed region itself (K launches between two CUDA events
    # on the launching stream) gives its average launch duration; the event-pair pass is reported next to it (it
    # puts two event records between consecutive kernels, which defeats the programmatic-dependent-launch overlap).
    pair_ms = per_launch_ms
    per_launch_ms = ms_total / K
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    if N == 1:
        tp = os.path.join(ROOT, "profiles", "gather_fixed_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]  # per launch, from the committed ncu --set full capture
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "dds_gather_kernel<FIXED>", "per_launch_ms": per_launch_ms,
                "per_launch_event_pair_ms": pair_ms, "per_launch_event_pair_ms_p10_p50_p90": pair_pcts,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "note": note}

    if rank == 0:
        cpu = None
        if N == 1 and not args.no_cpu_baseline:
            try:
                _, info = cpu_reference_run(args.cpu_samples, args.cpu_batch, 8, 2)
                cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
                cpu["samples_per_s"] = info["samples_per_s"]
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "configs[1]: 10M fixed-length 1024-float32 samples, uniform-random batch fetch",
                           "total_samples": total, "row_bytes": ROW_BYTES, "batch_per_gpu": B,
                           "bytes_per_step_per_gpu": step_bytes, "store_bytes": total * ROW_BYTES,
                           "l2": "inputs larger than L2 (random rows of a %.1f GB shard per GPU; 268 MB output)"
                                 % (nrows * ROW_BYTES / 1e9),
                           "parallelism": f"store sharded over {N} GPU(s), VMM peer mappings, no collective",
                           "queue": "K independent batches queued asynchronously on one stream with DDS_OVERLAP into two "
                                    "alternating output buffers (double-buffered prefetch); the last two are checked "
                                    "against the generator after the timed region",
                           "gather_geometry": {"ctas": geom[0].value, "warps_per_cta": geom[1].value,
                                               "stages": geom[2].value, "chunk_bytes": geom[3].value,
                                               "smem_bytes": geom[4].value}},
                "samples_per_s": N * K * B / (ms_total * 1e-3), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    barrier()
    store.free()
    store.close()
    if N > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if not os.path.exists(os.path.join(ROOT, "ddstore_b200", "libddstore_b200.so")):
        import __graft_entry__  # fresh checkout: build the native pieces in-tree first
        __graft_entry__.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
