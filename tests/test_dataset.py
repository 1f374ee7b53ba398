"""Loader adapter: DistDataset / RaggedDataset over the batched fetch, against the source
arrays. The DataLoader path goes through __getitems__ (one launch per variable per batch)."""
import threading
import uuid

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _world(P, body):
    from ddstore_b200 import ShmComm
    key = "ds" + uuid.uuid4().hex[:10]
    errs, res = [], [None] * P

    def run(r):
        try:
            comm = ShmComm(key, r, P)
            res[r] = body(comm, r)
            comm.close()
        except BaseException:  # noqa: BLE001
            import traceback
            errs.append(traceback.format_exc())

    th = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert not errs, "\n".join(errs)
    return res


def test_distdataset_per_sample_and_batched_loader():
    import torch
    from ddstore_b200.dataset import DistDataset, make_loader
    rng = np.random.default_rng(3)
    N, P = 1000, 2
    images = rng.integers(0, 2**32, size=(N, 1, 8, 8), dtype=np.uint32).view(np.float32)
    labels = rng.integers(0, 10, size=N)
    data = [(images[i], int(labels[i])) for i in range(N)]

    def body(comm, r):
        ds = DistDataset(data, "train", comm=comm, device=0)
        assert len(ds) == N
        for idx in (0, 499, 500, 999, int(rng.integers(N))):
            val, lab = ds[idx]  # the reference's per-sample path (distdataset.py:79-92): a CPU tensor and an int
            assert not val.is_cuda and val.shape == (1, 8, 8) and val.numpy().tobytes() == images[idx].tobytes() and lab == labels[idx]
        seen = []
        loader = make_loader(ds, batch_size=64, rank=r, world_size=P, shuffle=True, seed=7)
        for epoch in range(2):
            loader.sampler.set_epoch(epoch)
            order = []
            for vals, labs in loader:
                ds_idx = None
                assert vals.is_cuda and vals.shape[1:] == (1, 8, 8)
                order.append((vals.cpu().numpy(), labs.cpu().numpy()))
            got = np.concatenate([o[0] for o in order])
            gl = np.concatenate([o[1] for o in order])
            idx = np.array(list(iter(loader.sampler)))  # same epoch -> same permutation
            assert got.tobytes() == images[idx].tobytes() and np.array_equal(gl, labels[idx])
            seen.append(idx)
        assert not np.array_equal(seen[0], seen[1])  # epoch shuffle
        ds.free()
        return set(seen[0].tolist())

    parts = _world(P, body)
    assert parts[0] | parts[1] == set(range(N))


def test_ddstore_width_replica_groups():
    """ddstore_width < comm size: the communicator is split into replica groups of that width exactly as the reference
    does (examples/vae/distdataset.py:25-30); every group holds the WHOLE dataset sharded over its own members, the
    groups' stores are disjoint, and any rank can fetch any sample from its own group"""
    from ddstore_b200.dataset import DistDataset
    rng = np.random.default_rng(11)
    N, P, width = 600, 4, 2
    images = rng.integers(0, 2**32, size=(N, 6), dtype=np.uint32).view(np.float32)
    labels = rng.integers(0, 10, size=N)
    data = [(images[i], int(labels[i])) for i in range(N)]

    def body(comm, r):
        ds = DistDataset(data, "w", comm=comm, ddstore_width=width, device=0)
        assert ds.ddstore_comm_size == width and ds.ddstore_comm_rank == r % width and ds.ddstore.size == width
        info = ds.ddstore.query("wdata")
        assert info["nranks"] == width and info["lenlist"] == [N // 2, N] and len(ds) == N
        ids = rng.integers(0, N, size=97)
        vals, labs = ds.__getitems__(ids.tolist())
        assert vals.cpu().numpy().tobytes() == images[ids].tobytes() and np.array_equal(labs.cpu().numpy(), labels[ids])
        v, lab = ds[int(ids[0])]
        assert v.numpy().tobytes() == images[ids[0]].tobytes() and lab == labels[ids[0]]
        base = info["local_base"]
        ds.free()
        ds.ddstore.close()
        ds.ddstore_comm.close()
        return base

    bases = _world(P, body)
    assert len(set(bases)) == P  # four distinct shards: two per replica group


def test_ragged_dataset_config4_shape():
    import torch
    from ddstore_b200.dataset import RaggedDataset
    rng = np.random.default_rng(8)
    P, per = 2, 300
    world = []
    for r in range(P):
        n = rng.integers(8, 200, size=per)
        e = 8 * n
        feat = rng.integers(0, 2**32, size=(int(n.sum()), 16), dtype=np.uint32).view(np.float32)
        edge = rng.integers(-2**40, 2**40, size=(int(e.sum()), 2), dtype=np.int64)
        world.append((n, e, feat, edge))

    def sample(i):
        r, j = divmod(i, per)
        n, e, feat, edge = world[r]
        ns, es = np.concatenate([[0], np.cumsum(n)]), np.concatenate([[0], np.cumsum(e)])
        return feat[ns[j]:ns[j + 1]], edge[es[j]:es[j + 1]]

    def body(comm, r):
        n, e, feat, edge = world[r]
        ds = RaggedDataset({"node_feat": feat, "edge_index": edge}, {"node_feat": n, "edge_index": e}, comm=comm, device=0)
        assert len(ds) == P * per
        ids = rng.integers(0, P * per, size=97).tolist()
        batch = ds.__getitems__(ids)
        f, fo = batch["node_feat"]
        ed, eo = batch["edge_index"]
        fo, eo = fo.cpu().numpy(), eo.cpu().numpy()
        for k, i in enumerate(ids):
            sf, se = sample(i)
            assert f[fo[k]:fo[k + 1]].cpu().numpy().tobytes() == sf.tobytes()
            assert ed[eo[k]:eo[k + 1]].cpu().numpy().tobytes() == se.tobytes()
        # the same dataset through the double-buffered prefetch loader (overlapped multi-array fetches on a side stream,
        # with consumer work queued between the batches), a whole shuffled epoch, ragged last batch included
        from ddstore_b200.dataset import RaggedPrefetchLoader
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(ds, num_replicas=P, rank=r, shuffle=True, seed=3)
        sampler.set_epoch(1)
        order = list(iter(sampler))
        seen, acc = 0, torch.zeros((), device="cuda:0")
        for batch in RaggedPrefetchLoader(ds, sampler, batch_size=64):
            f, fo = batch["node_feat"]
            ed, eo = batch["edge_index"]
            acc = acc + torch.nan_to_num(f).abs().sum() + ed.sum()   # consumer work on the current stream
            fo, eo = fo.cpu().numpy(), eo.cpu().numpy()
            nb = len(fo) - 1
            for k in range(0, nb, 7):
                sf, se = sample(order[seen + k])
                assert f[fo[k]:fo[k + 1]].cpu().numpy().tobytes() == sf.tobytes()
                assert ed[eo[k]:eo[k + 1]].cpu().numpy().tobytes() == se.tobytes()
            assert fo[-1] == f.shape[0] and eo[-1] == ed.shape[0]
            seen += nb
        assert seen == len(order)
        ds.free()
        return True

    assert all(_world(P, body))


def test_prefetch_loader_matches_plain_loader():
    import torch
    from ddstore_b200.dataset import DistDataset, PrefetchLoader
    from torch.utils.data.distributed import DistributedSampler
    rng = np.random.default_rng(5)
    N, P = 700, 2
    images = rng.integers(0, 2**32, size=(N, 3, 4), dtype=np.uint32).view(np.float32)
    labels = rng.integers(0, 10, size=N)
    data = [(images[i], int(labels[i])) for i in range(N)]

    def body(comm, r):
        ds = DistDataset(data, "t", comm=comm, device=0)
        sampler = DistributedSampler(ds, num_replicas=P, rank=r, shuffle=True, seed=1)
        sampler.set_epoch(3)
        order = np.array(list(iter(sampler)))
        got_v, got_l = [], []
        acc = torch.zeros((), device="cuda:0")
        for vals, labs in PrefetchLoader(ds, sampler, batch_size=48):
            acc = acc + torch.nan_to_num(vals).abs().sum()  # consumer work queued on the current stream
            got_v.append(vals.clone())
            got_l.append(labs.clone())
        v = torch.cat(got_v).cpu().numpy()
        lab = torch.cat(got_l).cpu().numpy()
        assert v.tobytes() == images[order].tobytes() and np.array_equal(lab, labels[order])
        # the same epoch through a device-resident sampler: no per-batch index copy at all
        from ddstore_b200.dataset import DeviceBatchSampler
        dsamp = DeviceBatchSampler(len(ds), 48, rank=r, world_size=P, seed=1, device="cuda:0")
        dsamp.set_epoch(3)
        got2 = torch.cat([vals.clone() for vals, _ in PrefetchLoader(ds, dsamp, batch_size=48)]).cpu().numpy()
        assert got2.tobytes() == images[order].tobytes()
        # small batches grouped: 5 batches per launch, same batches out (host-side and device-resident samplers)
        dsamp16 = DeviceBatchSampler(len(ds), 16, rank=r, world_size=P, seed=1, device="cuda:0")
        dsamp16.set_epoch(3)
        for samp in (sampler, dsamp16):
            outs = [(v.clone(), l.clone()) for v, l in PrefetchLoader(ds, samp, batch_size=16, group=5)]
            assert [len(v) for v, _ in outs] == [16] * (len(order) // 16) + ([len(order) % 16] if len(order) % 16 else [])
            assert torch.cat([v for v, _ in outs]).cpu().numpy().tobytes() == images[order].tobytes()
            assert np.array_equal(torch.cat([l for _, l in outs]).cpu().numpy(), labels[order])
        ds.free()
        return True

    assert all(_world(P, body))


def test_streaming_ingest_into_initialised_shard():
    from ddstore_b200 import PyDDStore
    from ddstore_b200.dataset import ingest_chunks
    rng = np.random.default_rng(6)

    def body(comm, r):
        store = PyDDStore(comm, device=0)
        nrows = 5000 + 100 * r
        full = rng.integers(0, 2**32, size=(nrows, 24), dtype=np.uint32).view(np.float32)
        store.init("z", nrows, 24, 4)
        wrote = ingest_chunks(store, "z", (full[i:i + 777] for i in range(0, nrows, 777)))
        assert wrote == nrows
        first = ([0] + store.query("z")["lenlist"])[r]
        back = np.zeros_like(full)
        store.get("z", back, first)
        assert back.tobytes() == full.tobytes()
        with pytest.raises(ValueError):
            store.update("z", np.zeros((2, 24), np.float32), nrows - 1)  # past the end of the local shard
        store.free()
        store.close()
        return True

    assert all(_world(2, body))


def test_device_batch_sampler_equals_distributed_sampler():
    import torch
    from ddstore_b200.dataset import DeviceBatchSampler
    from torch.utils.data.distributed import DistributedSampler

    class L:
        def __len__(self):
            return 1003

    for rank in range(3):
        ref = DistributedSampler(L(), num_replicas=3, rank=rank, shuffle=True, seed=11)
        dev = DeviceBatchSampler(1003, 64, rank=rank, world_size=3, seed=11, device="cuda:0")
        for epoch in (0, 5):
            ref.set_epoch(epoch)
            dev.set_epoch(epoch)
            got = torch.cat([b for b in dev]).cpu().tolist()
            assert got == list(iter(ref)) and all(b.is_cuda for b in dev)
            assert len(dev) == (len(got) + 63) // 64
