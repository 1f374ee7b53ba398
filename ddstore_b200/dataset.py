"""ddstore_b200/dataset.py -- the loader side of the hot path.

`DistDataset` is the reference's `examples/vae/distdataset.py:13-92` re-built on the batched fetch:
same constructor shape (`data, label, comm, ddstore_width`), same per-sample `__getitem__` contract, plus
`__getitems__` -- which torch's DataLoader calls with the whole index batch -- so a batch costs one kernel
launch per variable and lands packed in HBM instead of one blocking round trip per sample followed by
`torch.tensor` + collate + `.to(device)` copies (`examples/vae/distdataset.py:84-88`, `vae-ddp.py:244`).

Epoch shuffle is the caller's sampler exactly as in the reference (`DistributedSampler`, `vae-ddp.py:216`):
`make_loader()` wires `DistributedSampler(shuffle=True)` + `set_epoch` + the identity collate.

(The reference flattens every sample into a disp=1 variable and then passes the SAMPLE index as the ROW index
(`distdataset.py:63,70,84`), so it returns floats [idx, idx+784) rather than image idx; here a sample is one row
of width `sample_size`, which is what the reference evidently meant.)
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from .comm import as_dds_comm
from .store import PyDDStore


def nsplit(a, n):
    """contiguous near-equal split (same arithmetic as examples/vae/distdataset.py:9-11)"""
    k, m = divmod(len(a), n)
    return (a[i * k + min(i, m):(i + 1) * k + min(i + 1, m)] for i in range(n))


class DistDataset(Dataset):
    """Fixed-shape samples + integer labels, sharded over the ranks of `comm` by contiguous blocks.

    data: sequence of (tensor/ndarray, label) pairs -- every rank passes the same sequence (like the reference) and
    keeps only its block; or pass `local_only=True` when `data` already is this rank's block.
    """

    def __init__(self, data, label, comm=None, ddstore_width=None, device=None, local_only=False):
        super().__init__()
        self.label = label
        self.comm = as_dds_comm(comm)
        self.rank, self.comm_size = self.comm.Get_rank(), self.comm.Get_size()
        if ddstore_width is not None and ddstore_width != self.comm_size:
            # the reference splits the communicator into replica groups of this width (distdataset.py:25-30);
            # here a store spans one NVSwitch box, so pass the per-box communicator instead
            raise NotImplementedError("pass the per-box communicator instead of ddstore_width")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ddstore = PyDDStore(self.comm, device=self.device.index)

        if local_only:
            mine = list(range(len(data)))
            counts = [int(c) for c in self._allgather_int(len(data))]
            self.total_ns = sum(counts)
        else:
            self.total_ns = len(data)
            mine = list(nsplit(range(len(data)), self.comm_size))[self.rank]
        vals, labels = [], []
        for i in mine:
            d, lab = data[i]
            d = d.cpu().numpy() if hasattr(d, "cpu") else np.asarray(d)
            vals.append(np.ascontiguousarray(d).reshape(-1))
            labels.append(lab)
        self.sample_shape = tuple(np.asarray(data[mine[0]][0]).shape) if mine else ()
        arr = np.stack(vals) if vals else np.zeros((0, 1), np.float32)
        self.sample_size = arr.shape[1]
        self.dtype = torch.from_numpy(arr[:0]).dtype
        self.ddstore.add(f"{self.label}data", np.ascontiguousarray(arr))
        self.ddstore.add(f"{self.label}labels", np.ascontiguousarray(np.array(labels, dtype=np.int32).reshape(-1, 1)))

    def _allgather_int(self, v):
        parts = self.comm.allgather_bytes(int(v).to_bytes(8, "little"))
        return [int.from_bytes(p, "little") for p in parts]

    def len(self):
        return self.total_ns

    def __len__(self):
        return self.total_ns

    # ---- the reference's per-sample contract (distdataset.py:79-92)
    def get(self, idx):
        val = torch.empty(self.sample_size, dtype=self.dtype, device=self.device)
        lab = torch.empty(1, dtype=torch.int32, device=self.device)
        self.ddstore.get(f"{self.label}data", val.view(1, -1), int(idx))
        self.ddstore.get(f"{self.label}labels", lab.view(1, 1), int(idx))
        return val.view(self.sample_shape), int(lab.item())

    def __getitem__(self, idx):
        return self.get(idx)

    # ---- the batched contract: DataLoader hands over the whole index list
    def __getitems__(self, indices):
        B = len(indices)
        idx = np.asarray(indices, dtype=np.int64)
        vals = torch.empty((B, self.sample_size), dtype=self.dtype, device=self.device)
        labs = torch.empty((B, 1), dtype=torch.int32, device=self.device)
        self.ddstore.get_batch(f"{self.label}data", idx, out=vals, count=1)
        self.ddstore.get_batch(f"{self.label}labels", idx, out=labs, count=1)
        return vals.view((B,) + self.sample_shape), labs.view(B)

    @staticmethod
    def collate(batch):
        """identity: __getitems__ already returns the collated, device-resident batch"""
        return batch

    def epoch_begin(self):
        self.ddstore.epoch_begin()

    def epoch_end(self):
        self.ddstore.epoch_end()

    def free(self):
        self.ddstore.free()


class RaggedDataset(Dataset):
    """Variable-length, multi-array samples (configs 3 and 4): every variable is a 2-D array of rows, and sample i
    owns rows [row_start[v][i], row_start[v][i] + row_count[v][i]) of variable v -- the HydraGNN-style layout the
    reference's get(name, arr, start) with count = arr.shape[0] implies (src/pyddstore.pyx:84-87).
    The (start, count) tables of ALL samples are kept on the device, so a batch needs only the sample ids."""

    def __init__(self, local_arrays, local_counts, comm=None, device=None):
        """local_arrays: {name: 2-D ndarray of this rank's rows}; local_counts: {name: int64[n_local_samples]}"""
        super().__init__()
        self.comm = as_dds_comm(comm)
        self.rank, self.comm_size = self.comm.Get_rank(), self.comm.Get_size()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ddstore = PyDDStore(self.comm, device=self.device.index)
        self.names = list(local_arrays)
        self.row_bytes, self.dtypes, self.widths = {}, {}, {}
        self.starts, self.counts = {}, {}
        n_local = len(next(iter(local_counts.values())))
        for name in self.names:
            arr = np.ascontiguousarray(local_arrays[name])
            cnt = np.ascontiguousarray(local_counts[name], dtype=np.int64)
            assert cnt.sum() == arr.shape[0] and len(cnt) == n_local
            self.ddstore.add(name, arr)
            first_row = ([0] + self.ddstore.query(name)["lenlist"])[self.rank]
            local_start = first_row + np.concatenate([[0], np.cumsum(cnt)[:-1]])
            # every rank learns every sample's (start, count): 16 B per sample per variable
            blobs = self.comm.allgather_bytes(len(cnt).to_bytes(8, "little"))
            sizes = [int.from_bytes(b, "little") for b in blobs]
            pad = max(sizes)
            buf = np.zeros((2, pad), np.int64)
            buf[0, :len(cnt)], buf[1, :len(cnt)] = local_start, cnt
            parts = self.comm.allgather_bytes(buf.tobytes())
            tabs = [np.frombuffer(p, np.int64).reshape(2, pad)[:, :n] for p, n in zip(parts, sizes)]
            all_start = np.concatenate([t[0] for t in tabs])
            all_count = np.concatenate([t[1] for t in tabs])
            self.ddstore.set_sample_index(name, all_start, all_count)  # device-resident (start, count) of every sample
            self.counts[name] = all_count
            self.row_bytes[name] = arr.dtype.itemsize * int(np.prod(arr.shape[1:], dtype=np.int64))
            self.dtypes[name] = torch.from_numpy(arr[:0]).dtype
            self.widths[name] = arr.shape[1:]
        self.total_ns = int(len(self.counts[self.names[0]]))

    def __len__(self):
        return self.total_ns

    def __getitem__(self, idx):
        return self.__getitems__([idx])

    def __getitems__(self, indices):
        """-> {name: (packed rows tensor [sum(count), ...width], int64 row offsets per sample [B+1])}.
        One launch chain per variable; the sample-id -> (start, count) lookup happens on the device."""
        ids = np.asarray(indices, dtype=np.int64)
        d_ids = torch.from_numpy(ids).to(self.device, non_blocking=True)
        out = {}
        for name in self.names:
            rows = int(self.counts[name][ids].sum())  # host-side size of the packed result (sizes only, no data)
            buf = torch.empty((max(rows, 1),) + tuple(self.widths[name]), dtype=self.dtypes[name], device=self.device)
            offs = torch.empty(len(ids) + 1, dtype=torch.int64, device=self.device)
            self.ddstore.get_samples(name, d_ids, out=buf, offsets=offs)
            out[name] = (buf[:rows], offs // self.row_bytes[name])
        return out

    collate = staticmethod(lambda batch: batch)

    def free(self):
        self.ddstore.free()


def make_loader(dataset, batch_size, rank=0, world_size=1, shuffle=True, seed=0, drop_last=False):
    """DataLoader over a DistDataset/RaggedDataset with the reference's epoch shuffle (DistributedSampler,
    examples/vae/vae-ddp.py:216-219): call loader.sampler.set_epoch(e) at the top of every epoch."""
    sampler = DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=shuffle, seed=seed,
                                 drop_last=drop_last)
    return DataLoader(dataset, batch_size=batch_size, sampler=sampler, num_workers=0, collate_fn=dataset.collate,
                      drop_last=drop_last)
