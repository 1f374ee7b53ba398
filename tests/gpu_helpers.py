"""helpers for the -m gpu parity tests: thread-rank worlds over one or more GPUs (test infrastructure)"""
import threading
import uuid

import numpy as np


def run_world(P, body, devices=None, timeout=300):
    """Run `body(store, rank)` on P thread-ranks (ShmComm across threads). devices[r] = CUDA ordinal of
    rank r (default: all on device 0 -- owner lookup and peer-table logic are exercised even on one GPU;
    with >= P GPUs pass range(P) for real NVLink peer loads)."""
    from ddstore_b200 import PyDDStore, ShmComm
    key = "g" + uuid.uuid4().hex[:12]
    res, errs = [None] * P, []

    def run(r):
        comm = store = None
        try:
            comm = ShmComm(key, r, P)
            store = PyDDStore(comm, device=(devices[r] if devices else 0))
            res[r] = body(store, r)
            store.free()
        except BaseException as e:  # noqa: BLE001
            import traceback
            errs.append((r, traceback.format_exc()))
        finally:
            if store is not None:
                store.close()
            if comm is not None:
                comm.close()

    th = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join(timeout) for t in th]
    assert not errs, "\n".join(f"rank {r}: {tb}" for r, tb in errs)
    return res


def packed_nbytes(counts, row):
    return int(np.clip(np.asarray(counts, np.int64), 0, None).sum()) * row
