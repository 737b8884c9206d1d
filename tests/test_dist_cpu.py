"""N>1 host logic on CPU: world_size-2 gloo run of the replica sharding (independent systems, no data-path collective)."""
import os
import socket

import numpy as np
import pytest

from bwgr_b200 import dist as bd


def test_partition_balanced():
    assert bd.partition(100, 8) == [(0, 13), (13, 26), (26, 39), (39, 52), (52, 64), (64, 76), (76, 88), (88, 100)]
    assert bd.partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n, w in ((1, 1), (7, 2), (20, 8), (0, 3)):
        parts = bd.partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [e - s for s, e in parts]
        assert max(sizes) - min(sizes) <= 1


def _stub_fit(Ycols, row_mask=None, scale=1.0):
    """Stands in for bw.em_fit on CPU: per-system outputs with the shapes of the real result dict."""
    n, k = Ycols.shape
    w = np.ones_like(Ycols) if row_mask is None else row_mask.astype(float)
    mu = (Ycols * w).sum(0) / w.sum(0)
    return {"mu": mu * scale, "b": np.tile(mu, (5, 1)), "h2": np.full(k, 0.5), "its": np.full(k, 7, dtype=np.int32)}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(50, 5))
    mask = rng.random((50, 5)) < 0.8
    out = bd.fit_sharded(_stub_fit, Y, row_mask=mask, scale=2.0)
    q.put((rank, {k: v.tolist() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_fit_sharded_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(50, 5))
    mask = rng.random((50, 5)) < 0.8
    want = _stub_fit(Y, row_mask=mask, scale=2.0)
    for r in (0, 1):  # every rank holds the full, ordered result
        for key, v in want.items():
            assert np.allclose(np.asarray(res[r][key]), v), (r, key)


def test_fit_sharded_single_process():
    Y = np.arange(12.0).reshape(4, 3)
    out = bd.fit_sharded(_stub_fit, Y)
    assert np.allclose(out["mu"], Y.mean(0)) and out["b"].shape == (5, 3)
