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


def test_cv_tasks_cover_every_fit_once():
    for nf, nt, w in ((5, 20, 8), (5, 20, 1), (3, 2, 4), (5, 20, 7)):
        per_rank = bd.cv_tasks(nf, nt, w)
        assert len(per_rank) == w
        seen = sorted((f, t) for groups in per_rank for f, ts in groups for t in ts)
        assert seen == [(f, t) for f in range(nf) for t in range(nt)]
        sizes = [sum(len(ts) for _, ts in groups) for groups in per_rank]
        assert max(sizes) - min(sizes) <= 1
        assert max(len(groups) for groups in per_rank) <= 2 or w < nf  # a rank's contiguous chunk spans at most two folds


class _FakeStore:
    def __init__(self, keep):
        self.keep = keep

    def close(self):
        pass


def test_fit_cv_sharded_single_process():
    rng = np.random.default_rng(1)
    n, k = 30, 3
    Y = rng.normal(size=(n, k))
    folds = [np.arange(0, 10), np.arange(10, 20), np.arange(20, 30)]

    def fit(Yc, store, scale=1.0):
        assert Yc.shape[0] == len(store.keep) == 20
        return {"mu": Yc.mean(0) * scale, "b": np.tile(Yc.mean(0), (4, 1)), "hat": np.zeros((20, Yc.shape[1]))}

    out = bd.fit_cv_sharded(fit, _FakeStore, Y, folds, scale=3.0)
    want = np.array([Y[np.setdiff1d(np.arange(n), f)][:, t].mean() * 3.0 for f in folds for t in range(k)])
    assert np.allclose(out["mu"], want) and out["b"].shape == (4, 9) and "hat" not in out


def _cv_fit(Yc, store, scale=1.0):
    return {"mu": Yc.mean(0) * scale, "b": np.tile(Yc.mean(0), (4, 1)), "hat": np.zeros((len(store.keep), Yc.shape[1]))}


def _cv_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(30, 3))
    folds = [np.arange(0, 10), np.arange(10, 20), np.arange(20, 30)]
    loaded = []

    def load(keep):
        loaded.append(len(keep))
        return _FakeStore(keep)

    out = bd.fit_cv_sharded(_cv_fit, load, Y, folds, scale=3.0)
    q.put((rank, {k: v.tolist() for k, v in out.items()}, len(loaded)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_fit_cv_sharded_gloo_world2():
    """3 folds x 3 traits over two ranks: rank 0 gets tasks 0-4 (folds 0, 1), rank 1 tasks 5-8 (folds 1, 2); both end up with the
    full fold-major result, and neither builds more row-subset stores than the folds it touches."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(30, 3))
    folds = [np.arange(0, 10), np.arange(10, 20), np.arange(20, 30)]
    want = bd.fit_cv_sharded(_cv_fit, _FakeStore, Y, folds, scale=3.0)  # one rank, same data
    assert want["mu"].shape == (9,) and "hat" not in want
    for rank, res, nstores in got:
        assert nstores == 2, (rank, nstores)
        for key, v in want.items():
            assert np.allclose(np.asarray(res[key]), v), (rank, key)


class _StubStore:
    def __init__(self, Xk, device):
        self.X = np.asarray(Xk, dtype=np.float64)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


def _stub_panel_fit(model, yk, g):
    """Stands in for a solver of the panel: a ridge-flavoured closed form whose strength depends on the model name."""
    lam = {"m_weak": 1e3, "m_strong": 1e1}[model]
    xc = g.X - g.X.mean(0)
    return {"b": xc.T @ (yk - yk.mean()) / ((xc * xc).sum(0) + lam)}


def _cv_driver_case():
    rng = np.random.default_rng(2)
    X = rng.integers(0, 3, size=(40, 12)).astype(np.int8)
    y = X[:, :3].sum(1) + rng.normal(size=40)
    from bwgr_b200 import api
    return X, y, api._cv_holdouts(40, k=5, n=3, llo=None, seed=1)


def _cv_driver_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bwgr_b200 import api
    X, y, hs = _cv_driver_case()
    out = api._cv_run(("m_weak", "m_strong"), _stub_panel_fit, y, X, hs, None, False, True, store=_StubStore)
    q.put((rank, out["cv"], out["beta"].tolist(), out["hat"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_cv_driver_holdouts_over_two_ranks():
    """emCV / mcmcCV driver under torch.distributed (gloo, world 2): three hold-outs dealt 2 + 1, every rank returns the same
    summary as the single-process run."""
    import torch.multiprocessing as mp
    from bwgr_b200 import api
    X, y, hs = _cv_driver_case()
    want = api._cv_run(("m_weak", "m_strong"), _stub_panel_fit, y, X, hs, None, False, True, store=_StubStore)
    assert list(want["cv"]) == ["CV_1", "CV_2", "CV_3"] and want["beta"].shape == (12, 2) and want["hat"].shape == (40, 2)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cv_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, cv, beta, hat in got:
        assert cv == want["cv"], rank
        assert np.allclose(beta, want["beta"]) and np.allclose(hat, want["hat"]), rank
