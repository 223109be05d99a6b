"""world_size-2 gloo tests of the sharding plumbing (dist.py), CPU only.

The per-rank compute (pdc_gls_dev / pdc_pdm_dev on the rank's GPU) is replaced
by an oracle-backed stand-in returning CPU tensors, so what is covered is the
partition, the packed all-gather and the cross-rank arg-extremum reduction.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from periodicity_b200 import dist as pdist


def test_shard_bounds_cover_grid_exactly():
    for n in (1, 2, 7, 100, 101, 10_000_001):
        for world in (1, 2, 3, 8):
            covered = []
            L0 = None
            for r in range(world):
                a, b, L = pdist.shard_bounds(n, r, world)
                L0 = L0 or L
                assert L == L0 and 0 <= a <= b <= n and b - a <= L
                covered += list(range(a, b)) if n <= 101 else [(a, b)]
            if n <= 101:
                assert covered == list(range(n))
            else:
                assert covered[0][0] == 0 and covered[-1][1] == n
                assert all(covered[i][1] == covered[i + 1][0] for i in range(world - 1))


def test_reduce_best_nan_and_ties():
    assert pdist.reduce_best([1.0, 3.0, 3.0], [5, 20, 9], +1) == (9, 3.0)      # first occurrence
    assert pdist.reduce_best([np.nan, 2.0], [-1, 4], +1) == (4, 2.0)
    assert pdist.reduce_best([0.5, 0.25, 0.25], [0, 70, 30], -1) == (30, 0.25)
    idx, val = pdist.reduce_best([np.nan, np.nan], [-1, -1], +1)
    assert idx == -1 and np.isnan(val)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gls_compute(t, y, w, fmin, df, j0, n_local, fit_mean, psd_scale, device):
    from oracle import cport
    err = None if w is None else np.asarray(w) ** -0.5
    p = cport.gls_exact(t, y, err, fmin, df, n_local, fit_mean, psd_scale is not None, j0=j0)
    return (torch.from_numpy(p), torch.tensor([int(np.nanargmax(p))]), torch.tensor([float(np.nanmax(p))], dtype=torch.float64))


def _pdm_compute(t, x, periods, nb, nc, device):
    from oracle import cport
    th = cport.pdm(t, x, periods, nb, nc)
    return torch.from_numpy(th), torch.tensor([int(np.nanargmin(th))]), torch.tensor([float(np.nanmin(th))], dtype=torch.float64)


def _sl_compute(t, m, periods, device):
    from oracle import stringlength_numpy
    ell = stringlength_numpy.string_lengths(np.asarray(t), np.asarray(m), periods)
    return torch.from_numpy(ell), torch.tensor([int(np.nanargmin(ell))]), torch.tensor([float(np.nanmin(ell))], dtype=torch.float64)


def _batch_compute(t, y, w, offsets, fmin, df, nf, fit_mean, psd_scale, want_power, device):
    from oracle import cport
    B = len(offsets) - 1
    P = np.stack([cport.gls_exact(t[offsets[b]:offsets[b + 1]], y[offsets[b]:offsets[b + 1]], None, fmin[b], df[b], nf)
                  for b in range(B)])
    return (torch.from_numpy(P) if want_power else None, torch.from_numpy(np.nanargmax(P, axis=1)),
            torch.from_numpy(np.nanmax(P, axis=1)))


def _survey_inputs():
    rng = np.random.default_rng(17)
    sizes = [120, 80, 200, 150, 90]                       # 5 curves over 2 ranks: ragged, uneven split
    ts = [np.sort(rng.uniform(0, 30, n)) for n in sizes]
    ys = [np.sin(2 * np.pi * t / rng.uniform(1, 4)) + 0.3 * rng.standard_normal(t.size) for t in ts]
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    df = np.array([1 / (t[-1] - t[0]) / 5 for t in ts])
    return ts, ys, offsets, 0.5 * df, df


def _worker(rank, world, port, outdir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        t = np.sort(rng.uniform(0, 50, 300))
        y = np.sin(2 * np.pi * t * 1.37) + 0.3 * rng.standard_normal(300)
        df = 1 / (t[-1] - t[0]) / 5
        nf = 501                                               # odd: last shard is short
        power, idx, val = pdist.gls_sharded(t, y, None, 0.5 * df, df, nf, True, None, compute=_gls_compute)
        periods = np.linspace(0.3, 3.0, 77)
        theta, pidx, pval = pdist.pdm_sharded(t, y, periods, 5, 2, compute=_pdm_compute)
        from oracle import stringlength_numpy
        ell, lidx, lval = pdist.stringlength_sharded(t, stringlength_numpy.scale(y), periods, compute=_sl_compute)
        ts, ys, offsets, f0, dfs = _survey_inputs()
        bp, barg, bmx = pdist.gls_batch_sharded(np.concatenate(ts), np.concatenate(ys), None, offsets, f0, dfs, 64,
                                                want_power=True, compute=_batch_compute)
        _, barg2, bmx2 = pdist.gls_batch_sharded(np.concatenate(ts), np.concatenate(ys), None, offsets, f0, dfs, 64,
                                                 want_power=False, compute=_batch_compute)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), power=power, idx=idx, val=val, theta=theta,
                 pidx=pidx, pval=pval, ell=ell, lidx=lidx, lval=lval, bp=bp, barg=barg, bmx=bmx, barg2=barg2, bmx2=bmx2)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_calls_world2_gloo(tmp_path):
    from oracle import cport
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    t = np.sort(rng.uniform(0, 50, 300))
    y = np.sin(2 * np.pi * t * 1.37) + 0.3 * rng.standard_normal(300)
    df = 1 / (t[-1] - t[0]) / 5
    want = cport.gls_exact(t, y, None, 0.5 * df, df, 501)
    want_theta = cport.pdm(t, y, np.linspace(0.3, 3.0, 77), 5, 2)
    from oracle import stringlength_numpy
    want_ell = stringlength_numpy.string_lengths(t, stringlength_numpy.scale(y), np.linspace(0.3, 3.0, 77))
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        # every rank ends with the full periodogram and the global peak
        np.testing.assert_allclose(z["power"], want, rtol=1e-12)
        assert int(z["idx"]) == int(np.nanargmax(want)) and float(z["val"]) == float(np.nanmax(z["power"]))
        np.testing.assert_allclose(z["theta"], want_theta, rtol=1e-12)
        assert int(z["pidx"]) == int(np.nanargmin(want_theta))
        np.testing.assert_array_equal(z["ell"], want_ell)
        assert int(z["lidx"]) == int(np.nanargmin(want_ell)) and float(z["lval"]) == float(np.nanmin(want_ell))
        # batch sharding: every rank ends with every curve's periodogram and peak
        ts, ys, offsets, f0, dfs = _survey_inputs()
        for b in range(5):
            ref = cport.gls_exact(ts[b], ys[b], None, f0[b], dfs[b], 64)
            np.testing.assert_allclose(z["bp"][b], ref, rtol=1e-12)
            assert z["barg"][b] == z["barg2"][b] == np.nanargmax(ref)
            assert z["bmx"][b] == z["bmx2"][b] == np.nanmax(z["bp"][b])
