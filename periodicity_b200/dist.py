"""Torch-tensor entry points and multi-GPU sharding (one process per GPU).

The trial-frequency search partitions trivially: every
frequency / trial period / light curve is independent.  Rank ``r`` of ``W``
evaluates a contiguous, equally padded slice of the grid through the
device-pointer C ABI (``pdc_gls_dev`` / ``pdc_pdm_dev``) and the ranks exchange
ONE packed all-gather per call: ``[slice values (L doubles), local best value,
local best global index]``; every rank then reduces the ``W`` (best, index)
pairs, so all ranks end with the full periodogram and the global arg-extremum.
``torch.distributed`` (NCCL over NVLink for CUDA tensors, gloo in the CPU tests)
is plumbing only; there is no collective inside the hot kernels because the path
has no exchange step.

The reference's counterpart is the ``multiprocessing.Pool`` fan-out of
``phase.py:185-186`` (and nothing for GLS).
"""
import math

import numpy as np

from . import _ffi


def shard_bounds(n_units, rank, world):
    """(start, stop, L): contiguous slice of rank ``rank`` and the padded slice length L."""
    L = max(1, math.ceil(n_units / world))
    start = min(n_units, rank * L)
    stop = min(n_units, start + L)
    return start, stop, L


def reduce_best(best_vals, best_idx, sign):
    """NaN-ignoring arg-extremum over per-rank candidates, first occurrence on ties.

    ``sign=+1`` maximum (np.nanargmax, reference core.py:202-205), ``-1`` minimum.
    Returns (index, value); (-1, nan) if there is no candidate."""
    best_vals = np.asarray(best_vals, dtype=np.float64)
    best_idx = np.asarray(best_idx, dtype=np.int64)
    ok = (best_idx >= 0) & ~np.isnan(best_vals)
    if not ok.any():
        return -1, float("nan")
    v = np.where(ok, best_vals, -np.inf if sign > 0 else np.inf)
    target = v.max() if sign > 0 else v.min()
    cand = best_idx[ok & (v == target)]
    return int(cand.min()), float(target)


def _torch():
    import torch
    return torch


def _dist_info(group=None):
    torch = _torch()
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


def all_gather_packed(local_vals, local_best, local_arg, L, group=None):
    """All-gather ``[vals padded to L, best, arg]`` (float64 tensor on any device).

    Returns (vals_all [W, L], best_all [W], arg_all [W]) as tensors on the input device."""
    torch = _torch()
    dist, rank, world = _dist_info(group)
    packed = torch.full((L + 2,), float("nan"), dtype=torch.float64, device=local_vals.device)
    packed[: local_vals.numel()] = local_vals
    packed[L] = local_best
    packed[L + 1] = local_arg
    if dist is None or world == 1:
        allp = packed.unsqueeze(0)
    else:
        allp = torch.empty((world, L + 2), dtype=torch.float64, device=local_vals.device)
        dist.all_gather_into_tensor(allp.view(-1), packed, group=group)
    return allp[:, :L], allp[:, L], allp[:, L + 1]


# ---------------------------------------------------------------------------
# torch-tensor entry points (single device, stream ordered, no host sync)
# ---------------------------------------------------------------------------
def gls_torch(t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, j0=0, ctx=None):
    """GLS power for CUDA float64 tensors; returns (power[nf], argmax[1] int64, max[1]) tensors."""
    torch = _torch()
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("t must be a contiguous CUDA float64 tensor")
    ctx = ctx or _ffi.default_context(t.device.index)
    y = y.contiguous()
    w = None if w is None else w.contiguous()
    power = torch.empty(int(nf), dtype=torch.float64, device=t.device)
    arg = torch.empty(1, dtype=torch.int64, device=t.device)
    mx = torch.empty(1, dtype=torch.float64, device=t.device)
    flags = (_ffi.GLS_FIT_MEAN if fit_mean else 0) | (_ffi.GLS_PSD if psd_scale is not None else 0)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    ctx.gls_dev(t.data_ptr(), y.data_ptr(), 0 if w is None else w.data_ptr(), t.numel(), fmin, df, j0, nf,
                flags, 1.0 if psd_scale is None else psd_scale, power.data_ptr(), arg.data_ptr(),
                mx.data_ptr(), stream)
    return power, arg, mx


def pdm_torch(t, x, periods, nb, nc, ctx=None):
    """PDM theta for CUDA float64 tensors; returns (theta[np], argmin[1] int64, min[1]) tensors."""
    torch = _torch()
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("t must be a contiguous CUDA float64 tensor")
    ctx = ctx or _ffi.default_context(t.device.index)
    x = x.contiguous()
    periods = periods.contiguous()
    theta = torch.empty(periods.numel(), dtype=torch.float64, device=t.device)
    arg = torch.empty(1, dtype=torch.int64, device=t.device)
    mn = torch.empty(1, dtype=torch.float64, device=t.device)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    ctx.pdm_dev(t.data_ptr(), x.data_ptr(), t.numel(), periods.data_ptr(), periods.numel(), nb, nc,
                theta.data_ptr(), arg.data_ptr(), mn.data_ptr(), stream)
    return theta, arg, mn


def ce_torch(t, x, periods, nphi, nm, ctx=None):
    """Conditional entropy for CUDA float64 tensors; returns (h[np], argmin[1] int64, min[1]) tensors."""
    torch = _torch()
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("t must be a contiguous CUDA float64 tensor")
    ctx = ctx or _ffi.default_context(t.device.index)
    x = x.contiguous()
    periods = periods.contiguous()
    h = torch.empty(periods.numel(), dtype=torch.float64, device=t.device)
    arg = torch.empty(1, dtype=torch.int64, device=t.device)
    mn = torch.empty(1, dtype=torch.float64, device=t.device)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    ctx.ce_dev(t.data_ptr(), x.data_ptr(), t.numel(), periods.data_ptr(), periods.numel(), nphi, nm,
               h.data_ptr(), arg.data_ptr(), mn.data_ptr(), stream)
    return h, arg, mn


def stringlength_torch(t, m, periods, ctx=None):
    """String length for CUDA float64 tensors; returns (ell[np], argmin[1] int64, min[1]) tensors."""
    torch = _torch()
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("t must be a contiguous CUDA float64 tensor")
    ctx = ctx or _ffi.default_context(t.device.index)
    m = m.contiguous()
    periods = periods.contiguous()
    ell = torch.empty(periods.numel(), dtype=torch.float64, device=t.device)
    arg = torch.empty(1, dtype=torch.int64, device=t.device)
    mn = torch.empty(1, dtype=torch.float64, device=t.device)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    ctx.stringlength_dev(t.data_ptr(), m.data_ptr(), t.numel(), periods.data_ptr(), periods.numel(),
                         ell.data_ptr(), arg.data_ptr(), mn.data_ptr(), stream)
    return ell, arg, mn


def gls_batch_torch(t, y, w, offsets, fmin, df, nf, fit_mean=True, psd_scale=None, want_power=True, ctx=None):
    """Batched GLS for CUDA tensors (curves back to back); offsets/fmin/df are host arrays."""
    torch = _torch()
    ctx = ctx or _ffi.default_context(t.device.index)
    B = len(offsets) - 1
    power = torch.empty((B, int(nf)), dtype=torch.float64, device=t.device) if want_power else None
    arg = torch.empty(B, dtype=torch.int64, device=t.device)
    mx = torch.empty(B, dtype=torch.float64, device=t.device)
    flags = (_ffi.GLS_FIT_MEAN if fit_mean else 0) | (_ffi.GLS_PSD if psd_scale is not None else 0)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    ctx.gls_batch_dev(t.data_ptr(), y.data_ptr(), 0 if w is None else w.data_ptr(), offsets, fmin, df, nf, flags,
                      psd_scale, 0 if power is None else power.data_ptr(), arg.data_ptr(), mx.data_ptr(), stream)
    return power, arg, mx


# ---------------------------------------------------------------------------
# sharded calls (numpy in, numpy out on every rank)
# ---------------------------------------------------------------------------
def _device_compute_gls(t, y, w, fmin, df, j0, n_local, fit_mean, psd_scale, device):
    torch = _torch()
    dev = torch.device("cuda", _ffi.default_context(device).device)
    tt = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).to(dev, non_blocking=True)
    yy = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64)).to(dev, non_blocking=True)
    ww = None if w is None else torch.as_tensor(np.ascontiguousarray(w, dtype=np.float64)).to(dev, non_blocking=True)
    power, arg, mx = gls_torch(tt, yy, ww, fmin, df, n_local, fit_mean, psd_scale, j0=j0,
                               ctx=_ffi.default_context(device))
    return power, arg, mx


def gls_sharded(t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, device=None, group=None, compute=None):
    """Frequency-grid-sharded GLS.  Every rank passes the same inputs and gets the full result.

    ``compute(t, y, w, fmin, df, j0, n_local, fit_mean, psd_scale, device)`` must return
    float64 tensors ``(power[n_local], argmax[1] (local, int64), max[1])``; the default runs
    ``pdc_gls_dev`` on this rank's GPU.  (The CPU gloo tests inject a stand-in.)
    """
    torch = _torch()
    _, rank, world = _dist_info(group)
    start, stop, L = shard_bounds(nf, rank, world)
    compute = compute or _device_compute_gls
    if stop > start:
        power, arg, mx = compute(t, y, w, fmin, df, start, stop - start, fit_mean, psd_scale, device)
        garg = torch.where(arg >= 0, arg + start, arg).to(torch.float64)
        best = mx.reshape(())
        garg = garg.reshape(())
    else:
        ref = compute(t, y, w, fmin, df, 0, 1, fit_mean, psd_scale, device)[0]  # keeps device/dtype
        power = ref[:0]
        best = torch.tensor(float("nan"), dtype=torch.float64, device=ref.device)
        garg = torch.tensor(-1.0, dtype=torch.float64, device=ref.device)
    vals, bests, args = all_gather_packed(power, best, garg, L, group)
    full = vals.reshape(-1)[:nf].cpu().numpy()
    idx, val = reduce_best(bests.cpu().numpy(), args.cpu().numpy().astype(np.int64), +1)
    return full, idx, val


def _device_compute_pdm(t, x, periods, nb, nc, device):
    torch = _torch()
    dev = torch.device("cuda", _ffi.default_context(device).device)
    tt = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).to(dev, non_blocking=True)
    xx = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev, non_blocking=True)
    pp = torch.as_tensor(np.ascontiguousarray(periods, dtype=np.float64)).to(dev, non_blocking=True)
    return pdm_torch(tt, xx, pp, nb, nc, ctx=_ffi.default_context(device))


def pdm_sharded(t, x, periods, nb, nc, device=None, group=None, compute=None):
    """Period-grid-sharded PDM; theta returned in the order of ``periods`` on every rank."""
    torch = _torch()
    _, rank, world = _dist_info(group)
    periods = np.ascontiguousarray(periods, dtype=np.float64)
    npd = periods.size
    start, stop, L = shard_bounds(npd, rank, world)
    compute = compute or _device_compute_pdm
    if stop > start:
        theta, arg, mn = compute(t, x, periods[start:stop], nb, nc, device)
        garg = torch.where(arg >= 0, arg + start, arg).to(torch.float64).reshape(())
        best = mn.reshape(())
    else:
        ref = compute(t, x, periods[:1], nb, nc, device)[0]
        theta = ref[:0]
        best = torch.tensor(float("nan"), dtype=torch.float64, device=ref.device)
        garg = torch.tensor(-1.0, dtype=torch.float64, device=ref.device)
    vals, bests, args = all_gather_packed(theta, best, garg, L, group)
    full = vals.reshape(-1)[:npd].cpu().numpy()
    idx, val = reduce_best(bests.cpu().numpy(), args.cpu().numpy().astype(np.int64), -1)
    return full, idx, val


def _device_compute_sl(t, m, periods, device):
    torch = _torch()
    dev = torch.device("cuda", _ffi.default_context(device).device)
    tt = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).to(dev, non_blocking=True)
    mm = torch.as_tensor(np.ascontiguousarray(m, dtype=np.float64)).to(dev, non_blocking=True)
    pp = torch.as_tensor(np.ascontiguousarray(periods, dtype=np.float64)).to(dev, non_blocking=True)
    return stringlength_torch(tt, mm, pp, ctx=_ffi.default_context(device))


def stringlength_sharded(t, m, periods, device=None, group=None, compute=None):
    """Period-grid-sharded String Length (trial periods are independent): every rank evaluates a
    contiguous slice and one all-gather of ``[lengths, local min, local argmin]`` gives every rank the whole
    periodogram, in the order of ``periods``."""
    compute = compute or _device_compute_sl
    return pdm_sharded(t, m, periods, None, None, device=device, group=group,
                       compute=lambda t_, m_, p_, nb_, nc_, dev_: compute(t_, m_, p_, dev_))


def batch_shard_bounds(n_curves, rank, world):
    """Contiguous group of light curves owned by ``rank`` (survey workload)."""
    start, stop, _ = shard_bounds(n_curves, rank, world)
    return start, stop


def _device_compute_gls_batch(t, y, w, offsets, fmin, df, nf, fit_mean, psd_scale, want_power, device):
    torch = _torch()
    ctx = _ffi.default_context(device)
    dev = torch.device("cuda", ctx.device)
    a, e = int(offsets[0]), int(offsets[-1])
    tt = torch.as_tensor(np.ascontiguousarray(t[a:e], dtype=np.float64)).to(dev, non_blocking=True)
    yy = torch.as_tensor(np.ascontiguousarray(y[a:e], dtype=np.float64)).to(dev, non_blocking=True)
    ww = None if w is None else torch.as_tensor(np.ascontiguousarray(w[a:e], dtype=np.float64)).to(dev, non_blocking=True)
    return gls_batch_torch(tt, yy, ww, np.asarray(offsets) - a, fmin, df, nf, fit_mean, psd_scale, want_power, ctx=ctx)


def gls_batch_sharded(t, y, w, offsets, fmin, df, nf, fit_mean=True, psd_scale=None, want_power=False,
                      device=None, group=None, compute=None):
    """Light-curve-batch-sharded GLS (survey workload).

    Curves are split into contiguous groups, one per rank; each rank evaluates only its own curves
    (no replication) and ONE all-gather returns every curve's (argmax, max) -- and the powers when
    ``want_power`` -- to all ranks.  Returns ``(power [B, nf] or None, argmax [B], max [B])`` as numpy.
    """
    torch = _torch()
    dist, rank, world = _dist_info(group)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = offsets.size - 1
    fmin = np.ascontiguousarray(np.broadcast_to(fmin, (B,)), dtype=np.float64)
    df = np.ascontiguousarray(np.broadcast_to(df, (B,)), dtype=np.float64)
    if psd_scale is not None:
        psd_scale = np.ascontiguousarray(np.broadcast_to(psd_scale, (B,)), dtype=np.float64)
    b0, b1, L = shard_bounds(B, rank, world)
    compute = compute or _device_compute_gls_batch
    width = (nf if want_power else 0) + 2
    if b1 > b0:
        power, arg, mx = compute(t, y, w, offsets[b0:b1 + 1], fmin[b0:b1], df[b0:b1], nf, fit_mean,
                                 None if psd_scale is None else psd_scale[b0:b1], want_power, device)
        packed = torch.full((L, width), float("nan"), dtype=torch.float64, device=mx.device)
        if want_power:
            packed[: b1 - b0, :nf] = power
        packed[: b1 - b0, width - 2] = mx
        packed[: b1 - b0, width - 1] = arg.to(torch.float64)
    else:
        ref = compute(t, y, w, offsets[:2], fmin[:1], df[:1], nf, fit_mean,
                      None if psd_scale is None else psd_scale[:1], False, device)[2]
        packed = torch.full((L, width), float("nan"), dtype=torch.float64, device=ref.device)
    if dist is None or world == 1:
        allp = packed
    else:
        allp = torch.empty((world * L, width), dtype=torch.float64, device=packed.device)
        dist.all_gather_into_tensor(allp.view(-1), packed.view(-1), group=group)
    allp = allp[:B].cpu().numpy()
    power = allp[:, :nf] if want_power else None
    arg = np.where(np.isnan(allp[:, width - 1]), -1, allp[:, width - 1]).astype(np.int64)
    return power, arg, allp[:, width - 2]


# ---------------------------------------------------------------------------
# fused epilogue + all-gather over NVLink peer memory (no NCCL call on the data path)
# ---------------------------------------------------------------------------
_symm_cache = {}


def _symm_buffer(n_doubles, device, group):
    """Symmetric float64 buffer of ``n_doubles`` on every rank, mapped into every process
    (torch.distributed._symmetric_memory); cached per (size, device).

    TWO buffers per key, used alternately (every rank makes the same sequence of calls, so all ranks pick the same one).
    That removes the barrier a single buffer needs BEFORE the kernel ("every rank has finished reading the previous
    result"): call k+2 is the first to overwrite the buffer of call k, a peer's kernel of call k+2 starts only after it
    has passed the barrier that ends call k+1, and that barrier completes only when this rank's stream has reached it --
    i.e. after everything this rank enqueued between call k and call k+1, which includes its reads of result k (result
    views are valid until the next call on the same stream)."""
    torch = _torch()
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    key = (int(n_doubles), str(device))
    ent = _symm_cache.get(key)
    if ent is None:
        pair = []
        for _ in range(2):
            buf = symm_mem.empty(int(n_doubles), dtype=torch.float64, device=device)
            hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            pair.append((buf, hdl))
        ent = [pair, 0]
        _symm_cache[key] = ent
    ent[1] ^= 1
    return ent[0][ent[1]]


def gls_sharded_p2p_torch(t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, ctx=None, group=None):
    """Frequency-grid-sharded GLS for CUDA tensors with the all-gather fused into the epilogue kernel.

    Every rank evaluates its slice; the epilogue stores each power value directly into the symmetric
    result buffer of ALL ranks through NVLink peer mappings (``pdc_gls_dev_fanout``), so when the
    kernel ends the "all-gather" has already happened.  ONE device-side barrier on the current
    stream (``handle.barrier``) ends the call; two alternating result buffers make a barrier before the
    kernel unnecessary (``_symm_buffer``).  Returns views ``(power[nf], best[W, 2])``
    of this rank's symmetric buffer (valid until the next call on this stream) -- ``best[r] = (max, global argmax)``
    of rank r's slice."""
    torch = _torch()
    dist, rank, world = _dist_info(group)
    if dist is None or world < 2:
        raise RuntimeError("gls_sharded_p2p_torch needs an initialised process group with >= 2 ranks")
    if world > _ffi.MAX_PEERS:
        raise ValueError(f"at most {_ffi.MAX_PEERS} ranks")
    ctx = ctx or _ffi.default_context(t.device.index)
    buf, hdl = _symm_buffer(int(nf) + 2 * world, t.device, group)
    start, stop, _ = shard_bounds(nf, rank, world)
    fan = _fanout_struct(hdl, nf, world, rank)
    flags = (_ffi.GLS_FIT_MEAN if fit_mean else 0) | (_ffi.GLS_PSD if psd_scale is not None else 0)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    if stop > start:
        ctx.gls_dev_fanout(t.data_ptr(), y.data_ptr(), 0 if w is None else w.data_ptr(), t.numel(), fmin, df,
                           start, stop - start, flags, 1.0 if psd_scale is None else psd_scale, fan, stream)
    else:
        buf[int(nf) + 2 * rank: int(nf) + 2 * rank + 2] = torch.tensor([float("nan"), -1.0], dtype=torch.float64,
                                                                        device=t.device)
    hdl.barrier(channel=0)        # every rank's stores have landed everywhere (the only barrier: see _symm_buffer)
    return buf[: int(nf)], buf[int(nf):].view(world, 2)


def _fanout_struct(hdl, n_values, world, rank):
    fan = _ffi.Fanout()
    fan.world, fan.rank = world, rank
    ptrs = hdl.buffer_ptrs
    for r in range(world):
        fan.power[r] = int(ptrs[r])
        fan.best[r] = int(ptrs[r]) + 8 * int(n_values)
    return fan


def pdm_sharded_p2p_torch(t, x, periods, nb, nc, ctx=None, group=None):
    """Period-grid-sharded PDM for CUDA tensors, all-gather fused into the epilogue (``pdc_pdm_dev_fanout``).
    ``periods`` is the FULL grid (CUDA tensor) on every rank.  Returns views ``(theta[np], best[W, 2])``."""
    torch = _torch()
    dist, rank, world = _dist_info(group)
    if dist is None or world < 2:
        raise RuntimeError("pdm_sharded_p2p_torch needs an initialised process group with >= 2 ranks")
    ctx = ctx or _ffi.default_context(t.device.index)
    npd = periods.numel()
    buf, hdl = _symm_buffer(npd + 2 * world, t.device, group)
    start, stop, _ = shard_bounds(npd, rank, world)
    fan = _fanout_struct(hdl, npd, world, rank)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    if stop > start:
        ctx.pdm_dev_fanout(t.data_ptr(), x.data_ptr(), t.numel(), periods.data_ptr() + 8 * start, stop - start,
                           nb, nc, start, fan, stream)
    else:
        buf[npd + 2 * rank: npd + 2 * rank + 2] = torch.tensor([float("nan"), -1.0], dtype=torch.float64,
                                                              device=t.device)
    hdl.barrier(channel=0)
    return buf[:npd], buf[npd:].view(world, 2)


_dev_cache = {}
_pin_cache = {}


def _upload_cached(name, arr, dev):
    """Host float64 array -> cached (grow-only) device buffer on the current stream.  Pinned sources travel
    asynchronously at full PCIe rate; pageable ones are staged by the driver before the call returns."""
    torch = _torch()
    a = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64))
    key = (name, str(dev))
    buf = _dev_cache.get(key)
    if buf is None or buf.numel() < a.numel():
        buf = torch.empty(max(1, a.numel()), dtype=torch.float64, device=dev)
        _dev_cache[key] = buf
    view = buf[: a.numel()]
    view.copy_(a, non_blocking=True)
    return view


def _download(values, best, dev, copy):
    """Device result -> pinned staging -> numpy (one stream sync).  ``copy=False`` returns views of the staging
    buffer, valid until the next sharded call in this process."""
    torch = _torch()
    n = values.numel() + best.numel()
    key = str(dev)
    pin = _pin_cache.get(key)
    if pin is None or pin.numel() < n:
        pin = torch.empty(n, dtype=torch.float64).pin_memory()
        _pin_cache[key] = pin
    pin[: values.numel()].copy_(values, non_blocking=True)
    pin[values.numel(): n].copy_(best.reshape(-1), non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    out = pin[:n].numpy()
    vals, b = out[: values.numel()], out[values.numel():].reshape(best.shape)
    return (vals.copy(), b.copy()) if copy else (vals, b)


def pdm_sharded_p2p(t, x, periods, nb, nc, device=None, group=None, root=None, copy=True):
    """numpy in / numpy out wrapper of :func:`pdm_sharded_p2p_torch` (same return as ``pdm_sharded``).

    Every rank uploads the inputs (cached device buffers), evaluates its slice of the period grid and the fused
    gather leaves the whole theta array on every GPU.  ``root=None``: every rank downloads it; ``root=r``: only
    rank r does (the others return ``(None, index, value)``)."""
    torch = _torch()
    ctx = _ffi.default_context(device)
    dev = torch.device("cuda", ctx.device)
    _, rank, _ = _dist_info(group)
    tt = _upload_cached("pdm_t", t, dev)
    xx = _upload_cached("pdm_x", x, dev)
    pp = _upload_cached("pdm_p", periods, dev)
    theta, best = pdm_sharded_p2p_torch(tt, xx, pp, nb, nc, ctx=ctx, group=group)
    if root is not None and rank != root:
        b = best.cpu().numpy()
        idx, val = reduce_best(b[:, 0], b[:, 1].astype(np.int64), -1)
        return None, idx, val
    th, b = _download(theta, best, dev, copy)
    idx, val = reduce_best(b[:, 0], b[:, 1].astype(np.int64), -1)
    return th, idx, val


def gls_sharded_p2p(t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, device=None, group=None, root=None,
                    copy=True):
    """numpy in / numpy out wrapper of :func:`gls_sharded_p2p_torch` (same return as ``gls_sharded``); ``root`` and
    ``copy`` as in :func:`pdm_sharded_p2p`."""
    torch = _torch()
    ctx = _ffi.default_context(device)
    dev = torch.device("cuda", ctx.device)
    _, rank, _ = _dist_info(group)
    tt = _upload_cached("gls_t", t, dev)
    yy = _upload_cached("gls_y", y, dev)
    ww = None if w is None else _upload_cached("gls_w", w, dev)
    power, best = gls_sharded_p2p_torch(tt, yy, ww, fmin, df, nf, fit_mean, psd_scale, ctx=ctx, group=group)
    if root is not None and rank != root:
        b = best.cpu().numpy()
        idx, val = reduce_best(b[:, 0], b[:, 1].astype(np.int64), +1)
        return None, idx, val
    p, b = _download(power, best, dev, copy)
    idx, val = reduce_best(b[:, 0], b[:, 1].astype(np.int64), +1)
    return p, idx, val
