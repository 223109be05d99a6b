"""The integer planes of partial sums (gls_plane, gls_low, hist_plane) are shared by all calls on a ctx and are kept
ALL-ZERO between calls by the epilogues that read them; other entry points reuse some of the same buffers as plain
scratch (glsm.cu: gls_low) and mark them dirty.  Interleaving every entry point -- different shapes, growing and
shrinking buffers, failing calls in between -- must leave every result bit-identical to the first evaluation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_interleaved_entry_points_are_bit_reproducible():
    from periodicity_b200 import _ffi
    from test_gl_oracle import events
    ctx = _ffi.Context(0)
    rng = np.random.default_rng(91)
    t = np.sort(rng.uniform(0, 300.0, 9000))
    y = 5 + np.sin(2 * np.pi * t / 2.75) + rng.standard_normal(t.size)
    w = rng.uniform(0.5, 2.0, t.size)
    df = 1 / (t[-1] - t[0]) / 5
    periods = np.linspace(1.0, 9.0, 2500)
    ev = events(3000, 3.3, 92)
    off = np.array([0, 2000, 4500, 9000])
    Y = np.stack([y, y[::-1], 2 * y + 1])
    m = (y - y.max()) / (2 * (y.max() - y.min())) + 0.25

    calls = {
        "gls": lambda: ctx.gls(t, y, None, 0.5 * df, df, 12_345),
        "gls_w": lambda: ctx.gls(t, y, w, 0.5 * df, df, 7_001),
        "gls_dense": lambda: ctx.gls(t, y, None, 0.005 * df, 0.01 * df, 3_000),          # hundreds of sub-cycle bins
        "gls_small": lambda: ctx.gls(t[:200], y[:200], None, 0.5 * df, df, 300),
        "gls_freqs": lambda: ctx.gls_freqs(t, y, w, np.geomspace(1e-3, 30.0, 2_777)),
        "batch": lambda: ctx.gls_batch(t, y, w, off, np.full(3, 0.5 * df), np.full(3, df), 900),
        "multi": lambda: ctx.gls_multi(t, Y, None, 0.5 * df, df, 1_500),                   # uses gls_low as FP64 scratch
        "pdm": lambda: ctx.pdm(t, y, periods, 10, 2),
        "pdm_short": lambda: ctx.pdm(t[:1500], y[:1500], periods[:700], 5, 2),             # float2 (unpacked) columns
        "pdm_jd": lambda: ctx.pdm(t + 2_457_000.5, y, periods[:900], 10, 2),
        "aov": lambda: ctx.aov(t, y, periods, 8),
        "ce": lambda: ctx.ce(t, y, periods, 10, 5),
        "gl": lambda: ctx.gl(ev, periods[:600], 6, 4),
        "sl": lambda: ctx.stringlength(t[:1200], m[:1200], periods[:300]),
    }

    def snapshot(out):
        return [np.array(o, copy=True) for o in out if o is not None]

    first = {k: snapshot(f()) for k, f in calls.items()}
    order = list(calls)
    for rep in range(3):
        rng.shuffle(order)
        for k in order:
            if rep == 1 and k in ("pdm", "gls"):          # failing calls in between must not leave anything behind
                with pytest.raises(ValueError):
                    ctx.pdm(t, y, periods, 0, 2)
                with pytest.raises(ValueError):
                    ctx.gls(t, y, None, float("nan"), df, 100)
                with pytest.raises(ValueError):
                    ctx.ce(t, y, periods, 100_000, 100)
            got = snapshot(calls[k]())
            for a, b in zip(got, first[k]):
                np.testing.assert_array_equal(a, b, err_msg=f"{k} changed in repetition {rep}")
    ctx.close()
