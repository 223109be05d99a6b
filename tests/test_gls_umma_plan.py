"""Host logic of the tensor-core GLS kernels (csrc/gls_umma.cu::gls_umma_plan, exported as pdc_debug_umma_plan): how a
call is cut into tiles (128 or 256 fine indices x coarse blocks), sample splits and accumulation runs, and which kernel
takes it.  Pure host arithmetic: runs without a GPU."""
import itertools

import pytest

from periodicity_b200 import _ffi

GIB = 1 << 30


def check_invariants(B, nf, nmax, p, sm=148):
    fine = p["fine"]
    assert fine == (256 if p["path"] == 3 else 128)
    assert p["nC"] == -(-nf // fine)
    for nt, cpt, rows, rnd in ((p["nt1"], p["cpt1"], 4, 4), (p["nt2"], p["cpt2"], 2, 16 if p["path"] == 3 else 8)):
        assert nt >= 1 and cpt % rnd == 0
        assert rows * cpt <= 256 and (rows * cpt) % 16 == 0            # the MMA's N
        assert nt * cpt >= p["nC"] > (nt - 1) * cpt                    # the tiles cover every coarse block, none is empty
    assert p["nsplit"] >= 1 and -(-nmax // p["nsplit"]) <= 16384        # FP32 masters: at most 16384 samples per job
    assert p["chunk_stages"] in (4, 8, 16)
    assert p["jobs"] == B * (p["nt1"] + p["nt2"]) * p["nsplit"]
    assert 0 <= p["fine_bytes"] <= 16 * GIB + (1 << 22)
    assert (p["fine_bytes"] == 0) == (p["path"] == 1)
    if p["path"] == 3:
        assert B == 1 and nf >= 16384
    if p["path"] in (2, 3):
        assert B == 1
        per = -(-(-(-nmax // p["nsplit"])) // 64) * 64
        stages = per * p["nsplit"] // 16 + 16
        assert p["fine_bytes"] == stages * (65536 if p["path"] == 3 else 32768)


@pytest.mark.parametrize("B,nf,nmax", list(itertools.product(
    [1, 3, 256, 10000], [256, 300, 1600, 4096, 8321, 10000, 16384, 20000, 100000, 1234567, 10 ** 7],
    [9, 777, 4097, 20000, 65000, 10 ** 6])))
def test_plan_invariants(B, nf, nmax):
    check_invariants(B, nf, nmax, _ffi.umma_plan(B, nf, nmax))


def test_named_configs():
    c2 = _ffi.umma_plan(1, 100_000, 65_000)
    assert (c2["path"], c2["nt1"], c2["cpt1"], c2["nt2"], c2["cpt2"]) == (3, 7, 56, 4, 112)
    assert c2["nsplit"] == 13 and c2["chunk_stages"] == 16              # 143 clusters on 74 pairs of SMs: two waves
    c5 = _ffi.umma_plan(1, 10 ** 7, 10 ** 6)
    assert c5["path"] == 3 and c5["nsplit"] == 62 and c5["fine_bytes"] < 5 * GIB
    c4 = _ffi.umma_plan(10_000, 10_000, 20_000)
    assert c4["path"] == 1 and (c4["nt1"], c4["cpt1"], c4["nt2"], c4["cpt2"], c4["nsplit"]) == (2, 40, 1, 80, 2)
    shard = _ffi.umma_plan(1, 10 ** 7 // 8, 10 ** 6)                     # C5 on one of eight GPUs
    assert shard["path"] == 3 and shard["nsplit"] == 62


def test_knobs():
    assert _ffi.umma_plan(1, 100_000, 65_000, cg2=0)["path"] == 2       # one CTA per tile, fine operand precomputed
    assert _ffi.umma_plan(1, 100_000, 65_000, fine=0)["path"] == 1      # fine operand in the kernel: no pair kernel either
    assert _ffi.umma_plan(1, 8321, 4097, cg2=1)["path"] == 3            # forced from 4096 frequencies on
    assert _ffi.umma_plan(1, 4000, 4097, cg2=1)["path"] in (1, 2)
    assert _ffi.umma_plan(1, 1600, 3000, fine=1)["path"] == 2
    p = _ffi.umma_plan(1, 100_000, 65_000, nsplit=7, chunk=5)
    assert p["nsplit"] == 7 and p["chunk_stages"] == 6                  # run length rounded up to an even number of stages
    assert _ffi.umma_plan(1, 100_000, 65_000, sm_count=32)["nsplit"] != 13   # the splits follow the SM count


def test_scratch_cap_sends_very_long_curves_to_the_in_kernel_operand():
    p = _ffi.umma_plan(1, 10 ** 6, 10 ** 8)                              # 1e8 samples: 400 GB of fine images would be needed
    assert p["path"] == 1 and p["fine_bytes"] == 0
    check_invariants(1, 10 ** 6, 10 ** 8, p)
    p = _ffi.umma_plan(1, 10 ** 6, 3 * 10 ** 6)                          # 12 GB on the pair kernel: still precomputed
    assert p["path"] == 3


def test_invalid_arguments():
    with pytest.raises(ValueError):
        _ffi.umma_plan(0, 1000, 1000)
    with pytest.raises(ValueError):
        _ffi.umma_plan(1, 1000, 1000, sm_count=1)
