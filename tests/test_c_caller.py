"""The boundary is a C ABI: the header must be plain C, and a C program must be able to drive the library without
Python or torch (tests/c/c_caller.c: multi-device ctx, GLS, PDM, error path)."""
import os
import subprocess

import pytest

from conftest import ROOT

INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "periodicity_b200", "lib")
SRC = os.path.join(ROOT, "tests", "c", "c_caller.c")


def test_header_is_plain_c_and_the_c_caller_links(tmp_path):
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                           os.path.join(INC, "periodicity_b200.h")])
    exe = tmp_path / "c_caller"
    subprocess.check_call(["gcc", "-std=c99", "-D_GNU_SOURCE", "-Wall", "-O1", "-I", INC, SRC, "-o", str(exe), "-L", LIBDIR,
                           "-lperiodicity_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    assert exe.exists()


@pytest.mark.gpu
def test_c_caller_runs_on_the_gpu(tmp_path):
    exe = tmp_path / "c_caller"
    subprocess.check_call(["gcc", "-std=c99", "-D_GNU_SOURCE", "-Wall", "-O1", "-I", INC, SRC, "-o", str(exe), "-L", LIBDIR,
                           "-lperiodicity_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    for devs in (["0"], ["0", "0", "0"]):
        out = subprocess.run([str(exe)] + devs, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "c_caller ok" in out.stdout and f"{len(devs)} device(s)" in out.stdout
