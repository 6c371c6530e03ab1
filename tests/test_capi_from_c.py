"""The C ABI used from plain C (examples/capi_demo.c: gcc + libcudart, no Python / torch in the process): builds on
CPU (-m "not gpu": compile + link only), runs on the GPU and is checked bit for bit against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT
from util import shape_cloud

EXE = os.path.join(ROOT, "examples", "capi_demo")


def build_demo():
    from genpc_b200.csrc import build

    build.build()
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "capi_demo.c"), "-o", EXE, "-L", os.path.join(ROOT, "genpc_b200"),
           "-lgenpc_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "genpc_b200"),
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return EXE


def test_c_program_compiles_and_links_against_the_header():
    assert os.path.exists(build_demo())


@pytest.mark.gpu
def test_c_program_matches_oracle(cuda, tmp_path):
    exe = build_demo()
    a, b = shape_cloud(1, 1, 5000), shape_cloud(2, 1, 12345)
    fa, fb = tmp_path / "a.f32", tmp_path / "b.f32"
    a[0].tofile(fa)
    b[0].tofile(fb)
    out = subprocess.run([exe, str(fa), "5000", str(fb), "12345", str(tmp_path / "o")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    assert np.array_equal(np.fromfile(tmp_path / "o.dist1", np.float32), d1[0])
    assert np.array_equal(np.fromfile(tmp_path / "o.dist2", np.float32), d2[0])
    assert np.array_equal(np.fromfile(tmp_path / "o.idx1", np.int32), i1[0])
    assert np.array_equal(np.fromfile(tmp_path / "o.idx2", np.int32), i2[0])
    assert np.array_equal(np.fromfile(tmp_path / "o.fps", np.int32), oracle.fps(a, 64, 0)[0])
