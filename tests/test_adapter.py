"""The drop-in claim, compiled and run: oracle/make_adapter.py applies the adapter block of INTEGRATION.md section 2 to a scratch
copy of the reference's Terrain.cpp (lines 44-89, 113-115) and builds oracle/_ref_adapter/adapter_check from it, the
reference's other unmodified sources and librednoise_b200.so.  On the GPU the program runs the REAL constructor
Terrain::Terrain(depth, 500, 0.2f) (Core.cpp:65), whose field now comes from rn_terrain2d, and compares the whole Cell tree,
volcanoPos and the rand() stream position with the reference's own CPU path for the same seed."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def adapter(rnlib):
    import make_adapter
    exe = make_adapter.build()
    if exe is None:
        pytest.skip("oracle/_ref_adapter/adapter_check not built (reference tree absent)")
    return exe


def test_adapter_builds_from_integration_md_and_refuses_without_a_device(adapter):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([adapter, "8", "12345"], capture_output=True, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stderr


def test_adapter_anchors_match_the_reference():
    """The line numbers INTEGRATION.md cites are checked against the reference source whenever it is present."""
    import make_adapter
    ref = "/root/reference/renderEngine"
    if not os.path.isdir(ref):
        pytest.skip("reference tree absent")
    src = make_adapter.patched_source(ref)
    assert "rn_terrain2d(rn, &p" in src and "generateRecursive(NULL, Position(0,0,0)" in src
    assert "averageSquare(x, y, x+step, y+step);\r\n    for" not in src.split("Terrain::Terrain")[1].split("void Terrain::addVolcano")[0]


@pytest.mark.gpu
@pytest.mark.parametrize("depth,seed", [(5, 7), (8, 12345), (10, 777), (12, 12345)])
def test_adapted_constructor_builds_the_reference_cell_tree(adapter, depth, seed):
    r = subprocess.run([adapter, str(depth), str(seed)], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "IDENTICAL" in r.stdout
