"""The multi-GPU C ABI (rn_create_multi / rn_terrain2d_sharded / rn_gather / memory helpers) driven by a plain C++ program
(tests/cpp/test_multi_gpu.cpp: no Python, no torch, links librednoise_b200.so only).  CPU: it builds, links and reports "no
device" without one.  GPU: slabs on 2 devices (two parts on one device when the box has a single GPU) assembled by NCCL
all-gather, by peer copies and straight into host memory == the single-device field, which matches the reference digest."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cpp", "test_multi_gpu.cpp")
EXE = os.path.join(HERE, "cpp", "test_multi_gpu")


@pytest.fixture(scope="module")
def exe(rnlib):
    lib = rnlib.LIB_PATH
    deps = [SRC, os.path.join(ROOT, "include", "rednoise.h"), lib]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-o", EXE, SRC, "-L" + os.path.dirname(lib), "-lrednoise_b200",
                        "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return EXE


def test_cpp_program_links_and_refuses_without_a_device(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 77, r.stdout + r.stderr
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [2, 3])
def test_cpp_sharded_field_equals_single_device(exe, ndev):
    gold = json.load(open(os.path.join(HERE, "golden", "terrain2d_golden.json")))
    e = [f for f in gold["fields"] if f["depth"] == 12 and f["seed"] == 12345][0]
    r = subprocess.run([exe, str(ndev), "12", e["fnv64_post_volcano"]], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "DIFFERS" not in r.stdout
