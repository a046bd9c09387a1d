"""Generates tests/golden/terrain2d_golden.json from the REFERENCE'S OWN sources.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
It compiles oracle/_ref/libredref.so (oracle/Makefile: the unmodified reference
Terrain.cpp etc. + the MSVC-LCG rand() stand-in) and records, for a grid of (depth,
seed), the FNV-1a-64 digest of the raw field before and after addVolcano, the rand()
call count, sampled rows as raw float bits, and the LCG known answers.  The fixture is
what the `-m "not gpu"` and `-m gpu` tests compare against on the GPU box, where
/root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402


def fnv64(a: np.ndarray) -> str:
    return oracle_lib.load_oracle().fnv64(a)


def main():
    oracle_lib.build_ref()
    ref = oracle_lib.load_ref()
    assert ref is not None, "reference tree not available"
    cases = [(3, 12345), (4, 12345), (5, 12345), (6, 12345), (7, 4242), (8, 12345), (8, 777), (8, 0), (8, 0xFFFFFFFF),
             (9, 31337), (10, 12345), (10, 12346), (11, 2024), (12, 12345)]
    out = {"generator": "oracle/_ref/libredref.so (reference Terrain.cpp, MSVC LCG, g++ -O2 -ffp-contract=off)",
           "params": {"amplitude": 0.2, "size": 500.0}, "hash": "FNV-1a-64 of raw little-endian field bytes, i-major, 16 B/elem",
           "fields": [], "lcg_kat": {}, "rows": [], "small_fields": []}
    for depth, seed in cases:
        pre, calls = ref.terrain2d(depth, seed, volcano=0)
        post, _ = ref.terrain2d(depth, seed, volcano=1)
        y = post[..., 1].astype(np.float64)
        out["fields"].append({
            "depth": depth, "seed": seed, "rand_calls": int(calls), "samples": int(pre.shape[0] * pre.shape[1]),
            "fnv64_pre_volcano": fnv64(pre), "fnv64_post_volcano": fnv64(post),
            "post_mean_y": float(y.mean()), "post_min_y": float(y.min()), "post_max_y": float(y.max()),
        })
        if depth <= 4:
            out["small_fields"].append({"depth": depth, "seed": seed,
                                        "pre_volcano_bits": pre.view(np.uint32).reshape(-1).tolist(),
                                        "post_volcano_bits": post.view(np.uint32).reshape(-1).tolist()})
        if depth in (8, 10):
            W = 1 << depth
            for i in (0, 1, W // 2, W - 1, W):
                out["rows"].append({"depth": depth, "seed": seed, "row": i, "volcano": 1,
                                    "bits": post[i].view(np.uint32).reshape(-1).tolist()})
    # Cell quadtree (Terrain::generateRecursive run by the reference itself): digest of the rn_cell array
    out["cells"] = []
    for depth, seed in [(3, 12345), (5, 7), (8, 12345), (10, 12345), (11, 2024)]:
        _, cells = ref.cells(depth, seed)
        out["cells"].append({"depth": depth, "seed": seed, "count": int(len(cells)), "fnv64": fnv64(cells.view(np.uint8)),
                             "first": cells[:5].view(np.uint32).reshape(-1).tolist()})
    # Rock (Rock::generateBaseRecursive on six faces + Rock::generateRecursive, run by the reference itself)
    out["rocks"] = []
    for depth, state in [(1, 12345), (2, 12345), (3, 7), (4, 12345), (5, 99), (6, 12345), (7, 12345), (7, 0xDEADBEEF), (8, 4242), (9, 12345), (10, 12345)]:
        faces, cells, draws, state_after = ref.rock(state, depth, 0.3)
        e = {"depth": depth, "state": state, "amplitude": 0.3, "rand_calls": draws, "state_after": state_after,
             "fnv64_faces": fnv64(faces), "fnv64_cells": fnv64(cells.view(np.uint8))}
        if depth <= 2:
            e["faces_bits"] = faces.view(np.uint32).reshape(-1).tolist()
            e["cells_bits"] = cells.view(np.uint32).reshape(-1).tolist()
        out["rocks"].append(e)
    # Cloud (Cloud::generateBaseRecursive on two sheets + Cloud::generateRecursive, run by the reference itself)
    out["clouds"] = []
    for depth, state in [(2, 12345), (3, 7), (4, 12345), (5, 99), (7, 12345), (7, 0xDEADBEEF), (8, 4242), (10, 12345)]:
        faces, cells, draws, state_after = ref.cloud(state, depth, 0.3)
        e = {"depth": depth, "state": state, "amplitude": 0.3, "rand_calls": draws, "state_after": state_after,
             "fnv64_faces": fnv64(faces), "fnv64_cells": fnv64(cells.view(np.uint8))}
        if depth <= 2:
            e["faces_bits"] = faces.view(np.uint32).reshape(-1).tolist()
            e["cells_bits"] = cells.view(np.uint32).reshape(-1).tolist()
        out["clouds"].append(e)
    for seed in (1, 12345, 0, 0xFFFFFFFF):
        out["lcg_kat"][str(seed)] = ref.lcg_draws(seed, 16).tolist()
    path = os.path.join(HERE, "terrain2d_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")
    for e in out["fields"]:
        print(e["depth"], e["seed"], e["rand_calls"], e["fnv64_pre_volcano"], e["fnv64_post_volcano"])


if __name__ == "__main__":
    main()
