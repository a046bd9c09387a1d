#!/usr/bin/env python
"""oracle/make_adapter.py -- TEST INFRASTRUCTURE ONLY.

Proves the drop-in claim of INTEGRATION.md section 2 by compiling it: the adapter block is cut out of INTEGRATION.md, applied to a
SCRATCH copy of the reference's renderEngine/source/Terrain.cpp (outside the repository; the reference sources are never copied
into it), and built with the reference's other unmodified sources, oracle/ref_harness.cpp and librednoise_b200.so into
oracle/_ref_adapter/adapter_check (git-ignored, travels to the GPU box as a built binary like oracle/_ref).

    python oracle/make_adapter.py [--ref /root/reference/renderEngine]

Does nothing (keeps a prebuilt binary) when the reference tree is absent.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "_ref_adapter")
EXE = os.path.join(OUT_DIR, "adapter_check")
LIBDIR = os.path.join(ROOT, "rednoise_b200")


def blocks():
    """The two C++ blocks of INTEGRATION.md between <!-- adapter:NAME --> markers."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    out = {}
    for name in ("body", "cleanup"):
        m = re.search(r"<!-- adapter:%s -->\s*```cpp\n(.*?)```" % name, text, re.S)
        if not m:
            raise SystemExit(f"INTEGRATION.md: adapter block '{name}' not found")
        out[name] = m.group(1)
    return out


def patched_source(ref):
    src = open(os.path.join(ref, "source", "Terrain.cpp"), encoding="utf-8-sig").read().split("\n")
    # the anchors the patch relies on (1-based reference line numbers)
    expect = {44: "maxWidth = 1<<depth;", 58: "static unsigned int randomseed = (unsigned int)time(NULL);", 88: "addVolcano(size);",
              89: "volcanoPos *= size;", 112: "delete[] cells;", 115: "delete baseArray;"}
    for ln, what in expect.items():
        if what not in src[ln - 1]:
            raise SystemExit(f"Terrain.cpp:{ln} is not '{what}' -- the reference changed, adapt INTEGRATION.md")
    b = blocks()
    new = src[:43] + b["body"].rstrip("\n").split("\n") + src[89:112] + b["cleanup"].rstrip("\n").split("\n") + src[115:]
    return '#include "rednoise.h"\n' + "\n".join(new)


def build(ref="/root/reference/renderEngine", verbose=False):
    if not os.path.isdir(os.path.join(ref, "source")):
        print(f"reference tree {ref} absent: keeping prebuilt {EXE}" if os.path.exists(EXE) else f"reference tree {ref} absent: adapter not built")
        return EXE if os.path.exists(EXE) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="rn_adapter_") as tmp:
        patched = os.path.join(tmp, "Terrain_adapted.cpp")
        open(patched, "w").write(patched_source(ref))
        others = [os.path.join(ref, "source", f) for f in ("Cell.cpp", "OrientableCell.cpp", "ProceduralExample.cpp", "Vector3.cpp", "Matrix33.cpp",
                                                            "Rock.cpp", "Cloud.cpp", "PhysicalCell.cpp", "DynamicCell.cpp")]
        cmd = ["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-include", os.path.join(HERE, "adapter_shim.h"),
               "-I" + os.path.join(ref, "include"), "-I" + os.path.join(ROOT, "include"), "-I" + HERE, "-o", EXE,
               os.path.join(HERE, "adapter_main.cpp"), patched] + others + ["-L" + LIBDIR, "-lrednoise_b200", "-Wl,-rpath,$ORIGIN/../../rednoise_b200"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return EXE


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--ref") + 1] if "--ref" in sys.argv else "/root/reference/renderEngine"
    print(build(ref, verbose="-v" in sys.argv))
