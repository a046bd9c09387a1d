# SASS evidence of one kernel of the shipped library: resource usage, static opcode histogram, and the lines around the
# mnemonics that matter (packed FP32, 256-bit stores, cp.async, constant-memory tables).
#   python profiles/micro/sass_excerpt.py KERNEL_SUBSTRING MNEMONIC[,MNEMONIC...] > profiles/r02_<kernel>_sass.txt
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "rednoise_b200", "librednoise_b200.so")
kname, keys = sys.argv[1], sys.argv[2].split(",")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout.splitlines()
body, inside, title = [], False, ""
for ln in sass:
    if "Function :" in ln:
        if inside: break
        inside = kname in ln
        title = ln.strip()
        continue
    if inside and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): body.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
print("# cuobjdump -sass / -res-usage of rednoise_b200/librednoise_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false)")
print("#", title)
for k, ln in enumerate(res):
    if kname in ln:
        print("#", res[k + 1].strip() if k + 1 < len(res) else "")
        break
ops = collections.Counter(re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?", "", ln).split()[0].rstrip(";") for ln in body)
print("# %d SASS instructions; static opcode histogram:" % len(body))
print("#  " + ", ".join("%s %d" % kv for kv in ops.most_common(40)))
for key in keys:
    hits = [k for k, ln in enumerate(body) if key in ln]
    print("\n# ---- %s: %d instructions; context of the first occurrence(s) ----" % (key, len(hits)))
    shown = -100
    for h in hits[:2]:
        if h - shown < 12: continue
        for ln in body[max(0, h - 6):h + 7]: print(ln)
        print("        ...")
        shown = h
