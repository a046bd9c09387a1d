# Executed instructions per SOURCE LINE of one kernel: joins the per-instruction counts of an .ncu-rep (--set full) with the
# line table of the same build of the library (nvdisasm -g on the embedded cubin; the kernels are built with -lineinfo).
#   python profiles/micro/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [samples_per_launch] [top_n]
import csv, os, re, subprocess, sys, tempfile, collections, glob
rep, kname = sys.argv[1], sys.argv[2]
samples = float(sys.argv[3]) if len(sys.argv) > 3 else None
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "rednoise_b200", "librednoise_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for cub in sorted(glob.glob(os.path.join(tmp, "*.cubin")), key=len):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur, inside, acc = None, False, []
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            if inside and acc: break
            inside = kname in ln
            cur = None
            continue
        if not inside: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: acc.append((int(m.group(1), 16), cur, m.group(2).strip()))
    if acc:
        lines = acc
        break
if not lines: sys.exit("kernel not found in the library's cubins")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iA, iE, iStall = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [r for r in rows[2:] if len(r) > iE]
if len(data) != len(lines): print("warning: %d instructions in the report, %d in the library (different build?)" % (len(data), len(lines)))
per = collections.Counter(); st = collections.Counter(); tot = 0
for r, (off, cur, text) in zip(data, lines):
    n = int(r[iE]); per[cur] += n; st[cur] += int(r[iStall] or 0); tot += n
ts = sum(st.values()) or 1
print("kernel:", rows[0][1]); print("warp instructions executed:", tot, ("(%.1f thread-level per sample at 32 lanes)" % (tot * 32 / samples)) if samples else "")
byfile = collections.Counter()
for (k, n) in per.items(): byfile[k[0] if k else None] += n
print("by file:", dict(byfile))
src_cache = {}
def src(k):
    if not k: return ""
    f = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", k[0])
    if f not in src_cache: src_cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
    L = src_cache[f]
    return L[k[1] - 1].strip()[:110] if 0 < k[1] <= len(L) else ""
print("%-22s %12s %6s %6s  %s" % ("file:line", "warp inst", "share", "stall", "source"))
for k, n in per.most_common(topn):
    print("%-22s %12d %5.1f%% %5.1f%%  %s" % ("%s:%d" % k if k else "?", n, 100.0 * n / tot, 100.0 * st[k] / ts, src(k)))
