# Dynamic opcode mix of one kernel from an .ncu-rep captured with --set full --import-source on:
#   python profiles/micro/ncu_opmix.py REPORT.ncu-rep [samples_per_launch]
# Prints warp-instructions executed per opcode (and per output sample when the sample count is given) and the source lines
# (needs -lineinfo) that execute the most instructions.
import csv, re, subprocess, sys, collections
rep = sys.argv[1]
samples = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iT, iW, iStall = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter(); thr = collections.Counter(); stall = collections.Counter(); wf = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m: continue
    op = m.group(2)
    if op in ("LDS", "STS", "STG", "LDG", "LDGSTS", "MUFU"):
        m2 = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS]); op = m2.group(2)
    n = int(r[iE]); ops[op] += n; thr[op] += int(r[iT]); stall[op] += int(r[iStall] or 0); wf[op] += int(r[iW] or 0); tot += n
print("kernel:", rows[0][1])
print("warp instructions executed: %d%s" % (tot, "  (%.1f thread instructions per sample)" % (sum(thr.values()) / samples) if samples else ""))
print("%-18s %14s %7s %9s %9s" % ("opcode", "warp inst", "share", "stall smp", "sh wavefr"))
ts = sum(stall.values()) or 1
for op, n in ops.most_common(40):
    print("%-18s %14d %6.1f%% %8.1f%% %9d" % (op, n, 100.0 * n / tot, 100.0 * stall[op] / ts, wf[op]))
