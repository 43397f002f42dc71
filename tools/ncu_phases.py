"""Per-phase (split at BAR.SYNC) and per-source-line breakdown of one kernel from an ncu SASS page.
Usage: ncu_phases.py <src.csv> <nvdisasm -g -c output> <kernel mangled-name substring> <output bytes>"""
import csv
import re
import sys
from collections import defaultdict

src_csv, disasm, kname, nbytes = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ii, ss, si = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
counts = [(r[si].strip(), int(r[ii]), int(r[ss])) for r in rows[2:] if len(r) > ii and r[ii].isdigit()]
lines_txt = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines_txt) if l.startswith(".text.") and kname in l)
end = next((i for i in range(start + 1, len(lines_txt)) if lines_txt[i].lstrip().startswith(".section")), len(lines_txt))
cur, line_of = ("?", 0), []
for ln in lines_txt[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln) and not ln.strip().startswith("//"):
        line_of.append(cur)
print("sass rows", len(counts), "disasm instrs", len(line_of))
phase, agg, ops, lines = 0, defaultdict(lambda: [0, 0]), defaultdict(lambda: defaultdict(int)), defaultdict(lambda: [0, 0])
for (sass, c, s), k in zip(counts, line_of):
    if "BAR.SYNC" in sass:
        phase += 1
    agg[phase][0] += c
    agg[phase][1] += s
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    ops[phase][op] += c
    lines[(phase,) + k][0] += c
    lines[(phase,) + k][1] += s
tot = sum(v[0] for v in agg.values())
tots = max(1, sum(v[1] for v in agg.values()))
for p, v in agg.items():
    print(f"phase {p}: {100 * v[0] / tot:.1f}% instr  {100 * v[1] / tots:.1f}% samples  thread-instr/byte {v[0] * 32 / nbytes:.2f}")
    print("    ", [(k, round(c * 32 / nbytes, 2)) for k, c in sorted(ops[p].items(), key=lambda kv: -kv[1])[:18]])
src = {}
for k, (c, s) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:24]:
    f = k[1]
    if f not in src:
        try:
            src[f] = open(f"/root/repo/robust-object-detection_b200/csrc/{f}").read().split("\n")
        except Exception:
            src[f] = []
    t = src[f][k[2] - 1].strip()[:84] if 0 < k[2] <= len(src[f]) else ""
    print(f"ph{k[0]} {100 * c / tot:5.1f}%ins {100 * s / tots:5.1f}%smp {f}:{k[2]} {t}")
