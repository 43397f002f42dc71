"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: instructions executed and
stall samples per CUDA source line.  Usage: ncu_lines.py <src.csv> <nvdisasm -g -c output> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, disasm = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ii, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
counts = [(r[si].strip(), int(r[ii]), int(r[ss])) for r in rows[2:] if len(r) > ii and r[ii].isdigit()]
# nvdisasm: lines like  //## File "x.cu", line 123   followed by instructions  /*0010*/ OPC ...
cur = ("?", 0)
line_of = []
for ln in open(disasm):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln) and not ln.strip().startswith("//"):
        line_of.append(cur)
n = min(len(counts), len(line_of))
agg = defaultdict(lambda: [0, 0])
for (sass, c, s), key in zip(counts[:n], line_of[:n]):
    agg[key][0] += c
    agg[key][1] += s
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print(f"sass rows {len(counts)} disasm instrs {len(line_of)} total instr {tot} samples {tots}")
srcs = {}
for (f, l), (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(f"/root/repo/robust-object-detection_b200/csrc/{f}").read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{c:>11} {100 * c / tot:5.1f}%  samp {100 * s / max(tots, 1):5.1f}%  {f}:{l:<4} {text}")
