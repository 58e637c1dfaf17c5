"""Summarise `ncu -i x.ncu-rep --page source --print-source cuda,sass --csv`: stall samples per
function of ocp_warp.h / file, and the hottest source lines."""
import collections, csv, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[3] if len(sys.argv) > 3 else "ocp_warp.h"
src = open(os.path.join(ROOT, "vboc_b200", "csrc", SRC)).read().split("\n")
meth, name = {}, "?"
for i, l in enumerate(src, 1):
    m = re.match(r"\s+VB_(?:DEV|HD)\s+[\w:<>,\s\*&]+?\s+(\w+)\(", l)
    if m:
        name = m.group(1)
    meth[i] = name
cur_file, hdr, col = None, None, None
by_fn = collections.defaultdict(collections.Counter)
lines = []
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; col = {}
        for i, n in enumerate(hdr):
            col.setdefault(n, i)
        continue
    if hdr is None or not r[0].isdigit():
        continue
    ln = int(r[0])
    def num(v):
        try:
            return int(float(v))
        except ValueError:
            return 0
    n = num(r[col["# Samples"]])
    ex = num(r[col["Instructions Executed"]])
    key = meth.get(ln, "?") if cur_file == SRC else cur_file
    by_fn[key]["samples"] += n
    by_fn[key]["inst"] += ex
    for st in hdr:
        if st.startswith("stall_") and "Not Issued" not in st and r[col[st]] not in ("", "-"):
            by_fn[key][st] += int(float(r[col[st]]))
    lines.append((n, ex, cur_file, ln, r[1].strip()[:90]))
tot = sum(c["samples"] for c in by_fn.values())
toti = sum(c["inst"] for c in by_fn.values())
print("total samples", tot, "instructions", toti)
for k, c in sorted(by_fn.items(), key=lambda kv: -kv[1]["samples"]):
    top = [(a.replace("stall_", ""), round(100 * b / max(c["samples"], 1))) for a, b in c.most_common(7) if a.startswith("stall_")][:4]
    print(f"{k:18s} samples {100*c['samples']/tot:5.1f}%  inst {100*c['inst']/toti:5.1f}%  {top}")
print()
for n, ex, f, ln, s in sorted(lines, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{100*n/tot:5.2f}% {100*ex/toti:5.2f}%i {f}:{ln}  {s}")
