"""Per-source-line stall samples from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_source_lines.py file.csv [kernel substring] [top n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
fn, path, cur = None, None, {}
out = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ii = hdr.index("Instructions Executed")
    elif r[0].isdigit() and fn and want in fn:
        key = (fn.split("(")[0], path, int(r[0]), r[1].strip()[:110])
        s, i = (int(r[si]) if r[si].isdigit() else 0), (int(r[ii]) if r[ii].isdigit() else 0)
        a = out.setdefault(key, [0, 0])
        a[0] += s
        a[1] += i
byk = {}
for (k, p, ln, src), (s, i) in out.items():
    byk.setdefault(k, []).append((s, i, p, ln, src))
for k, lst in byk.items():
    tot = sum(x[0] for x in lst)
    print(f"== {k}: {tot} samples")
    for s, i, p, ln, src in sorted(lst, reverse=True)[:top]:
        print(f"  {100.0 * s / max(tot, 1):5.1f}%  inst {i:>10}  {p}:{ln}  {src}")
