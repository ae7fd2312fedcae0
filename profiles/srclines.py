"""Per source line: warp instructions executed and stall samples of each kernel in an ncu report.
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python profiles/srclines.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
def num(s):
    try: return int(s)
    except ValueError: return 0
kern, path, hdr = None, None, None
data = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = r[1].split("(")[0]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if r[0] and r[0].isdigit():
        d = data.setdefault(kern, collections.OrderedDict())
        k = (path, int(r[0]))
        a = d.setdefault(k, [r[1], 0, 0])
        a[1] += num(r[hdr["Instructions Executed"]])
        a[2] += num(r[hdr["Warp Stall Sampling (All Samples)"]])
for kern, d in data.items():
    ti = sum(a[1] for a in d.values()); ts = sum(a[2] for a in d.values())
    print(f"\n## {kern}: {ti/1e6:.1f} M warp instructions, {ts} stall samples")
    for (p, ln), a in sorted(d.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{p}:{ln:4d} instr {100*a[1]/max(ti,1):5.1f}%  samples {100*a[2]/max(ts,1):5.1f}%  {a[0].strip()[:110]}")
