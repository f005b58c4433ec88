"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: python scratch/ncu_src_lines.py export.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or r[0] == "": continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def f(name):
        try: return float(r[ix[name]])
        except Exception: return 0.0
    key = (cur_file, line)
    a = agg.setdefault(key, dict(src=r[1].strip(), smp=0.0, inst=0.0, st=collections.Counter()))
    a["smp"] += f("# Samples"); a["inst"] += f("Instructions Executed")
    for k in ("stall_long_sb", "stall_wait", "stall_barrier", "stall_short_sb", "stall_no_inst", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_math", "stall_sleep", "stall_membar", "stall_tex", "stall_dispatch"):
        if k in ix: a["st"][k[6:]] += f(k)
ts = sum(a["smp"] for a in agg.values()) or 1; ti = sum(a["inst"] for a in agg.values()) or 1
print(f"total samples {ts:.0f}  warp instructions {ti:.0f}")
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = ", ".join(f"{k}:{int(v)}" for k, v in a["st"].most_common(3) if v)
    print(f"{fn}:{ln:<4d} {100*a['smp']/ts:5.1f}% smp {100*a['inst']/ti:5.1f}% inst  {a['src'][:80]:80s} [{st}]")
