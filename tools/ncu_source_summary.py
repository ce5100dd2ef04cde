"""Summarise an `ncu --page source --csv` dump of one kernel: instruction mix, average active
lanes, stall reasons, hottest instructions.  usage: ncu_source_summary.py file.csv [ntop]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[h - 1][:2] if h else "")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
ci = {k: i for i, k in enumerate(hdr)}
I = lambda r, k: int(float(r[ci[k]] or 0))
tot_inst = sum(I(r, "Instructions Executed") for r in data)
tot_thr = sum(I(r, "Thread Instructions Executed") for r in data)
tot_samp = sum(I(r, "# Samples") for r in data)
print(f"SASS instrs {len(data)}  warp-inst {tot_inst}  thread-inst {tot_thr}  avg active lanes {tot_thr / max(tot_inst, 1):.2f}  samples {tot_samp}")
cls, clsT, clsS = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    toks = r[ci["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    cls[op] += I(r, "Instructions Executed")
    clsT[op] += I(r, "Thread Instructions Executed")
    clsS[op] += I(r, "# Samples")
print("op          warp-inst   share  active  samples-share")
for op, n in cls.most_common(22):
    print(f"{op:10s} {n:10d} {n / tot_inst:6.3f} {clsT[op] / max(n, 1):6.1f} {clsS[op] / max(tot_samp, 1):8.3f}")
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(I(r, k) for r in data) for k in st}
print("stalls:", [(k, v, round(v / max(tot_samp, 1), 3)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
print("hottest instructions: samples warp-inst active  source")
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:ntop]:
    print(f"{I(r, '# Samples'):7d} {I(r, 'Instructions Executed'):9d} {r[ci['Avg. Threads Executed']]:>5s}  {r[ci['Source']][:100]}")
