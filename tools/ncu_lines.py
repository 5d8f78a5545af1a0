"""Development aid: aggregate an ncu report's per-instruction samples by source line.
usage: ncu_lines.py report.ncu-rep [top] [--active]   (--active drops barrier-stall samples)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
active = "--active" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = {h: r[i] for i, h in enumerate(hdr)}
    st = {h: int(d[h]) for h in d if h.startswith("stall_") and "Not" not in h and d[h] not in ("", "-")}
    samples = int(d["# Samples"] or 0)
    if active:
        samples -= st.get("stall_barrier", 0)
    agg[(cur, ln)] = (samples, int(d["Instructions Executed"] or 0), r[1], st)
tot = sum(v[0] for v in agg.values())
print("total samples", tot, "(barrier stalls excluded)" if active else "")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = {a: b for a, b in v[3].items() if b > 0 and not (active and a == "stall_barrier")}
    tp = sorted(st.items(), key=lambda x: -x[1])[:3]
    tp = " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in tp)
    print(f"{k[0][:13]:13s}:{k[1]:4d} {100*v[0]/max(tot,1):5.1f}% inst={v[1]:9d} [{tp}] | {v[2].strip()[:72]}")
