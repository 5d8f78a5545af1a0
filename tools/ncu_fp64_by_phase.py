"""Development aid: executed FP64 instructions per phase from an ncu report (source page, CUDA + SASS view).
Each SASS row carries 'Instructions Executed' (warp level) and 'Predicated-On Thread Instructions Executed'; the CUDA
line it belongs to maps it to a phase by (file, function-range) below.
usage: ncu_fp64_by_phase.py report.ncu-rep items steps nodes algorithmic_flops_per_node_step"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
items, steps, nodes, alg = (float(x) for x in sys.argv[2:6])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))

# phase of a CUDA source line: function bodies located by scanning the sources for their signatures
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "csrc")
MARK = [("hadi_phases.cuh", "void hadi_phase_tables", "setup"), ("hadi_phases.cuh", "void hadi_phase_factor", "setup"),
        ("hadi_phases.cuh", "void hadi_phase_div1", "dividend"), ("hadi_phases.cuh", "void hadi_phase_div2", "dividend"),
        ("hadi_phases.cuh", "void hadi_phase_div3", "dividend"),
        ("hadi_phases.cuh", "void hadi_phase_explicit", "E"), ("hadi_phases.cuh", "void hadi_phase_solve_a1", "S1"),
        ("hadi_phases.cuh", "void hadi_phase_rhs2", "S2"), ("hadi_phases.cuh", "void hadi_phase_solve_a2", "S2"),
        ("hadi_phases.cuh", "void hadi_phase_project", "P"), ("hadi_phases.cuh", "double hadi_rcp_prep", "setup"),
        ("hadi_phases.cuh", "double hadi_div", "division (S1+S2+P)"),
        ("hadi_phases_fast.cuh", "void hadi_relay_factor", "setup"), ("hadi_phases_fast.cuh", "void hadi_tm2_bwd_quad", "S1"),
        ("hadi_phases_fast.cuh", "void hadi_relay_stage", "S1"), ("hadi_phases_fast.cuh", "void hadi_relay_solve_a1", "S1"),
        ("hadi_phases_fast.cuh", "void hadi_fast_solve_a2", "S2")]
ranges = {}
for fn in ("hadi_phases.cuh", "hadi_phases_fast.cuh", "hadi_kernel.cu"):
    src = open(os.path.join(CS, fn)).read().splitlines()
    starts = []
    for f, sig, ph in MARK:
        if f != fn:
            continue
        for k, ln in enumerate(src):
            if sig in ln and "(" in ln:
                starts.append((k + 1, ph))
                break
    # every function start (to end a range): lines beginning a definition at column 0 with HADI_HD / template / __device__
    defs = sorted({k + 1 for k, ln in enumerate(src) if re.match(r"^(HADI_HD|template|__device__|__global__|static|int |void |struct )", ln)})
    for st, ph in starts:
        nxt = [d for d in defs if d > st + 1]
        ranges.setdefault(fn, []).append((st - 2, (nxt[0] - 1) if nxt else len(src), ph))


def phase_of(fn, line):
    for a, b, ph in ranges.get(fn, []):
        if a <= line <= b:
            return ph
    return "other (" + fn + ")"


cur_file, cur_line, hdr = None, None, None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "":
        try:
            cur_line = int(r[0])
        except ValueError:
            pass
        continue
    sass = r[3]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    if not m:
        continue
    op = m.group(2)
    if not (op.startswith("DADD") or op.startswith("DMUL") or op.startswith("DFMA") or op.startswith("DSETP") or op.startswith("MUFU.RCP64")):
        continue
    d = dict(zip(hdr, r))
    warp = float(d["Instructions Executed"] or 0)
    thr = float(d["Predicated-On Thread Instructions Executed"] or 0)
    ph = phase_of(cur_file, cur_line)
    a = agg.setdefault(ph, {})
    key = op.split(".")[0]
    w, t = a.get(key, (0.0, 0.0))
    a[key] = (w + warp, t + thr)
unit = items * steps * nodes
print("executed FP64 instructions per grid-point-step (thread level, predicated on) and warp-level pipe issues per item-step;")
print("algorithmic flops per grid-point-step: %g (DFMA counts two)" % alg)
tot_thr = tot_warp = 0.0
for ph in sorted(agg):
    a = agg[ph]
    thr = sum(t * (2 if k == "DFMA" else 1) for k, (w, t) in a.items() if k in ("DADD", "DMUL", "DFMA"))
    warp = sum(w for k, (w, t) in a.items())
    tot_thr += thr
    tot_warp += warp
    det = " ".join("%s=%.1f" % (k, t / unit) for k, (w, t) in sorted(a.items()))
    print("%-22s flops/node-step %6.2f   warp-level FP64-pipe instructions per item-step %8.0f   [%s]" % (ph, thr / unit, warp / (items * steps), det))
print("algorithmic split (SURVEY 8d, American Douglas): E 45 (A0 18, A1 5, A2 9, Y0 6, lambda 1, A1-RHS 6), S1 5, S2 15 (RHS 6, solve 9), P 7")
print("%-22s flops/node-step %6.2f   warp-level FP64-pipe instructions per item-step %8.0f" % ("total", tot_thr / unit, tot_warp / (items * steps)))
print("executed / algorithmic = %.3f" % (tot_thr / unit / alg))
