#!/usr/bin/env python
"""Turns the scratch captures in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed).

    python tools/make_profiles.py r01

Writes profiles/<round>_launches.txt (per-kernel share of one bench step from the ncu launch list),
profiles/<round>_<kernel>.txt (headline counters + the hottest SASS lines with their stall reasons) and copies the
bench JSON lines."""
import csv, json, os, shutil, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def kernel_summary(rep, out):
    raw = list(csv.reader(ncu(['-i', rep, '--page', 'raw', '--csv']).splitlines()))
    hdr, units = raw[0], raw[1]
    lines = [f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (one launch; ~40 replays, cold-cache)"]
    for r in raw[2:]:
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for i, h in enumerate(hdr):
            if any(h.startswith(w) for w in WANT) and r[i] not in ('', 'n/a'):
                lines.append(f"  {h} [{units[i]}] = {r[i]}")
    sass = list(csv.reader(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']).splitlines()))
    h = sass[1]
    isrc, ins, ismp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    rows = []
    for r in sass[2:]:
        try:
            rows.append((int(r[ismp]), int(r[ins]), r[isrc], r))
        except Exception:
            pass
    tot_s = sum(x[0] for x in rows) or 1
    tot_i = sum(x[1] for x in rows) or 1
    lines.append(f"\nSASS: {len(rows)} instructions, {tot_i} warp-instructions executed, {tot_s} stall samples")
    mnem = OrderedDict()
    for x in rows:
        m = x[2].replace('@', ' ').split()
        m = [t for t in m if not t.startswith('P') and not t.startswith('!P') and not t.startswith('UP') and not t.startswith('!UP')] or ['?']
        key = m[0].split('.')[0]
        mnem[key] = mnem.get(key, 0) + x[1]
    top_m = sorted(mnem.items(), key=lambda kv: -kv[1])[:14]
    lines.append("instruction mix (share of executed warp-instructions): " + ", ".join(f"{k} {v / tot_i * 100:.1f}%" for k, v in top_m))
    proof = [k for k in mnem if k.startswith(('UTC', 'LDTM', 'STTM', 'UTMA', 'UBLKCP', 'SYNCS', 'HMMA'))]
    lines.append("Blackwell mnemonics present: " + (", ".join(f"{k} x{mnem[k]}" for k in proof) or "none"))
    lines.append("hottest SASS lines by stall samples (share of samples | share of executed | dominant stall | instruction):")
    for x in sorted(rows, key=lambda t: -t[0])[:18]:
        st = sorted(((int(x[3][i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
        lines.append(f"  {x[0] / tot_s * 100:5.2f}% | {x[1] / tot_i * 100:5.2f}% | {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} | {x[2][:80]}")
    open(out, 'w').write("\n".join(lines) + "\n")
    print("wrote", out)


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    # the launch list covers setup + (warmup + steps) bench steps; a step starts at each query_prep_kernel
    names = [r[4] for r in rows]
    ns = [float(r[-1]) for r in rows]
    starts = [i for i, n in enumerate(names) if n.startswith('query_prep_kernel')]
    lines = [f"# {os.path.basename(csv_path)}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)"]
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
        step = list(zip(names[a:b], ns[a:b]))
        tot = sum(t for _, t in step)
        lines.append(f"one device step = {len(step)} launches, {tot / 1e6:.3f} ms under ncu")
        for n, t in step:
            lines.append(f"  {t / 1e3:10.1f} us  {t / tot * 100:5.1f}%  {n[:110]}")
    lines.append("\nall launches (count, total ms):")
    agg = OrderedDict()
    for n, t in zip(names, ns):
        k = n.split('(')[0]
        c, s = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, s + t)
    for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"  {c:4d} x  {s / 1e6:10.3f} ms  {k[:110]}")
    open(out, 'w').write("\n".join(lines) + "\n")
    print("wrote", out)


for name in ('gemm', 'bm25', 'scan'):
    rep = os.path.join(G, f"prof_{name}.ncu-rep")
    if os.path.exists(rep):
        kernel_summary(rep, os.path.join(P, f"{rnd}_{name}_ncu.txt"))
if os.path.exists(os.path.join(G, "launches.csv")):
    launches(os.path.join(G, "launches.csv"), os.path.join(P, f"{rnd}_launches.txt"))
for f in ('bench.json', 'bench_ref.json'):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f"{rnd}_{f}"))
