"""Summarise an ncu report: key raw metrics + top stall instructions per kernel.
usage: python tools/ncu_top.py report.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); hdr = r[0]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active']
seen = set()
for row in r[2:]:
    name = row[hdr.index('Kernel Name')][:70]
    if name in seen: continue
    seen.add(name); print('==', name)
    for k in keys:
        if k in hdr: print(f'   {k} = {row[hdr.index(k)]} {r[1][hdr.index(k)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name']
done = set()
for si, s in enumerate(starts):
    name = rows[s][1][:70]
    if name in done: continue
    done.add(name)
    e = starts[si + 1] if si + 1 < len(starts) else len(rows)
    h = rows[s + 1]; body = rows[s + 2:e]
    isamp = h.index('# Samples'); isrc = h.index('Source')
    sc = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    tot = sum(int(x[isamp]) for x in body if x[isamp].isdigit())
    agg = {h[i]: 0 for i in sc}
    for x in body:
        for i in sc:
            if x[i].isdigit(): agg[h[i]] += int(x[i])
    print('== stalls', name, 'samples', tot)
    print('  ', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]})
    for x in sorted(body, key=lambda q: -int(q[isamp]) if q[isamp].isdigit() else 0)[:ntop]:
        st = {h[i]: int(x[i]) for i in sc if x[i].isdigit() and int(x[i]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"  {int(x[isamp]):6d} {100*int(x[isamp])/max(tot,1):5.1f}%  {x[isrc].strip()[:64]:64s} {st}")
