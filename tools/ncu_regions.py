"""Summarise an `ncu --page source --csv` dump by code region (runs of SASS with similar execution
counts): warp-instructions per tile, share, active lanes, stall-sample share and the opcode mix."""
import csv, sys
path, ntiles = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); ithr = hdr.index("Thread Instructions Executed")
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
reg = []; cur = None
for k, r in enumerate(data):
    ex = int(r[iex]); key = ex / ntiles
    if cur is None or abs(cur['key'] - key) > 0.12 * max(cur['key'], 1.0):
        cur = {'key': key, 'start': k, 'n': 0, 'ex': 0, 'thr': 0, 'smp': 0, 'ops': {}}; reg.append(cur)
    cur['n'] += 1; cur['ex'] += ex; cur['thr'] += int(r[ithr]); cur['smp'] += int(r[ismp])
    toks = r[isrc].split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + 1
print(f"warp-instr per tile: {tot / ntiles:.0f}   thread-instr per byte: {sum(int(r[ithr]) for r in data) / ntiles / 16384:.2f}   samples {tots}")
for g in reg:
    if g['ex'] / tot < 0.008 and g['smp'] / tots < 0.01: continue
    ops = sorted(g['ops'].items(), key=lambda x: -x[1])[:6]
    print(f"idx {g['start']:4d} n={g['n']:4d} x{g['key']:6.1f}/tile {g['ex'] / ntiles:6.0f} wi/tile ({g['ex'] / tot * 100:4.1f}%) smp {g['smp'] / tots * 100:4.1f}% lanes {g['thr'] / max(g['ex'], 1):4.1f} {ops}")
