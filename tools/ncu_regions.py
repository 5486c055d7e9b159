"""Folds the SASS source page of an .ncu-rep (captured with --import-source on) into code regions (runs of instructions with
similar execution counts): warp-instructions per 32 KiB tile, share, stall-sample share, active lanes, opcode mix.
    python tools/ncu_regions.py capture.ncu-rep <tiles in the capture>"""
import csv,sys,subprocess
rep=sys.argv[1]; nt=float(sys.argv[2])
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hi=[i for i,r in enumerate(rows) if 'Source' in r][0]
hdr=rows[hi]; data=rows[hi+1:]
isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples'); ithr=hdr.index('Thread Instructions Executed')
tot=sum(int(r[iex]) for r in data); tots=sum(int(r[ismp]) for r in data)
print(f"total warp-instr {tot}  per tile {tot/nt:.0f}  per KiB {tot/nt/32:.1f}  samples {tots}")
cur=None; regs=[]
for k,r in enumerate(data):
    ex=int(r[iex]); key=ex/nt
    if cur is None or abs(cur['key']-key)>0.15*max(cur['key'],1):
        cur={'key':key,'start':k,'n':0,'ex':0,'thr':0,'smp':0,'ops':{}}; regs.append(cur)
    cur['n']+=1; cur['ex']+=ex; cur['thr']+=int(r[ithr]); cur['smp']+=int(r[ismp])
    toks=r[isrc].split(); op=toks[1] if toks[0].startswith('@') else toks[0]
    cur['ops'][op]=cur['ops'].get(op,0)+1
for g in regs:
    if g['ex']/tot<0.008 and g['smp']/tots<0.01: continue
    ops=sorted(g['ops'].items(), key=lambda x:-x[1])[:6]
    print(f"idx {g['start']:4d} n={g['n']:4d} x{g['key']:7.1f}/tile {g['ex']/nt:7.0f} wi/tile ({g['ex']/tot*100:4.1f}%) smp {g['smp']/tots*100:4.1f}% lanes {g['thr']/max(g['ex'],1):4.1f} {ops}")
