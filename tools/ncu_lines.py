#!/usr/bin/env python3
"""Summarise an ncu report of an ensemble kernel: key metrics, instruction mix, and non-FP64 instructions attributed to
CUDA source lines (needs -lineinfo and --import-source on).  Usage: tools/ncu_lines.py report.ncu-rep [top]"""
import collections, csv, subprocess, sys, os
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 22
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread ','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum ','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct','smsp__average_warps_issue_stalled']
for h,u,v in zip(hdr,units,vals):
    if any(h.startswith(w.strip()) for w in want) and 'min' not in h and 'max' not in h:
        try:
            if float(v)==0: continue
        except: pass
        print(f"{h:90s} {u:10s} {v}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
out=collections.defaultdict(collections.Counter); thr=collections.Counter(); file=None; hdr=None; cur=None; files={}
for r in rows:
    if len(r)==2 and r[0]=='File Path': file=r[1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<9: continue
    if r[0].strip().isdigit(): cur=(file,int(r[0])); continue
    sass=r[3].strip()
    if not sass or cur is None: continue
    try: n=int(r[7] or 0); tn=int(r[8] or 0)
    except: continue
    toks=sass.split(); op=(toks[1] if toks[0].startswith('@') else toks[0])
    key='MOV' if (op.startswith('IMAD.MOV') or op=='MOV') else op.split('.')[0]
    out[cur][key]+=n; thr[key]+=tn
tot=collections.Counter()
for k,c in out.items():
    for op,n in c.items(): tot[op]+=n
T=sum(tot.values()); print('total warp instructions',T)
print('mix:',' '.join(f"{op} {n/T*100:.1f}%" for op,n in tot.most_common(18)))
fp=('DMUL','DADD','DFMA','DSETP')
print('FP64 share %.1f%%'%(sum(tot[o] for o in fp)/T*100))
def src_line(k):
    try:
        if k[0] not in files: files[k[0]]=open(k[0]).read().split('\n')
        return files[k[0]][k[1]-1].strip()[:80]
    except Exception: return ''
nf=collections.Counter({k:sum(n for op,n in c.items() if op not in fp) for k,c in out.items()})
print('--- non-FP64 instructions by CUDA line (% of all issued)')
for k,n in nf.most_common(top):
    c=out[k]; det=' '.join(f"{op}:{m/T*100:.2f}" for op,m in c.most_common(5) if op not in fp)
    print(f"  {os.path.basename(k[0])}:{k[1]:4d} {n/T*100:5.2f}%  [{det}]  {src_line(k)}")
