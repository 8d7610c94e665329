#!/usr/bin/env python
"""Aggregate an ncu report of dram_kernel by source function: executed warp instructions and stall samples."""
import csv, subprocess, io, sys
rep=sys.argv[1]; nsteps=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur=""; hdr=None; lines=[]
def fl(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0].isdigit():
        d=dict(zip(hdr[4:], r[4:]))
        lines.append((cur,int(r[0]),r[1],fl(d.get("# Samples",0)),fl(d.get("Instructions Executed",0))))
te=sum(l[4] for l in lines); ts=sum(l[3] for l in lines)
def marks_of(path, pats):
    src=open(path).read().split('\n'); out=[]
    for name,pat in pats:
        for i,l in enumerate(src):
            if pat in l: out.append((name,i+1)); break
    return sorted(out,key=lambda x:x[1])
mm=marks_of('transcriptioncycleinference_b200/csrc/tc_mcmc.cu',[('ss_batch_kernel','void __launch_bounds__(SS_THREADS, SS_MIN_CTAS'),('ss_stream_kernel','void __launch_bounds__(SS_THREADS, 2) ss_stream'),('chol_tiled','bool chol_tiled('),('dmma','void dmma_m8n8k4'),('tma helpers','#define TMA_MAXST'),('chol_global','bool chol_global('),('cand/state structs','struct Cand'),('gen_increments_tma','void gen_increments_tma('),('flush_run','void flush_run('),('emit_s2','void emit_s2('),('generate','void generate('),('cand_bounds','void cand_bounds('),('resolve_dr','int resolve_dr('),('adapt','int adapt('),('dram_kernel(main)','dram_kernel(const __grid_constant__'),('after','RNG dump / FP64 peak')])
md=marks_of('transcriptioncycleinference_b200/csrc/tc_device.cuh',[('philox/draw','philox_round('),('normal_pair','void normal_pair('),('chi2_draw','double chi2_draw('),('exp/log','double tc_exp('),('warp_sum','double warp_sum('),('views','struct CellView'),('scan','void scan_counts_sequential('),('first_lag','int first_lag('),('rows_pairs','void rows_pairs('),('ss_eval','double ss_eval(')])
agg={}
for f,l,s,sa,ie in lines:
    key=f
    for tag,marks in (('tc_mcmc.cu',mm),('tc_device.cuh',md)):
        if f==tag:
            key=f+':?'
            for i,(n,a) in enumerate(marks):
                b=marks[i+1][1] if i+1<len(marks) else 10**9
                if a<=l<b: key=n
    agg.setdefault(key,[0,0]); agg[key][0]+=sa; agg[key][1]+=ie
print("total warp instr %.3e (%.0f per unit), samples %d"%(te,te/nsteps,ts))
for k,(sa,ie) in sorted(agg.items(), key=lambda x:-x[1][1]): print("%-28s inst %5.1f%% (%7.0f/unit)  samples %5.1f%%"%(k,100*ie/te,ie/nsteps,100*sa/ts))
