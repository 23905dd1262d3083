#!/usr/bin/env python
"""Aggregates an ncu report (captured with --import-source on; kernels built with -lineinfo) by CUDA source line:
share of warp instructions, share of stall samples, active lanes per instruction.

    python scripts/ncu_source_lines.py report.ncu-rep 0 40 > profiles/<name>.txt
"""
import csv,sys,subprocess
rep=sys.argv[1]; npts=float(sys.argv[2]) if len(sys.argv)>2 else None
txt=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
cur_file=None; hdr=None; agg=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if r[0] and r[0].isdigit():
        d=dict(zip(hdr,r))
        def f(k):
            try: return float(d.get(k,'0') or 0)
            except: return 0.0
        agg.append((cur_file,int(r[0]),r[1].strip(),f('Instructions Executed'),f('Warp Stall Sampling (All Samples)'),f('Thread Instructions Executed')))
ti=sum(a[3] for a in agg); ts=sum(a[4] for a in agg)
print("total warp inst %.4g samples %.0f"%(ti,ts))
n=int(sys.argv[3]) if len(sys.argv)>3 else 45
for a in sorted(agg,key=lambda a:-a[3])[:n]:
    print("%5.1f%% inst %5.1f%% samp lanes %4.1f %s:%d  %s"%(100*a[3]/ti,100*a[4]/ts,a[5]/max(a[3],1),a[0],a[1],a[2][:90]))
