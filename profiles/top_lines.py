"""Executed instructions and stall samples per CUDA source line of one kernel, from an ncu report captured with
`--set full --import-source on` (kernels are built with -lineinfo).

  python profiles/top_lines.py gpurun_out/r2i_prof_all.ncu-rep ::regex:gl2_bwd_q:1 30

This view found the two integer divisions in the tile prologue of the gl2 backward (3 % of its instructions,
7 % of its time) and shows how much of every kernel's instruction count is mbarrier polling.
"""
import csv,collections,sys,subprocess
rep, kid, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-id',kid],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
cur=None; hdr=None
agg=collections.defaultdict(lambda:[0,0,''])
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': fn=r[1][:90]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None: continue
    try: ln=int(r[0])
    except: continue
    ia=hdr.index('Address'); ie=hdr.index('Instructions Executed'); iw=hdr.index('Warp Stall Sampling (All Samples)')
    if r[ia]=='' or r[ia]=='-':
        try: k=int(r[ie]); w=int(r[iw])
        except: continue
        agg[(cur,ln)][0]+=k; agg[(cur,ln)][1]+=w; agg[(cur,ln)][2]=r[1].strip()[:110]
tot=sum(v[0] for v in agg.values()); tw=sum(v[1] for v in agg.values())
print(fn); print('total',tot,tw)
for (f,ln),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:n]:
    print('%-12s %5d %5.1f%% instr %5.1f%% samp  %s'%(f[:12],ln,100*v[0]/tot,100*v[1]/max(tw,1),v[2]))
