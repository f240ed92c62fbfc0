import csv, sys, collections, subprocess
rep, kern = sys.argv[1], sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 40
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name",f"regex:{kern}","--launch-count","1","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur_file=None; hdr=None; ix=None; last=('?',0)
agg=collections.defaultdict(lambda:[0,0,collections.Counter(),""])
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; ix={}; 
    if r[0]=="Line No":
        for i,h in enumerate(hdr): ix.setdefault(h,i)
        continue
    if hdr is None or len(r)<len(hdr): continue
    try:
        smp=int(r[ix['# Samples']]); ins=int(r[ix['Instructions Executed']])
    except: continue
    if not r[0].strip().isdigit(): continue      # SASS rows repeat what their source row already aggregates
    last=(cur_file,int(r[0])); agg[last][3]=r[1]
    key=last
    a=agg[key]; a[0]+=smp; a[1]+=ins
    for h in ('stall_barrier','stall_long_sb','stall_short_sb','stall_wait','stall_math','stall_mio','stall_lg','stall_branch_resolving','stall_no_inst'):
        try: a[2][h]+=int(r[ix[h]])
        except: pass
tot=sum(a[0] for a in agg.values()); tin=sum(a[1] for a in agg.values())
print('total samples',tot,'warp instr',tin)
top=sorted(agg.items(), key=lambda kv:-kv[1][0])[:topn]
for (f,l),a in sorted(top):
    st=", ".join(f"{k[6:]}={v/max(a[0],1):.2f}" for k,v in a[2].most_common(2))
    print(f"{f}:{l:<5d} smp={a[0]/tot:.3f} ins={a[1]/tin:.3f} [{st}] {a[3].strip()[:95]}")
