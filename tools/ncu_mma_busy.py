"""How busy is the MMA-issuing warp?  From the per-instruction samples of `ncu --page source --csv`
exports (one file per kernel launch): the share of all samples that fall in the MMA warp's loop
(barrier wait .. last UTC*MMA .. commit), next to the share ONE warp would have if it were sampled
all the time, and the SASS instructions executed per MMA in that loop.  (Warps suspended in
mbarrier.try_wait with a time hint are not sampled, so the ratio approximates the busy fraction.)

usage: python tools/ncu_mma_busy.py <source.csv>:<warps in the CTA> ...
  ncu -i rep.ncu-rep --page source --csv --launch-skip N --launch-count 1 > source.csv
"""
import csv,sys
def load(p):
    rows=list(csv.reader(open(p)))
    hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
    data=[]
    for r in rows[2:]:
        try: n=int(r[ix['Instructions Executed']]); s=int(r[ix['# Samples']])
        except: continue
        data.append((n,s,r[ix['Source']]))
    return rows[0][1][:70], data
for p,warps in [(a.rsplit(':',1)[0], int(a.rsplit(':',1)[1])) for a in sys.argv[1:]]:
    name,data=load(p)
    # first instance only: cut at first EXIT after the last UTC of the first copy
    mma=[i for i,(n,s,src) in enumerate(data) if 'UTCHMMA' in src or 'UTCIMMA' in src or 'UTCQMMA' in src]
    if not mma: print(name,'no mma'); continue
    # take first contiguous group (gap<200)
    grp=[mma[0]]
    for i in mma[1:]:
        if i-grp[-1]<200: grp.append(i)
        else: break
    a=grp[0]; b=grp[-1]
    # expand to loop: go back to nearest preceding TRYWAIT within 80 and forward to next BRA within 40
    lo=a
    for i in range(a,max(a-80,0),-1):
        if 'TRYWAIT' in data[i][2]: lo=i
    hi=b
    for i in range(b,min(b+40,len(data))):
        if data[i][2].strip().startswith('@P0   BRA') or 'BRA 0x' in data[i][2]: hi=i; break
    tot=sum(s for n,s,_ in data)
    # if the csv holds two copies of the kernel, halve
    dup = 2 if len([1 for n,s,src in data if 'EXIT' in src])>=4 and len(mma)>len(grp) else 1
    reg=sum(s for n,s,_ in data[lo:hi+1])
    ninstr=sum(n for n,s,_ in data[lo:hi+1])
    nm=sum(n for n,s,src in data[lo:hi+1] if 'UTC' in src and 'MMA' in src)
    print('%-70s warps=%2d mma-region samples %.1f%% (one warp = %.1f%%)  instr per MMA %.1f  MMAs %d'%(name, warps, 100.0*reg*dup/tot, 100.0/warps, ninstr/max(nm,1), nm))
