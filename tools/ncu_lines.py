"""Per-source-line summary of an `ncu --page source --csv` export: samples, instructions, shared wavefronts, top stalls."""
import csv, sys
def I(s):
    try: return int(s)
    except Exception: return 0
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=None; cur_file=None
lines=[]
for r in rows:
    if r and r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if r and r[0]=="Function Name": continue
    if hdr and r and r[0] not in ("",):
        d=dict(zip(hdr,r))
        st={k:I(d[k]) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
        lines.append((cur_file,r[0],r[1][:100],I(r[6]),I(r[7]),I(d['L1 Wavefronts Shared']),I(d['L1 Wavefronts Shared Ideal']),st))
tot=sum(l[3] for l in lines); print("total samples",tot, "total inst", sum(l[4] for l in lines), "wavefronts", sum(l[5] for l in lines), "ideal", sum(l[6] for l in lines))
print("--- by samples")
for l in sorted(lines,key=lambda l:-l[3])[:topn]:
    top=sorted(l[7].items(),key=lambda kv:-kv[1])[:3]
    top=[(k.replace('stall_',''),v) for k,v in top if v]
    print(f"{l[0]}:{l[1]} s={l[3]} ({100*l[3]/tot:.1f}%) inst={l[4]} wf={l[5]}/{l[6]} {top} | {l[2]}")
agg={}
for l in lines:
    for k,v in l[7].items(): agg[k]=agg.get(k,0)+v
print("--- stall totals", sorted(agg.items(),key=lambda kv:-kv[1])[:10])
