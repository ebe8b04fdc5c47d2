"""Per-op-class stall sampling of the hot loop from an ncu source page:
   ncu -i rep --page source --csv > src.csv ; python tools/ncu_loop_stalls.py src.csv"""
import csv, re, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def col(r,name):
    try: return float(r[ix[name]])
    except: return 0.0
addr=[int(r[0],16) for r in data]
src=[r[1].strip() for r in data]
best=None
for k,s in enumerate(src):
    m=re.search(r'BRA 0x([0-9a-f]+)',s)
    if m:
        t=int(m.group(1),16)
        if t<addr[k] and (addr[k]-t)//16>300 and t in addr: best=(addr.index(t),k)
s0,e0=best
tot=sum(col(r,'# Samples') for r in data); inloop=sum(col(r,'# Samples') for r in data[s0:e0+1])
print('loop: %d instructions; samples total %d, in loop %d (%.1f %%)'%(e0-s0+1,tot,inloop,100*inloop/tot))
reasons=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.Counter()
for r in data[s0:e0+1]:
    for h in reasons: agg[h]+=col(r,h)
print('loop stall reasons:', {k:int(v) for k,v in agg.most_common(8)})
byop=collections.defaultdict(collections.Counter); cnt=collections.Counter()
prev=None
for r in data[s0:e0+1]:
    t=re.sub(r'^@!?U?P\d+\s+','',r[1].strip()); op=t.split()[0]; cls=op.split('.')[0]
    a=re.findall(r'R(\d+)',t)
    if cls=='FFMA2': cls='S1' if '9.99' in t else ('S23' if a[1]==a[2] else ('A1' if 'reuse' in t and prev not in ('A1','A2') else ('A2' if 'reuse' in t else ('A3' if prev in ('A1','A2') else 'A?'))))
    if cls=='FMUL2': cls='Q1' if a[1]==a[2] else 'Q2'
    if cls=='FADD2': cls='F_first' if 'reuse' in t else ('F_second' if prev=='F_first' else 'F?')
    prev=cls
    cnt[cls]+=1
    for h in reasons: byop[cls][h]+=col(r,h)
    byop[cls]['samples']+=col(r,'# Samples')
for cls in sorted(byop, key=lambda c:-byop[c]['samples']):
    c=byop[cls]; print('%-9s n=%3d samples/instr %7.1f  '%(cls,cnt[cls],c['samples']/cnt[cls]), {k[6:]:round(v/cnt[cls],1) for k,v in c.most_common(6) if k!='samples' and v/cnt[cls]>=0.5})
# outside loop top contributors
out=collections.Counter()
for k,r in enumerate(data):
    if not (s0<=k<=e0): out[src[k][:50]]+=col(r,'# Samples')
print('outside loop top:', out.most_common(8))
