import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
fname=None; hdr=None
agg=collections.Counter(); src={}; stall=collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] in('File Path','File Name'): fname=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; ci=hdr.index('# Samples'); st=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if hdr is None or len(r)<=ci: continue
    if r[2]!='-' : continue   # only source-level summary rows (Address '-')
    try:n=int(r[ci])
    except: continue
    key=(fname,int(r[0])); agg[key]+=n; src[key]=r[1].strip()[:100]
    for i,h in st:
        try: stall[key][h]+=int(r[i])
        except: pass
tot=sum(agg.values()); print('total samples',tot)
for key,n in agg.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 40):
    top=', '.join(f"{h[6:]}:{v}" for h,v in stall[key].most_common(3))
    print(f"{100*n/tot:5.1f}% {key[0]}:{key[1]:<4d} {src[key][:90]}   [{top}]")
