import csv,sys,subprocess,collections
rep=sys.argv[1]
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=[i for i,x in enumerate(rows) if x and x[0]=='Line No']
hdr=rows[hi[0]]
iL=0; iS=1; iE=hdr.index('Instructions Executed'); iW=hdr.index('Warp Stall Sampling (All Samples)')
agg=collections.Counter(); st=collections.Counter(); txt={}
for x in rows[hi[0]+1:]:
    if len(x)<=iE: continue
    try: n=int(x[iE]); ln=int(x[iL])
    except: continue
    agg[ln]+=n; txt[ln]=x[iS].strip()[:100]
    try: st[ln]+=int(x[iW])
    except: pass
tot=sum(agg.values()); tots=sum(st.values()) or 1
print("total inst", tot)
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for ln,n in agg.most_common(N): print(f"{100*n/tot:5.1f}% inst {100*st[ln]/tots:5.1f}% stall | L{ln}: {txt[ln]}")
