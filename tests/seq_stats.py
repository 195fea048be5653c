"""CPU tool: sequence statistics of a level-3 / 64 KiB frame made by the reference (match distances against the
decode kernel's 4 KiB ring, exact dependency depth per batch of 32 / 64 sequences, length mix).  python tests/seq_stats.py"""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
import zxc_ctypes as z, zxc_corpus as zc
ref=z.ZxcLib(z.REF_SO)
data=zc.silesia_shaped(64<<20,seed=1)
fr=ref.compress(data,level=3,block_size=65536,seekable=0)
b=fr.tobytes()
p=16; blocks=[]
while True:
    t=b[p]; comp=int.from_bytes(b[p+3:p+7],'little')
    if t==255: break
    blocks.append((t,p+8,comp)); p+=8+comp
print(len(blocks),"blocks")
def varint(buf,pos):
    b0=buf[pos]
    if b0<0x80: return b0,pos+1
    if b0<0xC0: return (b0&0x3F)|(buf[pos+1]<<6),pos+2
    return (b0&0x1F)|(buf[pos+1]<<5)|(buf[pos+2]<<13),pos+3
offs_all=[];ml_all=[];ll_all=[];depth32=[];depth64=[]
rng=np.random.default_rng(0)
sel=rng.choice(len(blocks),120,replace=False)
for bi in sel:
    t,d,comp=blocks[bi]
    if t!=1: continue
    n_seq=int.from_bytes(b[d:d+4],'little'); n_lit=int.from_bytes(b[d+4:d+8],'little'); enc_lit=b[d+8]; enc_off=b[d+11]
    desc=4 if enc_lit else 0
    lit_comp=int.from_bytes(b[d+12:d+16],'little') if enc_lit else n_lit
    tok=np.frombuffer(b,np.uint8,n_seq,d+12+desc+lit_comp)
    o0=d+12+desc+lit_comp+n_seq
    if enc_off: off=np.frombuffer(b,np.uint8,n_seq,o0).astype(np.int64)+1; ext=o0+n_seq
    else: off=np.frombuffer(b,'<u2',n_seq,o0).astype(np.int64)+1; ext=o0+2*n_seq
    ll=(tok>>4).astype(np.int64); ml=(tok&15).astype(np.int64)
    for i in np.nonzero((ll==15)|(ml==15))[0]:
        if ll[i]==15: v,ext=varint(b,ext); ll[i]+=v
        if ml[i]==15: v,ext=varint(b,ext); ml[i]+=v
    ml+=5
    tot=ll+ml; end=np.cumsum(tot); mdst=end-ml; src_lo=mdst-off; src_end=np.minimum(mdst,src_lo+ml)
    offs_all.append(off); ml_all.append(ml); ll_all.append(ll)
    for B,acc in ((32,depth32),(64,depth64)):
        for s in range(0,n_seq,B):
            e=min(n_seq,s+B); dep=np.zeros(e-s,np.int64)
            for j in range(s,e):
                # blockers: earlier lanes in batch whose dest [mdst,mend) intersects [src_lo,src_end)
                lo=src_lo[j]; hi=src_end[j]
                k=np.nonzero((mdst[s:j]+ml[s:j]>lo)&(mdst[s:j]<hi))[0]
                dep[j-s]=(dep[k].max()+1) if k.size else 0
            acc.append(dep.max()+1)
off=np.concatenate(offs_all); ml=np.concatenate(ml_all); ll=np.concatenate(ll_all)
print("seqs",off.size,"avg ll %.2f ml %.2f"%(ll.mean(),ml.mean()))
for lim in (256,1024,2048,3500,4096,8192,16384,32768,65536):
    m=off<=lim
    print("off<=%6d: %.1f%% of matches, %.1f%% of match bytes"%(lim,100*m.mean(),100*ml[m].sum()/ml.sum()))
print("match rounds per batch of 32: mean %.2f ; per batch of 64: mean %.2f"%(np.mean(depth32),np.mean(depth64)))
print("ll<=20: %.1f%%  ml<=20: %.1f%%  ml<=32 %.1f%% off<ml %.1f%% off<36 %.1f%%"%(100*(ll<=20).mean(),100*(ml<=20).mean(),100*(ml<=32).mean(),100*(off<ml).mean(),100*(off<36).mean()))
print("literal bytes share %.1f%%"%(100*ll.sum()/(ll.sum()+ml.sum())))
