import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200 import ops, _lib as L
B,H,nt,ni,nc=int(sys.argv[1]) if len(sys.argv)>1 else 1,24,512,1024,1024
S=nt+ni+nc
g=torch.Generator(device="cuda").manual_seed(1)
q,k,v=[torch.randn((B,H,S,128),generator=g,device="cuda").bfloat16() for _ in range(3)]
out=torch.empty((B*S,H*128),device="cuda",dtype=torch.bfloat16)
orb=ops.make_out_row_base(B,nt,ni,nc,"cuda")
for _ in range(3): ops.attention(q,k,v,out,orb,n_cond=nc)
n=B*H*S//256
tr=torch.zeros(n*6,dtype=torch.int64,device="cuda")
L.lib.lx_attention_debug_cta_trace.argtypes=[C.c_void_p]
L.lib.lx_attention_debug_cta_trace(tr.data_ptr())
ops.attention(q,k,v,out,orb,n_cond=nc)
torch.cuda.synchronize()
L.lib.lx_attention_debug_cta_trace(None)
t=tr.cpu().view(n,6)
import collections
by=collections.defaultdict(list)
for r in t.tolist(): by[r[0]].append(r)
setup=[r[2]-r[1] for r in t.tolist()]; first=[r[3]-r[2] for r in t.tolist()]; loop=[r[4]-r[3] for r in t.tolist()]; epi=[r[5]-r[4] for r in t.tolist()]
import statistics as st
print("per-CTA clk: setup %d  first_S %d  loop %d  epilogue+exit %d   total %d"%(st.median(setup),st.median(first),st.median(loop),st.median(epi),st.median([r[5]-r[1] for r in t.tolist()])))
gaps=[]
for sm,rs in by.items():
    rs.sort(key=lambda r:r[1])
    for a,b in zip(rs,rs[1:]): gaps.append(b[1]-a[5])
if gaps: print("gap between consecutive CTAs on one SM (exit -> next entry): median %d  max %d  n %d"%(st.median(gaps),max(gaps),len(gaps)))
allt=[(max(r[5] for r in rs)-min(r[1] for r in rs)) for rs in by.values()]
print("SMs used", len(by), "busy span per SM: median", st.median(allt), "max", max(allt))
