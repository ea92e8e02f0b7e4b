import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200 import ops, _lib as L
B,H,nt,ni,nc=1,24,512,1024,1024
S=nt+ni+nc
g=torch.Generator(device="cuda").manual_seed(1)
q,k,v=[torch.randn((B,H,S,128),generator=g,device="cuda").bfloat16() for _ in range(3)]
out=torch.empty((B*S,H*128),device="cuda",dtype=torch.bfloat16)
orb=ops.make_out_row_base(B,nt,ni,nc,"cuda")
for _ in range(3): ops.attention(q,k,v,out,orb,n_cond=nc)
dbg=torch.zeros(2*20*8+2*20*4,dtype=torch.int64,device="cuda")
L.lib.lx_attention_debug_timeline.argtypes=[C.c_void_p]
L.lib.lx_attention_debug_timeline(dbg.data_ptr())
ops.attention(q,k,v,out,orb,n_cond=nc)
torch.cuda.synchronize()
L.lib.lx_attention_debug_timeline(None)
arr=dbg.cpu()[2*20*8:].view(2,20,4)
t=dbg.cpu()[:2*20*8].view(2,20,8)
t0=t[0,0,0].item()
names=["loop","s_full","tmem_ld","max","exp+P_st","st_wait+arrive"]
for gi in range(2):
    g_=gi
    print(f"query tile {gi}")
    print("iter  start   "+"  ".join(f"{n:>14s}" for n in names[1:]))
    for i in range(20):
        row=t[gi,i].tolist()
        d=[row[j]-row[j-1] for j in range(1,6)]
        print(f"{i:3d} {row[0]-t0:8d}  "+"  ".join(f"{x:14d}" for x in d)+f"   total {row[5]-row[0]}   | arrive@{row[5]-t0} mma_saw_pfull@{row[6]-t0} qk_issued@{row[7]-t0} next_s_full@{(t[gi,i+1,1].item()-t0) if i<19 else 0}")
for gi in range(2):
    for i in range(10,14): print('arrive by warp quarter', gi, i, [x-t[0,0,0].item() for x in arr[gi,i].tolist()], 'mma saw', t[gi,i,6].item()-t[0,0,0].item())
# throughput
import time
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for Bb in (1,4):
    q,k,v=[torch.randn((Bb,H,S,128),generator=g,device="cuda").bfloat16() for _ in range(3)]
    out=torch.empty((Bb*S,H*128),device="cuda",dtype=torch.bfloat16)
    orb=ops.make_out_row_base(Bb,nt,ni,nc,"cuda")
    for _ in range(5): ops.attention(q,k,v,out,orb,n_cond=nc)
    e0.record()
    for _ in range(20): ops.attention(q,k,v,out,orb,n_cond=nc)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"B={Bb}: {ms:.4f} ms  {4*Bb*H*S*S*128/ms/1e9:.1f} TFLOP/s")
