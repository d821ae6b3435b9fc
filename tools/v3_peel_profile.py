import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
import sfft_b200.sfft as m
n,k=1<<26,2000
g=torch.Generator().manual_seed(1)
loc=torch.randint(0,n,(k,),generator=g)
xf=torch.zeros(n,dtype=torch.complex128,device='cuda'); xf[loc.cuda()]=1.0
x=(torch.fft.ifft(xf)*n).contiguous(); del xf
p=m.sfft(n,k,3)
for i in range(3):
    cnt=p.execute_device(x,None)
    cyc=p.debug_fetch("peel_cycles",np.int64,8)
    rounds=p.debug_fetch("rounds",np.int32,1)[0]
    names=["decode_mansour","decode_gauss","ans_accumulate","peel_apply(est)","peel_apply(ans)","counts","",""]
    print("count",cnt,"rounds",rounds,{nm:round(c/1.9e3,1) for nm,c in zip(names,cyc) if nm}, "us total", round(cyc.sum()/1.9e3,1))
