import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from moco_flow_b200 import ops
from tests.helpers import bf16_round, to_images
dev=torch.device('cuda:0')
n_tiles, pc, qc = 1, 128, 64
gen=torch.Generator().manual_seed(1)
rows=n_tiles*128
Pm=bf16_round(torch.randn(rows,pc,generator=gen)); Qm=bf16_round(torch.randn(rows,qc,generator=gen))
pi,qi=to_images(Pm),to_images(Qm)
rec=np.concatenate([pi.reshape(n_tiles,-1),qi.reshape(n_tiles,-1)],axis=1)
buf=torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
tb=rec.shape[1]*2
out=torch.zeros(pc,qc,device=dev); cs=torch.zeros(pc,device=dev)
ops.dw_gemm(buf,tb,0,pc,buf,tb,(pc//64)*16384,qc,out,pc,qc,n_tiles,cs)
torch.cuda.synchronize()
ref=Pm.double().sum(0)
h0=Pm[:64].double().sum(0); h1=Pm[64:].double().sum(0)
# sums over subsets of rows: groups of 16
g=[Pm[i*16:(i+1)*16].double().sum(0) for i in range(8)]
c=cs.cpu().double()
print('got',c[:6]); print('ref',ref[:6]); print('h0',h0[:6]); print('h1',h1[:6])
A=torch.stack(g,1)  # [128,8]
sol=torch.linalg.lstsq(A, c.unsqueeze(1)).solution.squeeze()
print('coeffs per 16-row group', sol)
coef=torch.linalg.solve(Pm.double().t(), c)   # c = P^T w  -> w per row
print('per-row weights (rounded):')
print(torch.round(coef*100)/100)
