import sys, os, time
ROOT='/root/repo'
for p in (ROOT, ROOT+'/relative-entropy-coding_b200', ROOT+'/tests'): sys.path.insert(0,p)
import numpy as np, torch, synth
import __graft_entry__ as g; g.build()
from irec_b200 import engine as E
dev=torch.device('cuda',0)
n_img=32
arrs=[synth.c2(8192, data_seed=i) for i in range(n_img)]
flat=[torch.from_numpy(np.concatenate([a[k] for a in arrs])).to(dev) for k in range(4)]
offs, nb, md = E.make_block_offsets(8192, 1000, dev, n_items=n_img)
for B in (20, 32, 64, 128, 256):
    out=E.beam_encode_blocks(*flat, None, offs, nb, md, 3.0, 36, B, 42)
    torch.cuda.synchronize(); t0=time.perf_counter()
    out=E.beam_encode_blocks(*flat, None, offs, nb, md, 3.0, 36, B, 42)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    na=np.array([len(i) for i in out.indices]); cand=(36+(na-1)*36*min(B,36*36)).sum()
    print(B, nb, round(dt*1e3,1),'ms', '%.3e cand/s'%(cand/dt))
