import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngmlr_b200 import synth, refindex, B200Aligner, PackedBatch, PackedReads, split_read
g=synth.random_genome(50_000_000,1)
contigs=[g[i*10_000_000:(i+1)*10_000_000] for i in range(5)]
ref=refindex.encode_reference(contigs); idx=refindex.build_index(ref)
pool=synth.pacbio_problems(8192, seed=2, median=8000, genome=g)
batch=PackedBatch.from_problems(pool); subs=PackedReads([s for p in pool for s in split_read(p.qry)])
al=B200Aligner(0); al.set_index(idx); al.set_reference(ref)
import torch
def T(f,*a):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(*a); torch.cuda.synchronize(); return (time.perf_counter()-t)*1e3, r
for it in range(3):
    t1,_=T(al.cs_upload,subs); t2,_=T(al.cs_run); t3,_=T(al.cs_fetch)
    t4,_=T(al.upload,batch); t5,_=T(al.run); t6,_=T(al.fetch)
    print('cs_upload %.1f cs_run %.1f cs_fetch %.1f | upload %.1f run %.1f fetch %.1f | total %.1f'%(t1,t2,t3,t4,t5,t6,t1+t2+t3+t4+t5+t6), al.stats()['host_pack_ms'], al.stats()['host_h2d_ms'], al.stats()['host_text_ms'])
