import sys, os, time, subprocess
sys.path.insert(0,os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np
from ngmlr_b200 import synth
root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d='/tmp/e2ebig'; os.makedirs(d, exist_ok=True)
rng=np.random.default_rng(5)
g=[synth.random_genome(2_500_000, 21), synth.random_genome(2_500_000, 22)]
with open(d+'/ref.fa','w') as f:
    for i,c in enumerate(g):
        f.write('>chr%d\n'%(i+1)); s=c.tobytes().decode()
        f.write('\n'.join(s[k:k+80] for k in range(0,len(s),80))+'\n')
n=int(sys.argv[1]) if len(sys.argv)>1 else 400
lens=synth.read_lengths(n, rng, median=8000)
bases=0
with open(d+'/reads.fq','w') as f:
    for i,L in enumerate(lens):
        c=g[i%2]; L=int(L); s0=int(rng.integers(1000,c.size-L-1000))
        r,_=synth.mutate(c[s0:s0+L], rng, err=0.15)
        if rng.integers(0,2): r=synth.revcomp(r)
        s=r.tobytes().decode(); bases+=len(s)
        f.write('@r%d\n%s\n+\n%s\n'%(i,s,'I'*len(s)))
print('reads',n,'bases',bases)
env=dict(os.environ, NGMLR_B200_LIB=root+'/ngmlr_b200/libngmlr_b200.so')
# build caches once (index) with the plain binary so both runs start from the same on-disk index
subprocess.run([root+'/oracle/_ref/ngmlr','-r',d+'/ref.fa','-q',d+'/reads.fq','-o',d+'/warm.sam','-t','32','--no-progress'],capture_output=True,env=env)
for name,exe,t,extra in (('cpu-t32',root+'/oracle/_ref/ngmlr',32,{}),('b200-t16',root+'/oracle/_ref/ngmlr_b200',16,{}),('b200-t64',root+'/oracle/_ref/ngmlr_b200',64,{}),('b200-t16-batcher',root+'/oracle/_ref/ngmlr_b200',16,{'NGMLR_B200_BATCH_WINDOW_US':'200'}),('b200-t64-batcher',root+'/oracle/_ref/ngmlr_b200',64,{'NGMLR_B200_BATCH_WINDOW_US':'200'})):
    t0=time.time()
    r=subprocess.run([exe,'-r',d+'/ref.fa','-q',d+'/reads.fq','-o',d+'/%s.sam'%name,'-t',str(t),'--no-progress'],capture_output=True,text=True,env=dict(env,**extra))
    dt=time.time()-t0
    last=[l for l in r.stderr.splitlines() if 'Done' in l]
    print(name,'wall %.1fs'%dt, 'rc',r.returncode, last[-1][:120] if last else r.stderr[-300:])
import e2e_data
a=e2e_data.sam_records(d+'/cpu-t32.sam'); b=e2e_data.sam_records(d+'/b200-t64-batcher.sam')
print('records',len(a),len(b),'identical',a==b)
