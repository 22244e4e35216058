import sys, os, subprocess, json
# usage: sweep.py scale  "ENV=V ENV2=V2" "..." ; runs g-style timing in subprocesses (env is read at session creation)
scale=sys.argv[1]
code='''
import sys; sys.path.insert(0,'.')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
g=G.rmat(%s); col=ppr.find_partitions_csr(g)
import os
s=ppr.Session(g,100,colour=col,hub_threshold=int(os.environ.get('HUB','0')))
best=1e9
for r in range(3):
    s.grank(50,100,30,0.85,-1.0); best=min(best,s.stats()['kernel_ms'])
print(best, s.stats()['overflow_requeues'])
''' % scale
for spec in sys.argv[2:]:
    env=dict(os.environ)
    for kv in spec.split():
        if '=' in kv: k,v=kv.split('=',1); env[k]=v
    r=subprocess.run([sys.executable,'-c',code],env=env,capture_output=True,text=True)
    print(f"rmat{scale} [{spec}] -> {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)
