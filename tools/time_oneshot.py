"""tools/time_oneshot.py <scale | ba:<nodes>> <iters> -- the one-shot C-ABI call (pprb200_grank) on R-MAT <scale>, PPR_NUM_GPUS from the
environment: wall time, device time and the library's own phase timing (PPRB200_HOST_TIMING=1)."""
import os, sys, time; sys.path.insert(0, '.')
os.environ.setdefault('PPRB200_HOST_TIMING', '1')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
name = sys.argv[1]; iters = int(sys.argv[2]); scale = name
g = G.barabasi_albert(int(name[3:]), 4) if name.startswith('ba:') else G.rmat(int(name))
for rep in range(3):
    t0 = time.perf_counter()
    r = ppr.grank_csr(g, 50, 100, iters, 0.85, -1.0)
    dt = time.perf_counter() - t0
    print(f"rmat{scale} it={iters} PPR_NUM_GPUS={os.environ.get('PPR_NUM_GPUS')} n_gpus={r.stats['n_gpus']}: wall {dt*1e3:.1f} ms kernel_ms {r.stats['kernel_ms']:.1f} "
          f"prep {r.stats['prep_ms']:.1f} d2h {r.stats['d2h_ms']:.1f} requeues {r.stats['overflow_requeues']}", flush=True)
