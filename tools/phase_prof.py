"""tools/phase_prof.py <scale> [iters] [hub] -- per-phase cycle shares of the order-free merge kernels (PPRB200_PROF=1)
on R-MAT <scale>: merge_dense_kernel (big / mid instantiation) or, with PPRB200_DENSE=0, merge_par_kernel."""
import sys, os, ctypes as C
sys.path.insert(0, '.')
os.environ['PPRB200_PROF'] = '1'
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G, _lib
scale = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6; hub = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = G.rmat(scale); col = ppr.find_partitions_csr(g)
s = ppr.Session(g, 100, colour=col, hub_threshold=hub)
lib = _lib.load()
dense = os.environ.get('PPRB200_DENSE', '1') != '0'
names = (['fetch', 'theta+pass1', 'compact', 'sketch scan', 'pass2+tail', 'select', 'write+norm', 'clear'] if dense else
         ['fetch', 'setup', 'accum/pass1', 'compact+tau(+redo)', 'pass2(+flush)', 'select', 'write+norm', 'clear'])
for rep in range(2):
    s.grank(50, 100, iters, 0.85, -1.0)
    st = s.stats(); l, ms = s.kernel_time(0)
    print(f"rmat{scale} hub>{hub} dense={dense}: kernel_ms {st['kernel_ms']:.2f} merge_ms {ms:.2f} ({ms/iters:.2f}/iteration) alg GB/s {st['algorithmic_bytes']/ms/1e6:.1f} "
          f"frac {st['algorithmic_bytes']/ms/1e6/6543.1:.3f} requeues {st['overflow_requeues']} merged {st['merged_entries']}")
    if dense:
        d = (C.c_ulonglong * 8)()
        lib.pprb200_debug_counters(s.handle, d)
        print("  dense kernels: nodes done %d, ran pass 2 %d, tau=0 %d; hubs (>1024 successors) seen %d / ran pass 2 %d; handed over: candidates>CMAX %d, tail full %d, old basket not full %d" % (d[0], d[1], d[2], d[3], d[6], d[4], d[5], d[7]))
    buf = np.zeros(2 * 148 * 8 * 8, dtype=np.uint64); n = C.c_int(0)
    lib.pprb200_debug_prof(s.handle, buf.ctypes.data_as(C.c_void_p), C.byref(n))
    buf = buf.reshape(2, 148 * 8, 8)
    for cls, nm in ((1, 'big'), (0, 'mid')):
        b = buf[cls].astype(np.float64); tot = b.sum(1); act = tot > 0
        if not act.any(): continue
        print(f"  {nm}: CTAs {act.sum()} total cyc/CTA mean {tot[act].mean():.3e} max {tot[act].max():.3e}")
        print("    mean share:", {k: round(float(b[act][:, i].sum() / tot[act].sum()), 3) for i, k in enumerate(names)})
