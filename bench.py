#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 GRank / MCCompletePathV2 hot paths.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rmat22|rmat16|rmat20mc|ba8m|ba8mmc|ring ...]
                    [--impl reference] [--hub-threshold T] [--no-mc] [--no-exact] [--no-e2e] [--no-cpu-baseline]

One "step" = one whole job of the hot path on one synthetic graph (GRank: init + `iterations` merge sweeps +
final top-K; MC: walks + combine rounds + top-K). Default workload = the configuration BASELINE.json's target is
quoted on: GRank on R-MAT scale 22 (4 194 304 nodes, 67 108 864 edges), K=50, L=100, 30 iterations, damping 0.85
(configs[3]; it fits one GPU), the same job at every --gpus N (strong scaling over source shards). The default line
also carries, under "mc", the second half of BASELINE's metric -- MCCompletePathV2 walk-steps/s on R-MAT scale 20,
R=1000 (configs[2]) -- with its own value / roofline / e2e / cpu_baseline, and under "exact_order" the same GRank job
with the reference's exact fma order everywhere (hub_threshold = never), so the price of strict parity is on record.

  value      node-iterations/s (walk-steps/s for MC) with graph + baskets resident in HBM (session API),
             timed with CUDA events on the launching stream, L2 flushed between steps
  e2e        same metric through the host-buffer C-ABI call (pprb200_grank / pprb200_mccompletepathv2): host
             preprocessing, H2D of the CSR from pinned memory and D2H of the baskets are inside the timed region;
             e2e_api (N=1): the reference-facing template API itself (ppr::grank on an unordered_map, map-of-maps out)
  roofline   merge kernels (GRank) / walk kernel (MC): algorithmic bytes (counted by the kernels) over
             the event-timed duration of those launches, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  bounded sample of the same workload on the host cores (rank 0, N=1)

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libppr_ref.so; the oracle port if that is absent) on
the same graph and parameters on all host cores: the full job where that takes seconds (R-MAT-16), otherwise one
step of the first two iterations (one sweep of each partition) -- the line says which in `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

DEFAULT_WORKLOAD = "rmat22"
WORKLOADS = {
    "ring": dict(kind="grank", gen="ring", scale=100, K=50, L=100, iterations=30, damping=0.85, tolerance=1e-3,
                 desc="grank on README's 100-node ring (BASELINE configs[0])"),
    "rmat16": dict(kind="grank", gen="rmat", scale=16, K=50, L=100, iterations=30, damping=0.85, tolerance=1e-3,
                   desc="GRank on R-MAT scale 16 (65536 nodes, 1048576 edges), K=50 L=100 30 it d=0.85 tol=1e-3 (BASELINE configs[1])"),
    "rmat18": dict(kind="grank", gen="rmat", scale=18, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 18"),
    "rmat20": dict(kind="grank", gen="rmat", scale=20, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 20"),
    "rmat22": dict(kind="grank", gen="rmat", scale=22, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 22 (4194304 nodes, 67108864 edges), K=50 L=100 30 it d=0.85 (BASELINE configs[3])"),
    "rmat20mc": dict(kind="mc", gen="rmat", scale=20, K=50, L=100, iterations=1000, damping=0.85, tolerance=0.0,
                     desc="MCCompletePathV2 on R-MAT scale 20 (1048576 nodes, 16777216 edges), K=50 L=100 R=1000 d=0.85 (BASELINE configs[2])"),
    "rmat16mc": dict(kind="mc", gen="rmat", scale=16, K=50, L=100, iterations=1000, damping=0.85, tolerance=0.0,
                     desc="MCCompletePathV2 on R-MAT scale 16, K=50 L=100 R=1000 d=0.85"),
    "ba8m": dict(kind="grank", gen="ba", scale=8388608, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                 desc="GRank on Barabasi-Albert 8388608 nodes m=4 symmetrised, K=50 L=100 30 it (BASELINE configs[4])"),
    "ba8mmc": dict(kind="mc", gen="ba", scale=8388608, K=50, L=100, iterations=1000, damping=0.85, tolerance=0.0,
                   desc="MCCompletePathV2 on Barabasi-Albert 8388608 nodes m=4 symmetrised, K=50 L=100 R=1000 (BASELINE configs[4])"),
}
COUNTERS = ("nonsink_node_iterations", "edge_reads", "merged_entries", "candidates", "truncations", "boundary_ties",
            "overflow_requeues", "walk_steps", "walks")
NEVER_HUB = 0xFFFFFFFF
# reference / CPU samples: graphs up to this many nodes run the whole job, larger ones the first two iterations
REFERENCE_FULL_JOB_NODES = 1 << 16
REFERENCE_SAMPLE_ITERATIONS = 2   # one sweep of each partition: exactly n node-iterations
REFERENCE_MC_SAMPLE_R = 50        # walk budget of the bounded MC sample on graphs above REFERENCE_FULL_JOB_NODES


def make_graph(w, numpy_only=False):
    from approximated_personalized_pagerank_b200 import graphs as G
    if w["gen"] == "ring":
        return G.ring(w["scale"])
    if w["gen"] == "rmat":
        return G.rmat_numpy(w["scale"]) if numpy_only else G.rmat(w["scale"])
    if w["gen"] == "ba":
        return G.barabasi_albert(w["scale"], 4)
    raise ValueError(w["gen"])


def committed_profile(key):
    """ncu-derived figures of the committed captures (profiles/traffic.json): DRAM bytes per step, L2 throughput"""
    try:
        return json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(key)
    except Exception:
        return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [x for x in sm if x > 0.5 * max(mx)] if sm else []
        return {"sm_mhz": statistics.median(busy or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def metric_of(w):
    return ("grank_node_iterations_per_s", "node-iterations/s") if w["kind"] == "grank" else ("mc_walk_steps_per_s", "walk-steps/s")


# ------------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (test infrastructure, executed only here and in cpu_baseline)
# ------------------------------------------------------------------------------------------------------------------
def reference_sample(w, g, cores, steps=1, warm=0, allow_reference=True):
    """Times the reference (oracle/_ref) or, where that is absent / too large, the oracle port on a bounded sample of
    workload w. Returns (units, seconds, kind, cores_used, sample description, config note)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_bindings as ob
    full = g.n <= REFERENCE_FULL_JOB_NODES
    have_ref = allow_reference and ob.have_ref()
    times, units = [], []
    if w["kind"] == "grank":
        it = w["iterations"] if full else min(w["iterations"], REFERENCE_SAMPLE_ITERATIONS)
        # the reference keeps ~9 KB of hash nodes per node (two maps of <= L entries): stay well inside the box's memory
        try:
            avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
        except (ValueError, OSError):
            avail = 0
        use_ref = have_ref and g.n * 12000 < 0.6 * avail
        kind = "reference" if use_ref else "port"
        for s in range(warm + steps):
            if use_ref:
                r = ob.ref_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
                sec = r.seconds
                if full:  # the tolerance may stop the run early: count what actually ran with the port's colouring
                    colour = ob.oracle_find_partitions(g.relabel(ob.ref_iteration_order(g)))
                    ran = ob.oracle_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores).stats["iterations_run"]
                    c0 = int((colour == 0).sum())
                    u = sum(c0 if (i & 1) == 0 else g.n - c0 for i in range(ran))
                else:
                    u = g.n * it // 2  # an even number of iterations visits every node it/2 times
            else:
                t0 = time.perf_counter()
                r = ob.oracle_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
                sec = time.perf_counter() - t0
                u = r.stats["node_iterations"]
            if s >= warm:
                times.append(sec); units.append(u)
        name = f"grankMulti nThreads={cores}" if use_ref else f"OpenMP oracle port, {cores} threads"
        sample = (f"{name}, the whole {it}-iteration job" if full else
                  f"{name}, first {it} of {w['iterations']} iterations (one sweep of each partition); node-iterations/s extrapolates to the 30-iteration job")
        note = "full job" if full else f"bounded sample: iterations={it} of {w['iterations']} (extrapolated, BASELINE.md 3)"
        return sum(units), sum(times), kind, cores, sample, note
    # MC: the reference is single-threaded by construction; hops are counted by the port at the same walk budget
    # (bounded: R = 50 on a million nodes is ~12 s on 16 threads; larger graphs get proportionally fewer walks per source)
    R = w["iterations"] if full else max(2, min(w["iterations"], REFERENCE_MC_SAMPLE_R * (1 << 20) // max(g.n, 1 << 20)))
    use_ref = have_ref
    kind = "reference" if use_ref else "port"
    hops = ob.oracle_mc(g, w["K"], w["L"], R, w["damping"], 1, 0, nthreads=cores).stats["walk_steps"]
    for s in range(warm + steps):
        if use_ref:
            sec = ob.ref_mc(g, w["K"], w["L"], R, w["damping"]).seconds
        else:
            t0 = time.perf_counter()
            ob.oracle_mc(g, w["K"], w["L"], R, w["damping"], 1, 3, nthreads=cores)
            sec = time.perf_counter() - t0
        if s >= warm:
            times.append(sec); units.append(hops)
    used = 1 if use_ref else cores
    sample = (f"{'mccompletepathv2 (single-threaded by construction)' if use_ref else 'OpenMP oracle port'} R={R}"
              + ("" if full else f" of {w['iterations']} (walk-steps/s extrapolates to the full walk budget)"))
    note = "full job" if full else f"bounded sample: R={R} of {w['iterations']} (extrapolated)"
    return sum(units), sum(times), kind, used, sample, note


def run_reference(args, w):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only; no GPU, libppr_b200.so not loaded)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    g = make_graph(w, numpy_only=True)
    cores = os.cpu_count() or 1
    full = g.n <= REFERENCE_FULL_JOB_NODES
    steps = max(1, args.steps) if full else 1
    warm = max(0, min(args.warmup, 1)) if full else 0
    u, sec, kind, used, sample, note = reference_sample(w, g, cores, steps, warm)
    metric, unit = metric_of(w)
    value = u / sec
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": w["desc"], "reference_sample": note},
            "cpu_baseline": {"value": value, "unit": unit, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(w, g):
    """bounded sample of the same workload on the host cores (rank 0, N=1 only): the OpenMP port for the big graphs
    (the reference's hash maps need tens of GB there; `--impl reference` runs it), the reference itself otherwise"""
    try:
        cores = os.cpu_count() or 1
        u, sec, kind, used, sample, _ = reference_sample(w, g, cores, 1, 0, allow_reference=g.n <= REFERENCE_FULL_JOB_NODES)
        unit = metric_of(w)[1]
        return {"value": u / sec, "unit": unit, "cores": used, "kind": kind, "sample": f"{sample}, {sec:.1f} s"}
    except Exception as e:  # the checker is optional for the product bench
        return {"value": None, "unit": None, "cores": 0, "kind": "unavailable", "sample": repr(e)}


# ------------------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def run_job(w, wname, args, cx, steps, warm, hub_threshold, want_e2e=True, want_cpu=True, sample_clocks=True):
    """One workload on cx.world GPUs: returns the JSON fields of its line (rank 0) or None (other ranks)."""
    import torch
    import approximated_personalized_pagerank_b200 as ppr
    from approximated_personalized_pagerank_b200 import _lib
    dist, rank, world, lib = cx.dist, cx.rank, cx.world, cx.lib
    g = make_graph(w)
    colour = ppr.find_partitions_csr(g) if w["kind"] == "grank" else np.zeros(g.n, dtype=np.uint8)
    stream = torch.cuda.current_stream()
    sess = ppr.Session(g, w["L"], colour=colour, hub_threshold=hub_threshold, rank=rank, world=world, stream=stream.cuda_stream)
    if world > 1:
        from approximated_personalized_pagerank_b200 import multigpu
        multigpu.connect(sess, dist)  # CUDA IPC handles of the basket buffers travel over the process group

    def one_step():
        if w["kind"] == "grank":
            sess.grank(w["K"], w["L"], w["iterations"], w["damping"], w["tolerance"])
        else:
            sess.mc(w["K"], w["L"], w["iterations"], w["damping"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        cx.flush.zero_()
        one_step()
    barrier()
    sampler = ClockSampler(cx.local_rank)
    if rank == 0 and sample_clocks:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    merge_ms, abytes, units, launches, walk_ms, walk_bytes = 0.0, 0, 0, 0, 0.0, 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        cx.flush.zero_()
        ev[i][0].record(stream)
        one_step()
        ev[i][1].record(stream)
        torch.cuda.synchronize()  # per-step counters are read after the step has drained (device-side counters, tiny D2H)
        st = sess.stats()
        merge_ms += sess.kernel_time(0)[1]; abytes += st["algorithmic_bytes"]; launches += sess.launches()
        if w["kind"] == "mc":
            walk_ms += sess.kernel_time(1)[1]; walk_bytes += st["walk_algorithmic_bytes"]
        units += st["node_iterations"] if w["kind"] == "grank" else st["walk_steps"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, merge_ms, walk_ms], dtype=torch.float64, device="cuda")
    stats = sess.stats()
    # basket bytes this rank stored into its peers (last step): a basket goes to the ranks that own a predecessor of its node
    peers_per_basket = float(world - 1)
    if world > 1 and w["kind"] == "grank":
        from approximated_personalized_pagerank_b200 import multigpu
        owner = multigpu.shard_owner(g, colour, hub_threshold, world)
        src_owner = np.repeat(owner, g.out_degree())
        need = np.zeros(g.n, dtype=np.int64)
        for r in range(world):
            m = np.zeros(g.n, dtype=bool)
            m[g.col[src_owner == r]] = True
            need += m & (owner != r)
        mine = owner == rank
        peers_per_basket = float(need[mine].mean()) if mine.any() else 0.0
        del src_owner, need
    pushed = int(stats["nonsink_node_iterations"] * 12 * ((w["L"] + 3) // 4 * 4) * peers_per_basket)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-rank shards of the job -> whole-job totals (node_iterations is already global: colour sizes x iterations)
        tot = torch.tensor([abytes, walk_bytes, launches, units if w["kind"] == "mc" else 0] + [stats[k] for k in COUNTERS],
                           dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        abytes, walk_bytes, launches = int(tot[0]), int(tot[1]), int(tot[2])
        if w["kind"] == "mc":
            units = int(tot[3])
        for i, k in enumerate(COUNTERS):
            stats[k] = int(tot[4 + i])
        mx = torch.tensor([pushed], dtype=torch.int64, device="cuda")
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pushed = int(mx[0])
    dev_ms, merge_ms, walk_ms = (float(x) for x in t.tolist())
    metric, unit = metric_of(w)

    # ---- e2e: host buffers through the one-shot C-ABI, pinned staging, H2D/D2H inside the timed region ----
    e2e = None
    e_steps = max(1, min(steps, 3))
    n, K = g.n, w["K"]
    if want_e2e and (world == 1 or rank == 0):
        import ctypes as C
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
        ids = torch.empty((n, K), dtype=torch.int32).pin_memory().numpy()
        sc = torch.empty((n, K), dtype=torch.float64).pin_memory().numpy()
        cnt = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    if want_e2e and world == 1:
        rp, cl = pin(g.row_ptr), pin(g.col)
        st = _lib.Stats()

        def e2e_step():
            if w["kind"] == "grank":
                _lib.check(lib.pprb200_grank(_lib.ptr(rp), _lib.ptr(cl), n, None, K, w["L"], w["iterations"], w["damping"],
                                             w["tolerance"], hub_threshold, _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
                return st.node_iterations
            _lib.check(lib.pprb200_mccompletepathv2(_lib.ptr(rp), _lib.ptr(cl), n, K, w["L"], w["iterations"], w["damping"],
                                                    ppr.api.DEFAULT_MC_SEED, ppr.api.DEFAULT_MC_ROUNDS, hub_threshold,
                                                    _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
            return st.walk_steps
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u = 0
        for _ in range(e_steps):
            u += e2e_step()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        h2d = g.row_ptr.nbytes + g.col.nbytes + g.n * 5 + 4 * int((g.out_degree() > 0).sum())
        e2e = {"value": u / sec, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n * K * 12 + n * 4),
               "ms_per_step": 1e3 * sec / e_steps, "host_prep_ms": st.prep_ms, "kernel_ms": st.kernel_ms, "d2h_ms": st.d2h_ms,
               "api": "pprb200_grank / pprb200_mccompletepathv2 (one-shot C-ABI, host CSR in, flat baskets out)"}
    elif want_e2e:
        # N GPUs end to end through the same one-shot C-ABI call: PPR_NUM_GPUS = N makes rank 0's process plan once and drive
        # all N devices itself (a session per device, peer access); the other ranks' processes only wait here
        sess.close()
        sess = None
        torch.cuda.empty_cache()
        lib.pprb200_release_cached_memory()
        barrier()
        # (the idle ranks wait on the HOST: an NCCL barrier is a kernel spinning on their GPUs, which rank 0 is about to use)
        dist.barrier(group=cx.cpu_group)
        if rank == 0:
            os.environ["PPR_NUM_GPUS"] = str(world)
            rp, cl = pin(g.row_ptr), pin(g.col)
            st = _lib.Stats()

            def e2e_step():
                if w["kind"] == "grank":
                    _lib.check(lib.pprb200_grank(_lib.ptr(rp), _lib.ptr(cl), n, None, K, w["L"], w["iterations"], w["damping"],
                                                 w["tolerance"], hub_threshold, _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
                    return st.node_iterations
                _lib.check(lib.pprb200_mccompletepathv2(_lib.ptr(rp), _lib.ptr(cl), n, K, w["L"], w["iterations"], w["damping"],
                                                        ppr.api.DEFAULT_MC_SEED, ppr.api.DEFAULT_MC_ROUNDS, hub_threshold,
                                                        _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
                return st.walk_steps
            e2e_step()
            t0 = time.perf_counter()
            u = 0
            for _ in range(e_steps):
                u += e2e_step()
            sec = time.perf_counter() - t0
            lib.pprb200_release_cached_memory()
            os.environ["PPR_NUM_GPUS"] = "1"
            e2e = {"value": u / sec, "unit": unit, "h2d_bytes_per_step": int(world * (g.row_ptr.nbytes + g.col.nbytes + g.n * 5)),
                   "d2h_bytes_per_step": int(n * K * 12 + n * 4), "ms_per_step": 1e3 * sec / e_steps, "n_gpus_used": int(st.n_gpus),
                   "host_prep_ms": st.prep_ms, "kernel_ms": st.kernel_ms, "d2h_ms": st.d2h_ms,
                   "api": f"pprb200_grank / pprb200_mccompletepathv2 with PPR_NUM_GPUS={world}: one process, one host plan, a session per device over peer access"}
        dist.barrier(group=cx.cpu_group)
        barrier()

    if world > 1:
        dist.barrier()
    if sess is not None:
        sess.close()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    peak *= world  # aggregate HBM bandwidth of the GPUs that share the job
    if w["kind"] == "mc":  # dominant kernel of the MC path by BASELINE's metric: the walk kernel (12 B per hop)
        rl_bytes, rl_ms, rl_kernel = walk_bytes, walk_ms, "mc_walk_kernel (12 B/hop + basket writes); combine rounds reported under roofline.combine"
    else:
        rl_bytes, rl_ms, rl_kernel = abytes, merge_ms, ("merge kernels (merge_dense_kernel big/mid, merge_par_kernel for split hubs and hand-overs, "
                                                        "merge_seq_kernel cascade; three streams), all launches of a step; event-timed per iteration on the session stream")
    achieved = (rl_bytes / 1e9) / (rl_ms / 1e3) if rl_ms > 0 else None
    prof = committed_profile(wname) if world == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "traffic": ({"dram_bytes_per_step": prof["dram_bytes_per_step"], "source": prof["source"]} if prof and "dram_bytes_per_step" in prof else None),
                "peak_source": peak_src, "kernel": rl_kernel, "algorithmic_bytes_per_step": rl_bytes // steps,
                "kernel_ms_per_step": rl_ms / steps, "share_of_step": rl_ms / dev_ms if dev_ms > 0 else None}
    if w["kind"] == "mc":
        hops = units
        roofline["combine"] = {"achieved": (abytes / 1e9) / (merge_ms / 1e3) if merge_ms > 0 else None, "unit": "GB/s",
                               "frac": (abytes / 1e9) / (merge_ms / 1e3) / peak if merge_ms > 0 else None,
                               "ms_per_step": merge_ms / steps, "algorithmic_bytes_per_step": abytes // steps}
        # the same walk rate against the sector-granular figure (two 32-byte sectors per hop) and the L2
        roofline["walk_steps_per_s_walk_kernel_only"] = hops / (walk_ms / 1e3) if walk_ms > 0 else None
        roofline["sector_granular"] = {"bytes_per_hop": 64, "achieved": (64.0 * hops / 1e9) / (walk_ms / 1e3) if walk_ms > 0 else None,
                                       "unit": "GB/s", "frac_of_hbm_peak": (64.0 * hops / 1e9) / (walk_ms / 1e3) / peak if walk_ms > 0 else None,
                                       "note": (f"CSR of this graph: {(4 * g.n_edges + 8 * g.n) / 1e6:.0f} MB against 126 MB of L2; "
                                                "the walk is bound by two dependent gathers per hop, not by HBM bandwidth")}
        roofline["l2_throughput"] = prof.get("l2") if prof else None
    out = {
        "metric": metric, "value": units / (dev_ms / 1e3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"], "K": w["K"], "L": w["L"], "iterations": w["iterations"], "damping": w["damping"],
                   "tolerance": w["tolerance"], "iterations_run": stats["iterations_run"], "nodes": g.n, "edges": g.n_edges,
                   "hub_threshold": "default (12)" if hub_threshold == 0 else ("never (exact fma order everywhere)" if hub_threshold == NEVER_HUB else hub_threshold),
                   "l2": "flushed between steps (512 MiB memset); within a step the working set is what it is",
                   "sharding": "single GPU" if world == 1 else f"sources sharded over {world} GPUs (same job: strong scaling)"},
        "wall_ms_per_step_incl_flush_and_stat_reads": wall_ms / steps,
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "counters": {k: stats[k] for k in COUNTERS},
    }
    if world > 1 and w["kind"] == "grank" and stats["iterations_run"]:
        per_it = pushed / stats["iterations_run"]
        it_ms = merge_ms / steps / stats["iterations_run"]
        out["nvlink"] = {"bytes_pushed_per_gpu_per_iteration": int(per_it), "achieved": per_it / 1e9 / (it_ms / 1e3) if it_ms > 0 else None,
                         "peak": 900.0, "unit": "GB/s", "frac": per_it / 1e9 / (it_ms / 1e3) / 900.0 if it_ms > 0 else None,
                         "peers_per_basket": peers_per_basket,
                         "note": "baskets stored by the producing CTA (publish_slot) into the peers that read them (need masks: ranks owning a predecessor), overlapped with the merge; max over ranks; the final completion push is not counted"}
    if world == 1 and want_cpu:
        out["cpu_baseline"] = cpu_baseline(w, g)
    return out


def e2e_api_cpp(wname):
    """e2e through the reference-facing template API itself: tests/cpp/api_bench_b200 times ppr::grank on an
    unordered_map<int, vector<int>> (relabel + C-ABI call + map-of-maps materialisation) and ppr::b200::grankFlat."""
    exe = ROOT / "tests" / "cpp" / "api_bench_b200"
    if not exe.exists():
        return None
    try:
        r = subprocess.run([str(exe), wname], capture_output=True, text=True, timeout=900)
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hub-threshold", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mc", action="store_true", help="skip the MC R-MAT-20 sub-line of the default workload")
    ap.add_argument("--no-exact", action="store_true", help="skip the exact-order second value")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch
    import torch.distributed as dist
    from approximated_personalized_pagerank_b200 import _lib

    cx = Ctx()
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.dist = dist
    cx.inproc_multi = False
    if args.gpus != cx.world and cx.world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={cx.world}")
    torch.cuda.set_device(cx.local_rank)
    cx.cpu_group = None
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local_rank))
        cx.cpu_group = dist.new_group(backend="gloo")  # host-side waits (no kernel on the GPU)
    cx.lib = _lib.load()
    if cx.lib.pprb200_device_count() < 1:
        raise SystemExit("bench.py needs an sm_100 GPU (no CPU fallback)")
    warm = max(3, args.warmup)
    steps = max(1, args.steps)
    cx.flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    line = run_job(w, args.workload, args, cx, steps, warm, args.hub_threshold, want_e2e=not args.no_e2e,
                   want_cpu=not args.no_cpu_baseline)
    default = args.workload == DEFAULT_WORKLOAD
    if default and not args.no_exact and args.hub_threshold == 0:
        # the same job with the reference's exact fma order for every node (bit-identical to the reference's arithmetic)
        ex = run_job(w, args.workload, args, cx, 1, 0, NEVER_HUB, want_e2e=False, want_cpu=False, sample_clocks=False)
        if line is not None and ex is not None:
            line["exact_order"] = {k: ex[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches")}
            line["exact_order"]["roofline_frac"] = ex["roofline"]["frac"]
            line["exact_order"]["config"] = "same workload, hub_threshold = 4294967295: every node on merge_seq_kernel (the reference's fma chain in successor order)"
    if default and not args.no_mc:
        mc = run_job(WORKLOADS["rmat20mc"], "rmat20mc", args, cx, min(steps, 5), 3, args.hub_threshold, want_e2e=not args.no_e2e,
                     want_cpu=not args.no_cpu_baseline, sample_clocks=False)
        if line is not None:
            line["mc"] = mc
    if cx.world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if cx.rank != 0 or line is None:
        return
    if cx.world == 1 and not args.no_e2e:
        del cx.flush
        torch.cuda.empty_cache()
        cx.lib.pprb200_release_cached_memory()  # the C++ program below is another process: give it the GPU's memory
        # (tests/cpp/api_bench.cc builds R-MAT graphs as std::unordered_map and calls ppr::grank: GRank R-MAT workloads only)
        line["e2e_api"] = e2e_api_cpp(args.workload) if (w["kind"] == "grank" and w["gen"] == "rmat") else None
    try:
        line["parity_report"] = json.loads((ROOT / "profiles" / "r2" / "parity_report.json").read_text())
    except Exception:
        line["parity_report"] = None
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
