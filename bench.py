#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 GRank / MCCompletePathV2 hot paths.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rmat16|rmat22|rmat20mc|ring|ba8m] [--impl reference]

One "step" = one whole job of the hot path on one synthetic graph (GRank: init + `iterations` merge sweeps +
final top-K; MC: walks + combine rounds + top-K). Default workload = BASELINE.json configs[1]: GRank on R-MAT
scale 16 (65 536 nodes, 1 048 576 edges), K=50, L=100, 30 iterations, damping 0.85, tolerance 1e-3.

  value      node-iterations/s (walk-steps/s for MC) with graph + baskets resident in HBM (session API),
             timed with CUDA events on the launching stream, L2 flushed between steps
  e2e        same metric through the host-buffer C-ABI call (pprb200_grank): host preprocessing, H2D of the
             CSR from pinned memory and D2H of the baskets are inside the timed region
  roofline   merge kernels only: algorithmic bytes (SURVEY.md 8d formula, counted by the kernels) / device time
             of the merge launches, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the reference's own grankMulti (oracle/_ref, all host cores) on a bounded sample (fewer iterations)

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libppr_ref.so; the oracle port if that is absent) on
the same graph and parameters with fewer iterations per step, and prints the same JSON line with "impl":"reference".
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

WORKLOADS = {
    # name: (kind, generator, args, K, L, iterations/R, damping, tolerance)
    "ring": dict(kind="grank", gen="ring", scale=100, K=50, L=100, iterations=30, damping=0.85, tolerance=1e-3,
                 desc="grank on README's 100-node ring (BASELINE configs[0])"),
    "rmat16": dict(kind="grank", gen="rmat", scale=16, K=50, L=100, iterations=30, damping=0.85, tolerance=1e-3,
                   desc="GRank on R-MAT scale 16 (65536 nodes, 1048576 edges), K=50 L=100 30 it d=0.85 tol=1e-3 (BASELINE configs[1])"),
    "rmat18": dict(kind="grank", gen="rmat", scale=18, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 18"),
    "rmat20": dict(kind="grank", gen="rmat", scale=20, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 20"),
    "rmat22": dict(kind="grank", gen="rmat", scale=22, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                   desc="GRank on R-MAT scale 22 (4194304 nodes, 67108864 edges), K=50 L=100 30 it (BASELINE configs[3])"),
    "rmat20mc": dict(kind="mc", gen="rmat", scale=20, K=50, L=100, iterations=1000, damping=0.85, tolerance=0.0,
                     desc="MCCompletePathV2 on R-MAT scale 20, K=50 L=100 R=1000 d=0.85 (BASELINE configs[2])"),
    "rmat16mc": dict(kind="mc", gen="rmat", scale=16, K=50, L=100, iterations=1000, damping=0.85, tolerance=0.0,
                     desc="MCCompletePathV2 on R-MAT scale 16, K=50 L=100 R=1000 d=0.85"),
    "ba8m": dict(kind="grank", gen="ba", scale=8388608, K=50, L=100, iterations=30, damping=0.85, tolerance=-1.0,
                 desc="GRank on Barabasi-Albert 8M nodes m=4 symmetrised (BASELINE configs[4])"),
}
COUNTERS = ("nonsink_node_iterations", "edge_reads", "merged_entries", "candidates", "truncations", "boundary_ties",
            "overflow_requeues", "walk_steps", "walks")
REFERENCE_SAMPLE_ITERATIONS = 4  # iterations per reference step (bounded sample of the 30-iteration job)


def make_graph(w):
    from approximated_personalized_pagerank_b200 import graphs as G
    if w["gen"] == "ring":
        return G.ring(w["scale"])
    if w["gen"] == "rmat":
        return G.rmat(w["scale"])
    if w["gen"] == "ba":
        return G.barabasi_albert(w["scale"], 4)
    raise ValueError(w["gen"])


def measured_traffic(workload, steps):
    """DRAM bytes of the dominant kernels per step from the committed ncu capture of this command (profiles/traffic.json)"""
    p = ROOT / "profiles" / "traffic.json"
    try:
        t = json.loads(p.read_text()).get(workload)
        return None if t is None else {"dram_bytes_per_step": t["dram_bytes_per_step"], "source": t["source"]}
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


def node_iterations_of(colour, iterations_run):
    c0 = int((colour == 0).sum()); c1 = int((colour == 1).sum())
    return sum(c0 if (i & 1) == 0 else c1 for i in range(iterations_run))


def run_reference(args, w):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_bindings as ob
    g = make_graph(w)
    cores = os.cpu_count() or 1
    it = min(w["iterations"], REFERENCE_SAMPLE_ITERATIONS) if w["kind"] == "grank" else w["iterations"]
    kind = "reference" if ob.have_ref() else "port"
    times, units = [], []
    colour = ob.oracle_find_partitions(g.relabel(ob.ref_iteration_order(g))) if (kind == "reference" and w["kind"] == "grank") else None
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))  # one warm-up is plenty for a CPU path measured in seconds
    for s in range(warm + steps):
        if w["kind"] == "grank":
            if kind == "reference":
                r = ob.ref_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
                sec = r.seconds
                u = node_iterations_of(colour, it)  # the sample's tolerance never triggers within 4 iterations on these graphs
            else:
                t0 = time.perf_counter()
                r = ob.oracle_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
                sec = time.perf_counter() - t0
                u = r.stats["node_iterations"]
        else:
            if kind == "reference":
                r = ob.ref_mc(g, w["K"], w["L"], w["iterations"], w["damping"])
                sec = r.seconds
                u = None
            else:
                t0 = time.perf_counter()
                r = ob.oracle_mc(g, w["K"], w["L"], w["iterations"], w["damping"], 1, 3, nthreads=cores)
                sec = time.perf_counter() - t0
                u = r.stats["walk_steps"]
        if s >= warm:
            times.append(sec); units.append(u)
    metric, unit = ("grank_node_iterations_per_s", "node-iterations/s") if w["kind"] == "grank" else ("mc_walk_steps_per_s", "walk-steps/s")
    if units[0] is None:  # reference MC does not count its hops: use the expected hops of the same walk budget from the port
        units = [ob.oracle_mc(g, w["K"], w["L"], w["iterations"], w["damping"], 1, 0, nthreads=cores).stats["walk_steps"]] * len(times)
    value = sum(units) / sum(times)
    sample = (f"{'grankMulti' if kind == 'reference' else 'oracle port'} nThreads={cores}, first {it} of {w['iterations']} iterations per step"
              if w["kind"] == "grank" else f"full mccompletepathv2 run R={w['iterations']} (single-threaded by construction)")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": w["desc"]},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores if w["kind"] == "grank" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(w, g):
    """bounded sample of the same workload on the host cores (rank 0, N=1 only)"""
    sys.path.insert(0, str(ROOT / "tests"))
    try:
        import oracle_bindings as ob
        cores = os.cpu_count() or 1
        if w["kind"] == "grank":
            it = min(w["iterations"], REFERENCE_SAMPLE_ITERATIONS)
            big = g.n > (1 << 18)
            if ob.have_ref() and not big:
                colour = ob.oracle_find_partitions(g.relabel(ob.ref_iteration_order(g)))
                r = ob.ref_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
                return {"value": node_iterations_of(colour, it) / r.seconds, "unit": "node-iterations/s", "cores": cores, "kind": "reference",
                        "sample": f"grankMulti nThreads={cores}, first {it} of {w['iterations']} iterations, {r.seconds:.1f} s"}
            it = 2 if big else it
            t0 = time.perf_counter()
            r = ob.oracle_grank(g, w["K"], w["L"], it, w["damping"], w["tolerance"], nthreads=cores)
            sec = time.perf_counter() - t0
            return {"value": r.stats["node_iterations"] / sec, "unit": "node-iterations/s", "cores": cores, "kind": "port",
                    "sample": f"OpenMP oracle port, {cores} threads, first {it} of {w['iterations']} iterations, {sec:.1f} s"}
        t0 = time.perf_counter()
        if ob.have_ref() and g.n <= (1 << 16):
            r = ob.ref_mc(g, w["K"], w["L"], w["iterations"], w["damping"])
            steps = ob.oracle_mc(g, w["K"], w["L"], w["iterations"], w["damping"], 1, 0, nthreads=cores).stats["walk_steps"]
            return {"value": steps / r.seconds, "unit": "walk-steps/s", "cores": 1, "kind": "reference",
                    "sample": f"full mccompletepathv2 R={w['iterations']}, {r.seconds:.1f} s; hops counted by the port at the same walk budget"}
        sub = min(w["iterations"], 50)
        r = ob.oracle_mc(g, w["K"], w["L"], sub, w["damping"], 1, 0, nthreads=cores)
        sec = time.perf_counter() - t0
        return {"value": r.stats["walk_steps"] / sec, "unit": "walk-steps/s", "cores": cores, "kind": "port",
                "sample": f"OpenMP oracle port walks only, R={sub}, {sec:.1f} s"}
    except Exception as e:  # the checker is optional for the product bench
        return {"value": None, "unit": None, "cores": 0, "kind": "unavailable", "sample": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="rmat16", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hub-threshold", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch
    import torch.distributed as dist
    import approximated_personalized_pagerank_b200 as ppr
    from approximated_personalized_pagerank_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    if lib.pprb200_device_count() < 1:
        raise SystemExit("bench.py needs an sm_100 GPU (no CPU fallback)")
    warm = max(3, args.warmup)
    steps = max(1, args.steps)

    g = make_graph(w)
    colour = ppr.find_partitions_csr(g) if w["kind"] == "grank" else np.zeros(g.n, dtype=np.uint8)
    stream = torch.cuda.current_stream()
    sess = ppr.Session(g, w["L"], colour=colour, hub_threshold=args.hub_threshold, rank=rank, world=world,
                       stream=stream.cuda_stream)
    if world > 1:
        from approximated_personalized_pagerank_b200 import multigpu
        multigpu.connect(sess, dist)  # CUDA IPC handles of the basket buffers travel over the process group
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

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
        flush.zero_()
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    merge_ms, merge_launches, abytes, units, launches, walk_ms, walk_bytes = 0.0, 0, 0, 0, 0, 0.0, 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record(stream)
        one_step()
        ev[i][1].record(stream)
        # per-step counters are read after the step has drained (device-side counters, tiny D2H)
        torch.cuda.synchronize()
        st = sess.stats()
        l, ms = sess.kernel_time(0)
        merge_ms += ms; merge_launches += l; abytes += st["algorithmic_bytes"]; launches += sess.launches()
        if w["kind"] == "mc":
            walk_ms += sess.kernel_time(1)[1]; walk_bytes += st["walk_algorithmic_bytes"]
        units += st["node_iterations"] if w["kind"] == "grank" else st["walk_steps"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, merge_ms, walk_ms], dtype=torch.float64, device="cuda")
    stats = sess.stats()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-rank shards of the job -> whole-job totals (node_iterations is already global: colour sizes x iterations)
        tot = torch.tensor([abytes, walk_bytes, launches, units if w["kind"] == "mc" else 0] +
                           [stats[k] for k in COUNTERS], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        abytes, walk_bytes, launches = int(tot[0]), int(tot[1]), int(tot[2])
        if w["kind"] == "mc":
            units = int(tot[3])
        for i, k in enumerate(COUNTERS):
            stats[k] = int(tot[4 + i])
    dev_ms, merge_ms, walk_ms = (float(x) for x in t.tolist())

    # ---- e2e: host buffers through the one-shot C-ABI, pinned staging, H2D/D2H inside the timed region ----
    e2e = None
    if not args.no_e2e and world == 1:
        import ctypes as C
        n, K = g.n, w["K"]
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
        rp, cl = pin(g.row_ptr), pin(g.col)
        ids = torch.empty((n, K), dtype=torch.int32).pin_memory().numpy()
        sc = torch.empty((n, K), dtype=torch.float64).pin_memory().numpy()
        cnt = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        st = _lib.Stats()

        def e2e_step():
            if w["kind"] == "grank":
                _lib.check(lib.pprb200_grank(_lib.ptr(rp), _lib.ptr(cl), n, None, K, w["L"], w["iterations"], w["damping"],
                                             w["tolerance"], args.hub_threshold, _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
                return st.node_iterations
            _lib.check(lib.pprb200_mccompletepathv2(_lib.ptr(rp), _lib.ptr(cl), n, K, w["L"], w["iterations"], w["damping"],
                                                    ppr.api.DEFAULT_MC_SEED, ppr.api.DEFAULT_MC_ROUNDS, args.hub_threshold,
                                                    _lib.ptr(ids), _lib.ptr(sc), _lib.ptr(cnt), C.byref(st)))
            return st.walk_steps
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u = 0
        e_steps = max(1, min(steps, 3))
        for _ in range(e_steps):
            u += e2e_step()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        h2d = g.row_ptr.nbytes + g.col.nbytes + g.n * 5 + 4 * int((g.out_degree() > 0).sum())
        e2e = {"value": u / sec, "unit": "node-iterations/s" if w["kind"] == "grank" else "walk-steps/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n * K * 12 + n * 4), "ms_per_step": 1e3 * sec / e_steps,
               "host_prep_ms": st.prep_ms, "kernel_ms": st.kernel_ms, "d2h_ms": st.d2h_ms}

    if world > 1 and not args.no_e2e:
        # N GPUs end to end through the public multi-GPU API: every rank preprocesses + uploads the CSR from host memory,
        # the ranks exchange IPC handles, run, and rank 0 reads the whole result back
        from approximated_personalized_pagerank_b200 import multigpu
        n, K = g.n, w["K"]
        ids = torch.empty((n, K), dtype=torch.int32).pin_memory().numpy()
        sc = torch.empty((n, K), dtype=torch.float64).pin_memory().numpy()
        cnt = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)

        def e2e_multi():
            s2 = ppr.Session(g, w["L"], colour=None if w["kind"] == "grank" else np.zeros(n, dtype=np.uint8),
                             hub_threshold=args.hub_threshold, rank=rank, world=world, stream=stream.cuda_stream)
            multigpu.connect(s2, dist)
            if w["kind"] == "grank":
                s2.grank(K, w["L"], w["iterations"], w["damping"], w["tolerance"])
            else:
                s2.mc(K, w["L"], w["iterations"], w["damping"])
            if rank == 0:
                s2.fetch(ids, sc, cnt)
            st2 = s2.stats()
            torch.cuda.synchronize()
            dist.barrier()
            s2.close()
            return st2["node_iterations"] if w["kind"] == "grank" else st2["walk_steps"]
        e2e_multi()
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(steps, 3))
        u = sum(e2e_multi() for _ in range(e_steps))
        barrier()
        sec = time.perf_counter() - t0
        tu = torch.tensor([u if w["kind"] == "mc" else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(tu)
        u = int(tu[0]) if w["kind"] == "mc" else u
        e2e = {"value": u / sec, "unit": "node-iterations/s" if w["kind"] == "grank" else "walk-steps/s",
               "h2d_bytes_per_step": int(world * (g.row_ptr.nbytes + g.col.nbytes + g.n * 5)), "d2h_bytes_per_step": int(n * K * 12 + n * 4),
               "ms_per_step": 1e3 * sec / e_steps}

    def finish():
        if world > 1:
            dist.barrier()
            sess.close()
            dist.destroy_process_group()

    if rank != 0:
        finish()
        return
    peak, peak_src = measured_peak()
    peak *= world  # aggregate HBM bandwidth of the GPUs that share the job
    if w["kind"] == "mc":  # dominant kernel of the MC path by BASELINE's metric: the walk kernel (12 B per hop)
        rl_bytes, rl_ms, rl_kernel = walk_bytes, walk_ms, "mc_walk_kernel (12 B/hop + basket writes); combine rounds reported under roofline.combine"
    else:
        rl_bytes, rl_ms, rl_kernel = abytes, merge_ms, ("merge kernels (merge_par_kernel big/mid + merge_seq_kernel cascade, on three streams), "
                                                        "all launches of a step; event-timed per iteration on the session stream")
    achieved = (rl_bytes / 1e9) / (rl_ms / 1e3) if rl_ms > 0 else None
    metric, unit = ("grank_node_iterations_per_s", "node-iterations/s") if w["kind"] == "grank" else ("mc_walk_steps_per_s", "walk-steps/s")
    line = {
        "metric": metric, "value": units / (dev_ms / 1e3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"], "K": w["K"], "L": w["L"], "iterations": w["iterations"], "damping": w["damping"],
                   "tolerance": w["tolerance"], "iterations_run": stats["iterations_run"], "nodes": g.n, "edges": g.n_edges,
                   "l2": "flushed between steps (512 MiB memset); within a step the working set is what it is",
                   "sharding": "single GPU" if world == 1 else f"sources sharded over {world} GPUs"},
        "wall_ms_per_step_incl_flush_and_stat_reads": wall_ms / steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "traffic": (measured_traffic(args.workload, steps) if world == 1 else None), "peak_source": peak_src, "kernel": rl_kernel,
                     "algorithmic_bytes_per_step": rl_bytes // steps, "kernel_ms_per_step": rl_ms / steps,
                     "share_of_step": rl_ms / dev_ms if dev_ms > 0 else None,
                     "combine": ({"achieved": (abytes / 1e9) / (merge_ms / 1e3) if merge_ms > 0 else None, "unit": "GB/s",
                                  "ms_per_step": merge_ms / steps, "algorithmic_bytes_per_step": abytes // steps}
                                 if w["kind"] == "mc" else None)},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "counters": {k: stats[k] for k in COUNTERS},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, g)
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
