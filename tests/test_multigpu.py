"""Multi-GPU path. CPU part: the host-side shard plan and the handle exchange plumbing under a
world_size-2 gloo group. GPU part (needs >= 2 B200s, `gpurun --gpus 2`): tests/multigpu_worker.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G, multigpu

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_plan_covers_every_non_sink_node_once_and_balances_work(world):
    g = G.rmat(12)
    colour = ppr.find_partitions_csr(g)
    owner = multigpu.shard_owner(g, colour, 0, world)
    deg = g.out_degree()
    assert (owner[deg == 0] == -1).all() and (owner[deg > 0] >= 0).all() and owner.max() == world - 1
    for c in (0, 1):  # per colour (= per iteration) the work is balanced up to the single largest node (LPT)
        work = np.array([deg[(owner == r) & (colour == c)].sum() for r in range(world)], dtype=np.float64)
        assert work.max() <= max(1.05 * work.mean(), deg[colour == c].max() + 0.05 * work.mean())


@pytest.mark.parametrize("world", [2, 3, 8])
def test_need_masks_name_exactly_the_ranks_that_read_a_basket(world):
    """During the iterations a basket is stored only into the ranks that own a predecessor of its node (publish_slot reads
    PeerDev::need); a mask that misses a reader means stale baskets on that rank, one that names too many wastes NVLink stores.
    Specification check of the plan on the CPU, by node: R-MAT (hubs read everywhere), BA, a ring (one reader each), sinks."""
    for g in (G.rmat(11), G.barabasi_albert(3000, 4), G.ring(64), G.from_edges(6, [0, 0, 1, 5], [1, 2, 2, 2])):
        colour = ppr.find_partitions_csr(g)
        owner, need = multigpu.need_mask(g, colour, 0, world)
        assert (owner == multigpu.shard_owner(g, colour, 0, world)).all()
        deg = g.out_degree()
        src = np.repeat(np.arange(g.n), deg)
        want = np.zeros(g.n, dtype=np.int64)
        np.bitwise_or.at(want, g.col, 1 << owner[src].astype(np.int64))   # every edge v -> u: owner(v) reads u's basket
        want[deg == 0] = 0                                                 # sinks have no basket to send
        assert (need.astype(np.int64) == want).all(), int((need != want).sum())
    # one rank: nothing to send
    g = G.rmat(9)
    owner, need = multigpu.need_mask(g, None, 0, 1)
    assert (need == 0).all() and set(np.unique(owner)) <= {-1, 0}


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = G.rmat(10)
    colour = ppr.find_partitions_csr(g)
    owner = multigpu.shard_owner(g, colour, 0, world)
    mine = torch.from_numpy((owner == rank).astype(np.int64))
    tot = mine.clone()
    dist.all_reduce(tot)
    ok = bool((tot.numpy() == (g.out_degree() > 0)).all())
    # the handle exchange pattern of multigpu.connect with a fake 192-byte blob per rank
    blob = torch.full((multigpu.IPC_BYTES,), rank, dtype=torch.uint8)
    out = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(out, blob)
    ok = ok and all(int(out[r][0]) == r and int(out[r][-1]) == r for r in range(world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_world2_gloo_shards_are_disjoint_and_handles_travel():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_multi_rank_session_refuses_to_run_before_attach():
    from approximated_personalized_pagerank_b200 import _lib
    if _lib.load().pprb200_device_count() < 1:
        pytest.skip("needs a GPU to create a session")
    g = G.rmat(8)
    s = ppr.Session(g, 100, rank=0, world=2)
    with pytest.raises(_lib.PprB200Error):
        s.grank(10, 20, 2, 0.85, -1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_gpus_bit_identical_to_the_oracle(world):
    """one process per GPU (torchrun): every rank ends with the same baskets, bit-identical to the single-process oracle --
    GRank and MCCompletePathV2 do not depend on the GPU count (skipped by device count: `gpurun --gpus N`)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29611 + world), str(ROOT / "tests" / "multigpu_worker.py"), "12"], capture_output=True, text=True,
                       timeout=1200)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_multi_gpu_behind_the_one_shot_c_abi(world, monkeypatch):
    """PPR_NUM_GPUS: pprb200_grank / pprb200_mccompletepathv2 shard the sources over N devices inside ONE
    process -- peer access, no IPC, same kernels -- and return the same bits as on one GPU and as the oracle."""
    import torch
    import oracle_bindings as ob
    from helpers import assert_bit_identical
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    g = G.rmat(13)
    colour = ppr.find_partitions_csr(g)
    monkeypatch.setenv("PPR_NUM_GPUS", str(world))
    got = ppr.grank_csr(g, 50, 100, 12, 0.85, 1e-3, colour=colour)
    assert got.stats["n_gpus"] == world
    want = ob.oracle_grank(g, 50, 100, 12, 0.85, 1e-3, colour=colour, hub_threshold=ppr.DEFAULT_HUB_THRESHOLD)
    assert_bit_identical(got, want, f"one process, {world} GPUs, grank")
    for k in ("iterations_run", "merged_entries", "nonsink_node_iterations", "truncations", "boundary_ties"):
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])
    mc = ppr.mccompletepathv2_csr(g, 50, 100, 200, 0.85)
    assert mc.stats["n_gpus"] == world
    wmc = ob.oracle_mc(g, 50, 100, 200, 0.85, ppr.api.DEFAULT_MC_SEED, ppr.api.DEFAULT_MC_ROUNDS, hub_threshold=ppr.DEFAULT_HUB_THRESHOLD)
    assert_bit_identical(mc, wmc, f"one process, {world} GPUs, mc")
    assert mc.stats["walk_steps"] == wmc.stats["walk_steps"]
    monkeypatch.setenv("PPR_NUM_GPUS", "1")
    one = ppr.grank_csr(g, 50, 100, 12, 0.85, 1e-3, colour=colour)
    assert one.stats["n_gpus"] == 1
    assert_bit_identical(got, one, "N GPUs vs one GPU")


@pytest.mark.gpu
def test_drop_in_cpp_program_on_two_gpus_equals_one_gpu():
    """the unchanged template API (ppr::grank on an unordered_map) reaches the multi-GPU path through PPR_NUM_GPUS: the
    drop-in check program prints the same baskets with 2 GPUs as with 1"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = ROOT / "tests" / "cpp" / "dropin_check_b200"
    if not exe.exists():
        pytest.skip("tests/cpp/dropin_check_b200 not built")
    outs = []
    for n in ("1", "2"):
        r = subprocess.run([str(exe), "random_full"], capture_output=True, text=True, timeout=600, env=dict(os.environ, PPR_NUM_GPUS=n))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and len(outs[0]) > 100


@pytest.mark.gpu
def test_a_peer_that_never_attaches_is_reported_not_ignored(monkeypatch):
    """world = 2 session whose peer never runs: the mailbox barrier gives up after PPRB200_PEER_TIMEOUT_MS, the run stops, and
    stats / fetch return PPRB200_ERR_CUDA instead of PPRB200_OK with garbage (both ranks' sessions live on the one GPU of the
    test box and are wired with pprb200_session_attach_local; only rank 0 ever runs)"""
    from approximated_personalized_pagerank_b200 import _lib
    monkeypatch.setenv("PPRB200_PEER_TIMEOUT_MS", "200")
    import ctypes as C
    g = G.rmat(9)
    s = ppr.Session(g, 100, rank=0, world=2)
    s2 = ppr.Session(g, 100, rank=1, world=2)   # the "peer": attached, but its run is never enqueued
    arr = (C.c_void_p * 2)(s.handle, s2.handle)
    _lib.check(s.lib.pprb200_session_attach_local(arr, 2))
    s.grank(10, 20, 4, 0.85, -1.0)   # rank 1 never runs: rank 0's first barrier times out
    with pytest.raises(_lib.PprB200Error, match="peer barrier timed out"):
        s.stats()
    with pytest.raises(_lib.PprB200Error, match="peer barrier timed out"):
        s.fetch()
