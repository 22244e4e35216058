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
def test_two_gpus_bit_identical_to_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(ROOT / "tests" / "multigpu_worker.py"), "12"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
