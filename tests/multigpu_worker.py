"""Worker of tests/test_multigpu.py: run under torchrun, one rank per GPU.

Checks, for GRank and MCCompletePathV2 on R-MAT graphs: every rank ends with the same baskets, and they are
bit-identical to the single-process oracle (hence to the 1-GPU run: the result does not depend on the GPU count)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import approximated_personalized_pagerank_b200 as ppr  # noqa: E402
from approximated_personalized_pagerank_b200 import graphs as G, multigpu  # noqa: E402
import oracle_bindings as ob  # noqa: E402
from helpers import assert_bit_identical  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    for hub in (0, 16):
        g = G.rmat(scale)
        colour = ppr.find_partitions_csr(g)
        sess = ppr.Session(g, 100, colour=colour, hub_threshold=hub, rank=rank, world=world)
        multigpu.connect(sess, dist)
        for (K, L, it, tol) in ((50, 100, 30, 1e-3), (10, 40, 7, -1.0)):
            sess.grank(K, L, it, 0.85, tol)
            got = sess.fetch()
            got.stats = sess.stats()
            want = ob.oracle_grank(g, K, L, it, 0.85, tol, colour=colour, hub_threshold=hub if hub else ppr.DEFAULT_HUB_THRESHOLD)
            assert_bit_identical(got, want, f"rank {rank}/{world} grank rmat{scale} hub {hub} K{K}")
            assert got.stats["iterations_run"] == want.stats["iterations_run"]
            t = torch.tensor([got.stats["merged_entries"], got.stats["nonsink_node_iterations"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            assert int(t[0]) == want.stats["merged_entries"] and int(t[1]) == want.stats["nonsink_node_iterations"], \
                "the ranks' shards do not add up to the whole job"
        for (K, L, R, rounds) in ((50, 100, 200, 3), (20, 30, 50, 0)):
            sess.mc(K, L, R, 0.85, rounds=rounds)
            got = sess.fetch()
            st = sess.stats()
            want = ob.oracle_mc(g, K, L, R, 0.85, ppr.api.DEFAULT_MC_SEED, rounds, hub_threshold=hub if hub else ppr.DEFAULT_HUB_THRESHOLD)
            assert_bit_identical(got, want, f"rank {rank}/{world} mc rmat{scale} hub {hub} R{R}")
            t = torch.tensor([st["walk_steps"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            assert int(t[0]) == want.stats["walk_steps"]
        torch.cuda.synchronize()
        dist.barrier()
        sess.close()
    if rank == 0:
        print(f"multigpu_worker: world {world} OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
