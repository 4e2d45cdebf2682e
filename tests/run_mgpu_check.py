"""Launched by torchrun (one rank per GPU): shard the C2-style precompute by query range, run the
NCCL exchange, and check the assembled global colptr / validity bits and the MC reduction against
the oracle on rank 0.  Prints MGPU_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mpb200
    from mpb200 import _lib, sharding
    import fixtures as fx
    lib = mpb200.init(local)
    stream = torch.cuda.Stream()                           # one explicit stream for the library AND torch (as bench.py)
    torch.cuda.set_stream(stream)
    _lib.check(lib.mpb200_set_stream(_lib.c_vp(stream.cuda_stream)))
    N = 200_001                                            # uneven shards on purpose
    V = fx.uniform_samples(N, 2, 4242)
    V = V[np.argsort(V[:, 0], kind="stable")]              # stripe order, as bench.py shards
    r = fx.fmt_radius(N, 2)
    q0, q1 = sharding.shard_range(N, rank, world)
    NN = mpb200.MetricNN(V)
    NN.set_query_range(q0, q1)
    CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H())
    SS = mpb200.UnitHypercube(2)
    ok = True
    results = {}
    for kind in ("peer", "nccl"):                          # direct peer stores (csrc/xchg.cu) and the NCCL all-gather form
        nnz = NN.build_table(r)
        NN.edges_free(NN.table, CC, SS, fetch=False)
        ex = sharding.make_exchange(q1 - q0, (nnz + 63) // 64 + 1, kind=kind)
        if kind == "peer":
            ex.attach(NN.table)                            # column lengths travel early, under the fill
        for _ in range(3):                                 # several epochs: the alternating buffer sets and the flag barrier
            NN.build_table(r)
            NN.edges_free(NN.table, CC, SS, fetch=False, count=False)
            ex.run(NN.table)
        torch.cuda.synchronize()
        results[kind] = ex.assemble()
        if hasattr(ex, "close"):
            ex.close()
    gcol, gbits = results["peer"]
    ok = ok and np.array_equal(gcol, results["nccl"][0]) and np.array_equal(gbits, results["nccl"][1])
    # MC: shard the rollout ids, reduce in rank order
    Bx = mpb200.PointRobotNDBoxes([mpb200.BoxBounds(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    P = mpb200.MCProblem(np.eye(2)[None], (np.eye(2) * 0.1)[None], np.eye(2), np.array([[0.2, 0.0]] * 2), [0.3, 0.7],
                         [[3.0, 0.0]])
    n_tot = 400_000
    a, b = sharding.shard_range(n_tot, rank, world)
    mc = sharding.allreduce_mc(mpb200.collision_probability(P, Bx, b - a, seed=9, first=a))
    if rank == 0:
        from oracle import oracle as orc
        fc, fr, fz = orc.KDTree(V).rball(r)
        fv, _ = orc.edges_free_csc(orc.Obstacles2D(fx.ISRR_2H), orc.StateSpace([0, 0], [1, 1]), V, fc, fr)
        got = np.unpackbits(gbits.view(np.uint8), bitorder="little")[:len(fv)]
        ok = ok and np.array_equal(gcol, fc) and np.array_equal(got, fv)
        whole = mpb200.collision_probability(P, Bx, n_tot, seed=9)
        ok = ok and mc["hits"] == whole["hits"] and abs(mc["S1"] - whole["S1"]) <= 1e-12 * whole["S1"]
        print("MGPU_OK" if ok else "MGPU_FAIL", "world", world, "nnz", len(fv), "mc_p", mc["p"], flush=True)
    NN.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
