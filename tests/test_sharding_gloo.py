"""world_size-2 gloo tests of the multi-GPU host logic (shard ranges, colptr / validity / table
all-gathers, deterministic MC reduction).  Shard contents come from the oracle, so this runs on
the CPU-only box."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp_

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mpb200
    from mpb200 import sharding
    from oracle import oracle as orc
    import fixtures as fx
    N = 4001
    V = fx.uniform_samples(N, 2, 77)
    r = fx.fmt_radius(N, 2)
    q0, q1 = sharding.shard_range(N, rank, world)
    tree = orc.KDTree(V)
    colptr, rowval, nzval = tree.rball(r, q0, q1)                       # this rank's shard
    O = orc.Obstacles2D(fx.ISRR_2H)
    So = orc.StateSpace([0, 0], [1, 1])
    valid, _ = orc.edges_free_csc(O, So, V, colptr, rowval, q0)
    chunks = np.packbits(valid, bitorder="little")
    chunks = np.concatenate([chunks, np.zeros((-len(chunks)) % 8, dtype=np.uint8)]).view(np.uint64)
    gcol, offs = sharding.allgather_colptr(colptr)
    gbits = sharding.allgather_bits(chunks, len(valid))
    grow, gnz = sharding.allgather_table(rowval, nzval)
    # the packed per-step exchange with UNEVEN shards (4001 columns over 2 ranks): capacities agreed collectively,
    # true column counts in the payload
    import torch
    ex = sharding.ValidityExchange(q1 - q0, (len(valid) + 63) // 64)
    ex.run_arrays(torch.from_numpy(colptr.copy()), torch.from_numpy(chunks.view(np.int64).copy()))
    xcol, xbits = ex.assemble()
    mc = sharding.allreduce_mc(dict(S1=0.1 * (rank + 1), S2=0.01 * (rank + 1), S0=10.0 + rank, n=10, hits=rank + 1))
    # reference: the unsharded oracle
    fc, fr, fz = tree.rball(r)
    fv, _ = orc.edges_free_csc(O, So, V, fc, fr)
    ok = (np.array_equal(gcol, fc) and np.array_equal(xcol, fc) and np.array_equal(xbits, gbits) and np.array_equal(grow, fr) and gnz.tobytes() == fz.tobytes()
          and np.array_equal(np.unpackbits(gbits.view(np.uint8), bitorder="little")[:len(fv)], fv)
          and offs[rank] == fc[q0] - 1 and mc["n"] == 10 * world and mc["hits"] == sum(range(1, world + 1))
          and abs(mc["S1"] - 0.1 * sum(range(1, world + 1))) < 1e-15)
    open(os.path.join(tmp, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    from mpb200 import sharding
    for N, G in ((10, 3), (1_000_000, 8), (5, 8), (0, 2)):
        rs = [sharding.shard_range(N, g, G) for g in range(G)]
        assert rs[0][0] == 0 and rs[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def test_concat_bitvectors():
    sys.path.insert(0, ROOT)
    from mpb200 import sharding
    rng = np.random.default_rng(0)
    parts, allbits = [], []
    for n in (0, 1, 63, 64, 65, 200, 0, 7):
        b = rng.integers(0, 2, n).astype(np.uint8)
        ch = np.packbits(b, bitorder="little")
        ch = np.concatenate([ch, np.zeros((-len(ch)) % 8, dtype=np.uint8)]).view(np.uint64)
        if n % 64 and len(ch):
            ch = ch.copy()
            ch[-1] |= np.uint64(0xFFFFFFFFFFFFFFFF) << np.uint64(n % 64)   # garbage above nbits must be masked
        parts.append((ch, n))
        allbits.append(b)
    out = sharding.concat_bitvectors(parts)
    exp = np.concatenate(allbits)
    got = np.unpackbits(out.view(np.uint8), bitorder="little")
    assert np.array_equal(got[:len(exp)], exp) and not got[len(exp):].any()


def test_world2_gloo_allgathers(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp_.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("ok%d" % r)).read() == "1"


def test_validity_exchange_wire_format_roundtrip():
    """ValidityExchange packs [int64 ncols | int32 column lengths | uint64 validity words] per rank; assemble() must
    rebuild the global colptr and the concatenated BitVector from the gathered buffer, with a different column
    count on every rank (host logic, no GPU/NCCL)."""
    import torch
    from mpb200 import sharding
    rng = np.random.Generator(np.random.PCG64(5))
    world, cap_cols = 3, 7                               # odd capacity: the int32 block is padded to 8 bytes
    ncols_of = [7, 6, 4]
    counts = [rng.integers(0, 40, size=n).astype(np.int64) for n in ncols_of]
    nnz = [int(c.sum()) for c in counts]
    cap = (max(nnz) + 63) // 64 + 2
    ex = sharding.ValidityExchange.__new__(sharding.ValidityExchange)
    ex.world, ex.ncols, ex.cap = world, cap_cols, cap
    ex.cnt_words = (cap_cols + 1) // 2
    ex.stride = 1 + ex.cnt_words + cap
    buf = np.zeros((world, ex.stride), dtype=np.int64)
    bits = []
    for g in range(world):
        buf[g, 0] = ncols_of[g]
        buf[g, 1:1 + ex.cnt_words].view(np.int32)[:ncols_of[g]] = counts[g]
        buf[g, 1:1 + ex.cnt_words].view(np.int32)[ncols_of[g]:] = 99        # stale padding must be ignored
        b = rng.integers(0, 2, size=nnz[g]).astype(np.uint8)
        bits.append(b)
        words = np.packbits(np.concatenate([b, np.zeros((-len(b)) % 64, np.uint8)]), bitorder="little").view(np.uint64)
        buf[g, 1 + ex.cnt_words:1 + ex.cnt_words + len(words)] = words.view(np.int64)
    ex.recv = torch.from_numpy(buf.reshape(-1).copy())
    colptr, chunks = ex.assemble()
    exp_colptr = np.concatenate([[1], 1 + np.cumsum(np.concatenate(counts))])
    assert np.array_equal(colptr, exp_colptr)
    allbits = np.concatenate(bits)
    got = np.unpackbits(chunks.view(np.uint8), bitorder="little")[:len(allbits)]
    assert np.array_equal(got, allbits)
