"""Multi-GPU sharding of the precompute: one process per GPU, query-range shards, replicated
samples and obstacles, and the exchange steps the host planner needs.

The reference has no distributed layer at all (single Julia thread); this is the B200-native
analogue of its only scaling axis, the sample count N.  Rank g owns the contiguous query
(column) range shard_range(N, g, G) of every table; the per-rank CSC shards concatenate in rank
order.  Collectives (torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests):
  * all-gather of per-column counts            -> the global colptr
  * all-gather of edge-validity words          -> the global BitVector (shards re-aligned by shift)
  * all-gather of rowval / nzval (optional)    -> the full table on every rank
  * all-gather + fixed-order sum of the MC sums (deterministic, unlike a tree all-reduce)
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(N, rank, world):
    """contiguous, balanced column range of `rank` (0-based half-open)"""
    base, rem = divmod(int(N), int(world))
    q0 = rank * base + min(rank, rem)
    return q0, q0 + base + (1 if rank < rem else 0)


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _gather_var(x, group=None):
    """all-gather of 1-D tensors of different lengths: sizes first, then padded payloads"""
    world = dist.get_world_size(group)
    n = torch.tensor([x.numel()], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=x.dtype, device=x.device)
    pad[:x.numel()] = x
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def allgather_colptr(local_colptr, group=None):
    """local 1-based colptr (ncols_g + 1) of every rank -> global 1-based colptr (N + 1), and the
    0-based entry offset of every rank's shard"""
    t = torch.as_tensor(np.diff(np.asarray(local_colptr)), dtype=torch.int64).to(_dev())
    parts = [p.cpu().numpy() for p in _gather_var(t, group)]
    counts = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
    colptr = np.empty(len(counts) + 1, dtype=np.int64)
    colptr[0] = 1
    np.cumsum(counts, out=colptr[1:])
    colptr[1:] += 1
    offs = np.concatenate([[0], np.cumsum([int(p.sum()) for p in parts])])
    return colptr, offs


def concat_bitvectors(parts):
    """[(uint64 chunks, nbits), ...] -> chunks of the concatenated BitVector (Julia layout)."""
    total = sum(n for _, n in parts)
    out = np.zeros((total + 63) // 64, dtype=np.uint64)
    pos = 0
    for chunks, n in parts:
        if n == 0:
            continue
        w = np.ascontiguousarray(chunks[:(n + 63) // 64], dtype=np.uint64).copy()
        if n % 64:
            w[-1] &= np.uint64((1 << (n % 64)) - 1)
        word, sh = divmod(pos, 64)
        if sh == 0:
            out[word:word + len(w)] |= w
        else:
            out[word:word + len(w)] |= w << np.uint64(sh)
            hi = w >> np.uint64(64 - sh)
            k = min(len(hi), len(out) - word - 1)
            out[word + 1:word + 1 + k] |= hi[:k]
        pos += n
    return out


def allgather_bits(local_chunks, nbits, group=None):
    """edge-validity words of every rank -> the global BitVector chunks (on every rank)"""
    x = torch.as_tensor(np.ascontiguousarray(local_chunks[:(nbits + 63) // 64]).view(np.int64)).to(_dev())
    words = _gather_var(x, group)
    n = torch.tensor([nbits], dtype=torch.int64, device=x.device)
    ns = [torch.zeros_like(n) for _ in range(dist.get_world_size(group))]
    dist.all_gather(ns, n, group=group)
    return concat_bitvectors([(w.cpu().numpy().view(np.uint64), int(k)) for w, k in zip(words, ns)])


def allgather_table(rowval, nzval, group=None):
    """variable-size all-gather of the shard's rowval / nzval -> full arrays on every rank"""
    rv = _gather_var(torch.as_tensor(np.asarray(rowval), dtype=torch.int64).to(_dev()), group)
    nz = _gather_var(torch.as_tensor(np.asarray(nzval), dtype=torch.float64).to(_dev()), group)
    return np.concatenate([p.cpu().numpy() for p in rv]), np.concatenate([p.cpu().numpy() for p in nz])


def allreduce_mc(res, group=None):
    """sum the Monte-Carlo shard sums in rank order (bit-reproducible for a fixed world size)"""
    x = torch.tensor([res["S1"], res["S2"], res["S0"], float(res["n"]), float(res["hits"])], dtype=torch.float64,
                     device=_dev())
    outs = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, x, group=group)
    tot = dict(S1=0.0, S2=0.0, S0=0.0, n=0, hits=0)
    for o in outs:                                     # fixed order
        o = o.cpu().tolist()
        tot["S1"] += o[0]; tot["S2"] += o[1]; tot["S0"] += o[2]; tot["n"] += int(o[3]); tot["hits"] += int(o[4])
    n = max(tot["n"], 1)
    p = tot["S1"] / n
    tot["p"], tot["se"] = p, (max(tot["S2"] / n - p * p, 0.0) / n) ** 0.5
    return tot


# ---- device-resident exchange (NCCL): no host round trip ----------------------------------------
class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, n, typestr):
    """torch view (no copy) of a device array owned by libmpb200: typestr "<i8", "<f8", "<u8" ..."""
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=torch.device("cuda", torch.cuda.current_device()))


def table_device_tensors(table):
    """(colptr, rowval, nzval, edge_words) torch views of a DeviceTable's arrays"""
    import ctypes
    from . import _lib
    cp, rv, nz, eb = _lib.c_vp(), _lib.c_vp(), _lib.c_vp(), _lib.c_vp()
    _lib.check(_lib.lib().mpb200_table_device_view(table.h, ctypes.byref(cp), ctypes.byref(rv), ctypes.byref(nz),
                                                   ctypes.byref(eb)))
    words = (table.nnz + 63) // 64
    return (device_tensor(cp.value, table.ncols + 1, "<i8"),
            device_tensor(rv.value, table.nnz, "<i8") if table.nnz else None,
            device_tensor(nz.value, table.nnz, "<f8") if table.nnz else None,
            device_tensor(eb.value, words, "<i8") if (eb.value and words) else None)


def _assemble(world, slots_counts, slots_words, ncols_of):
    """global 1-based colptr + BitVector chunks from per-rank (int32 counts, uint64 words)"""
    counts = [np.asarray(slots_counts[g][:ncols_of[g]], dtype=np.int64) for g in range(world)]
    allc = np.concatenate(counts) if counts else np.zeros(0, dtype=np.int64)
    colptr = np.empty(len(allc) + 1, dtype=np.int64)
    colptr[0] = 1
    np.cumsum(allc, out=colptr[1:])
    colptr[1:] += 1
    parts = [(slots_words[g], int(counts[g].sum())) for g in range(world)]
    return colptr, concat_bitvectors(parts)


class ValidityExchange:
    """NCCL / gloo form of the per-step exchange the host planner needs from every GPU (the fallback when CUDA IPC
    peer mapping is unavailable, and the form the CPU tests exercise): the shard column lengths and the
    edge-validity words, packed into ONE all-gather per step: [int64 ncols | int32 counts x max_ncols | uint64 words
    x cap] per rank.  Shards may have different column counts (N % world != 0): the capacities are agreed
    collectively (all-reduce MAX) at construction and each rank's true ncols travels in the payload."""

    def __init__(self, ncols, word_capacity, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        dev = _dev()
        caps = torch.tensor([int(ncols), int(word_capacity)], dtype=torch.int64, device=dev)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)   # every rank enters this: no rank can run ahead alone
        self.ncols = int(caps[0])                                   # capacity in columns (max over ranks)
        self.cap = int(caps[1])
        self.cnt_words = (self.ncols + 1) // 2            # int32 counts, padded to whole 8-byte words
        self.stride = 1 + self.cnt_words + self.cap       # int64 words per rank
        self.send = torch.zeros(self.stride, dtype=torch.int64, device=dev)
        self.recv = torch.empty(self.world * self.stride, dtype=torch.int64, device=dev)

    def run_arrays(self, colptr, words):
        """colptr: int64 tensor (ncols_g + 1); words: int64 tensor of validity words or None"""
        n = colptr.numel() - 1
        if n > self.ncols:
            raise RuntimeError("validity exchange built for <= %d columns per rank, table has %d" % (self.ncols, n))
        if words is not None and words.numel() > self.cap:
            raise RuntimeError("validity exchange capacity exceeded")
        self.send[0] = n
        counts = self.send[1:1 + self.cnt_words].view(torch.int32)[:n]
        counts.copy_(colptr[1:] - colptr[:-1])             # int64 differences -> int32 (column lengths fit easily)
        if words is not None:
            self.send[1 + self.cnt_words:1 + self.cnt_words + words.numel()].copy_(words)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)

    def run(self, table):
        """Stream-ordered on torch's current stream: either make that the library's launching stream
        (mpb200_set_stream) or use the waiting forms of the validity calls before calling this."""
        colptr, _, _, words = table_device_tensors(table)
        self.run_arrays(colptr, words)

    def assemble(self):
        """host-side: global colptr + BitVector chunks from the gathered shards"""
        buf = self.recv.cpu().numpy().reshape(self.world, self.stride)
        ncols_of = [int(buf[g, 0]) for g in range(self.world)]
        counts = [buf[g, 1:1 + self.cnt_words].view(np.int32) for g in range(self.world)]
        words = [buf[g, 1 + self.cnt_words:].view(np.uint64) for g in range(self.world)]
        return _assemble(self.world, counts, words, ncols_of)


class PeerExchange:
    """The same exchange as direct peer stores over NVLink (mpb200_xchg_*, csrc/xchg.cu): one pack-and-store kernel
    + a device-side flag barrier per step, no NCCL call and no staging copies in the data path.  torch.distributed
    is used once, at construction, to agree on capacities and to swap the CUDA IPC handles."""

    def __init__(self, ncols, word_capacity, group=None):
        import ctypes
        from . import _lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        caps = torch.tensor([int(ncols), int(word_capacity)], dtype=torch.int64, device=_dev())
        dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
        self.ncols, self.cap = int(caps[0]), int(caps[1])
        lib = _lib.lib()
        self.h = _lib.c_vp()
        handle = (ctypes.c_char * 64)()
        _lib.check(lib.mpb200_xchg_create(self.rank, self.world, self.ncols, self.cap, ctypes.byref(self.h), handle))
        mine = torch.frombuffer(bytearray(bytes(handle)), dtype=torch.uint8).to(_dev())
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        blob = bytes(allh.cpu().numpy().tobytes())
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(lib.mpb200_xchg_connect(self.h, buf))
        dist.barrier(group=group)   # every rank has mapped every buffer before the first push

    def attach(self, table):
        """later builds of `table` send their column lengths early, underneath the fill (mpb200_xchg_attach);
        call again after the table handle changes (the first build creates it)"""
        from . import _lib
        if table.h:
            _lib.check(_lib.lib().mpb200_xchg_attach(self.h, table.h))
            self._attached = table

    def run(self, table):
        from . import _lib
        _lib.check(_lib.lib().mpb200_xchg_push(self.h, table.h))

    def assemble(self):
        import ctypes
        from . import _lib
        recv, slot, coff, woff, status = _lib.c_vp(), _lib.c_i64(), _lib.c_i64(), _lib.c_i64(), _lib.c_i64()
        _lib.check(_lib.lib().mpb200_xchg_view(self.h, ctypes.byref(recv), ctypes.byref(slot), ctypes.byref(coff),
                                               ctypes.byref(woff), ctypes.byref(status)))
        if status.value:
            raise RuntimeError("peer exchange: rank %d never arrived at the barrier" % (status.value - 1))
        raw = device_tensor(recv.value, self.world * slot.value, "|u1").cpu().numpy().reshape(self.world, slot.value)
        ncols_of = [int(raw[g, :8].view(np.int64)[0]) for g in range(self.world)]
        counts = [raw[g, coff.value:coff.value + 4 * self.ncols].view(np.int32) for g in range(self.world)]
        words = [raw[g, woff.value:woff.value + 8 * self.cap].view(np.uint64) for g in range(self.world)]
        return _assemble(self.world, counts, words, ncols_of)

    def close(self):
        from . import _lib
        if self.h:
            t = getattr(self, "_attached", None)
            if t is not None and t.h:
                _lib.load().mpb200_xchg_attach(None, t.h)   # the table must not keep a pointer to a dead exchange
            torch.cuda.synchronize()
            dist.barrier(group=self.group)   # no peer may still be storing into a buffer that is about to be freed
            _lib.load().mpb200_xchg_destroy(self.h)
            self.h = _lib.c_vp()


def make_exchange(ncols, word_capacity, group=None, kind=None):
    """PeerExchange when the GPUs can map each other's memory, else the NCCL all-gather form.  The choice is
    collective (all ranks agree).  kind = "peer" / "nccl" forces one."""
    import os
    kind = kind or os.environ.get("MPB200_EXCHANGE", "auto")
    if kind != "nccl" and dist.get_backend(group) == "nccl":
        ok = torch.ones(1, dtype=torch.int64, device=_dev())
        ex = None
        try:
            ex = PeerExchange(ncols, word_capacity, group)
        except Exception:
            if kind == "peer":
                raise
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok):
            return ex
    return ValidityExchange(ncols, word_capacity, group)
