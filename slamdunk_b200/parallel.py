"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

The count path shards naturally: every output is a sum of independent integer
contributions of (read, UTR) pairs, so any partition of the reads gives bit-identical totals.
Two partitions are used:
  * UTR-owner regions (bench.py, synthetic atlas): each rank owns a contiguous genome region and
    the reads placed there; per-UTR counter rows are disjoint, one all-reduce of the int64
    [n_utr x 5] table assembles them (sum of disjoint rows);
  * read-batch slices (file inputs): each rank counts a contiguous range of 32-read tiles of the
    same BAM; counters AND the per-position arrays are all-reduced before sd_finalize turns the
    coverage differences into coverage (the scan is linear, so reduce-then-scan is exact).
There is no other collective on the path.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib as L


def env():
    """-> (rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Join the process group described by the environment; no-op for a single process."""
    rank, world, local_rank = env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of n_items for this rank (sizes differ by at most one)."""
    base, extra = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def sub_batch(batch, t0, t1):
    """View of tiles [t0, t1) of a HOST sd_batch (no copy): record and tile-table pointers advance,
    the column base pointers stay because tile offsets are absolute."""
    b = L.Batch()
    ctypes.memmove(ctypes.byref(b), ctypes.byref(batch), ctypes.sizeof(L.Batch))
    t0, t1 = int(t0), int(t1)
    b.n_tiles = t1 - t0
    first = t0 * batch.tile_reads
    b.n_reads = max(0, min(int(batch.n_reads), t1 * batch.tile_reads) - first)
    b.rec = batch.rec + first * ctypes.sizeof(L.ReadRec)
    b.tiles = batch.tiles + t0 * ctypes.sizeof(L.Tile)
    return b


def sub_wire(wire, t0, t1):
    """View of tiles [t0, t1) of a HOST sd_wire (no copy): the tile table and the per-record columns advance, the per-base,
    var and mp columns stay because the tile offsets into them are absolute."""
    w = L.Wire()
    ctypes.memmove(ctypes.byref(w), ctypes.byref(wire), ctypes.sizeof(L.Wire))
    t0, t1 = int(t0), int(t1)
    w.n_tiles = t1 - t0
    first = t0 * wire.tile_reads
    w.n_reads = max(0, min(int(wire.n_reads), t1 * wire.tile_reads) - first)
    w.tiles = wire.tiles + t0 * ctypes.sizeof(L.WireTile)
    for name, size in (("pos16", 2), ("mqf", 2), ("xi16", 2), ("nm16", 2), ("nmp16", 2), ("xi32", 4)):
        p = getattr(wire, name)
        setattr(w, name, p + first * size if p else None)
    return w


_HOST_GROUP = None


def host_group():
    """Process group for host-side (CPU tensor) exchanges: the default group when it is gloo, else a gloo group over the same
    ranks next to the NCCL one (made once; every rank must call this at the same point)."""
    global _HOST_GROUP
    if dist.get_backend() == "gloo":
        return None
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo")
    return _HOST_GROUP


_WIRE_REC_COLS = (("pos16", 2), ("mqf", 2), ("xi16", 2), ("nm16", 2), ("nmp16", 2), ("xi32", 4))


def _host_bytes(addr, n):
    """uint8 tensor over n bytes of host memory at addr (no copy)"""
    if not addr or n <= 0:
        return torch.zeros(0, dtype=torch.uint8)
    return torch.frombuffer((ctypes.c_uint8 * n).from_address(addr), dtype=torch.uint8)


def scatter_wire(wire, rank, world, src=0):
    """Rank `src` holds the wire form of a whole BAM (sd_decode_bam with SD_DECODE_WIRE); every rank gets the contiguous tile slice
    shard_range() assigns it: `src` a view of its own (no copy), the others a copy sent over the host group -- one rank inflates
    and decodes the file instead of all of them.  The tile table of a received slice is rebased to its own columns.
    -> (L.Wire with host pointers, object that keeps the buffers alive).  `wire` is ignored on ranks other than `src`."""
    import numpy as np
    grp = host_group()
    meta = [None]
    if rank == src:
        nt = int(wire.n_tiles)
        tiles = np.frombuffer((ctypes.c_uint8 * ((nt + 1) * 32)).from_address(wire.tiles), dtype=L.WIRE_TILE_DT)
        qb = int(wire.qual_bits)
        parts = []
        for r in range(world):
            t0, t1 = shard_range(nt, r, world)
            b0, b1 = int(tiles[t0]["base_off"]), int(tiles[t1]["base_off"])
            v0, v1 = int(tiles[t0]["var_off"]), int(tiles[t1]["var_off"])
            m0, m1 = int(tiles[t0]["mp_off"]), int(tiles[t1]["mp_off"])
            parts.append(dict(t0=t0, t1=t1, b0=b0, b1=b1, v0=v0, v1=v1, m0=m0, m1=m1))
        meta[0] = dict(parts=parts, n_reads=int(wire.n_reads), qual_bits=qb, qual_dict=list(wire.qual_dict), max_lseq=int(wire.max_lseq),
                       max_span=int(wire.max_span), max_tile_cig=int(wire.max_tile_cig), max_tile_mp=int(wire.max_tile_mp),
                       n_xi_dict=int(wire.n_xi_dict), has={k: bool(getattr(wire, k)) for k, _ in _WIRE_REC_COLS}, has_mp=bool(wire.mp))
    dist.broadcast_object_list(meta, src=src, group=grp)
    M = meta[0]
    qb = M["qual_bits"]

    def columns(w_or_none, part):
        """[(name, byte size, source address or None)] of one rank's slice, in a fixed order"""
        nt = part["t1"] - part["t0"]
        first = part["t0"] * 32
        cols = [("tiles", (nt + 1) * 32, (w_or_none.tiles + part["t0"] * 32) if w_or_none is not None else None)]
        for name, size in _WIRE_REC_COLS:
            if M["has"][name]:
                cols.append((name, nt * 32 * size, (getattr(w_or_none, name) + first * size) if w_or_none is not None else None))
        nb = part["b1"] - part["b0"]
        cols.append(("seq", nb // 4, (w_or_none.seq + part["b0"] // 4) if w_or_none is not None else None))
        cols.append(("qual", nb * qb // 8, (w_or_none.qual + part["b0"] * qb // 8) if w_or_none is not None else None))
        cols.append(("var", (part["v1"] - part["v0"]) * 4, (w_or_none.var + part["v0"] * 4) if w_or_none is not None else None))
        if M["has_mp"]:
            cols.append(("mp", (part["m1"] - part["m0"]) * 4, (w_or_none.mp + part["m0"] * 4) if w_or_none is not None else None))
        return cols

    xi_dict = torch.zeros(max(1, M["n_xi_dict"]), dtype=torch.float32)
    if rank == src and M["n_xi_dict"]:
        xi_dict.copy_(torch.frombuffer((ctypes.c_uint8 * (4 * M["n_xi_dict"])).from_address(wire.xi_dict), dtype=torch.float32))
    dist.broadcast(xi_dict, src=src, group=grp)
    if rank == src:
        for r in range(world):
            if r == src:
                continue
            for _name, nbytes, addr in columns(wire, M["parts"][r]):
                if nbytes:
                    dist.send(_host_bytes(addr, nbytes), dst=r, group=grp)
        mine = sub_wire(wire, M["parts"][src]["t0"], M["parts"][src]["t1"])
        return mine, (wire, xi_dict)
    part = M["parts"][rank]
    pin = torch.cuda.is_available()
    bufs = {}
    for name, nbytes, _ in columns(None, part):
        t = torch.zeros(nbytes + 16, dtype=torch.uint8, pin_memory=pin)      # the expansion reads a few words past a column's end
        if nbytes:
            dist.recv(t[:nbytes], src=src, group=grp)
        bufs[name] = t
    nt = part["t1"] - part["t0"]
    tiles = bufs["tiles"][:(nt + 1) * 32].numpy().view(L.WIRE_TILE_DT)
    tiles["base_off"] -= part["b0"]
    tiles["var_off"] -= part["v0"]
    tiles["mp_off"] -= part["m0"]
    w = L.Wire()
    w.n_tiles, w.tile_reads = nt, 32
    w.n_reads = max(0, min(M["n_reads"], part["t1"] * 32) - part["t0"] * 32)
    w.tiles = bufs["tiles"].data_ptr()
    for name, _size in _WIRE_REC_COLS:
        setattr(w, name, bufs[name].data_ptr() if name in bufs else None)
    w.seq, w.qual, w.var = bufs["seq"].data_ptr(), bufs["qual"].data_ptr(), bufs["var"].data_ptr()
    w.mp = bufs["mp"].data_ptr() if "mp" in bufs else None
    w.xi_dict = xi_dict.data_ptr() if M["n_xi_dict"] else None
    w.n_xi_dict = M["n_xi_dict"]
    w.qual_bits = M["qual_bits"]
    for i in range(16):
        w.qual_dict[i] = M["qual_dict"][i]
    w.max_lseq, w.max_span, w.max_tile_cig, w.max_tile_mp = M["max_lseq"], M["max_span"], M["max_tile_cig"], M["max_tile_mp"]
    return w, (bufs, xi_dict)


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of integer tensors (NCCL on GPU tensors, gloo on CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def reduce_engine_outputs(eng, positions=True, group=None):
    """All-reduce an engine's outputs on its stream: counters + stats, and the per-position arrays
    when the reads (not the UTRs) were partitioned.  Call before eng.finalize().  Returns after the reduction has
    completed, so that a decision taken from the reduced values is the same on every rank."""
    with torch.cuda.stream(eng.stream):
        ts = [eng.counters, eng.stats]
        if positions and eng.n_positions:
            ts += [eng.pos_cov, eng.pos_tc]
        allreduce_sum_(ts, group)
    eng.synchronize()      # callers read eng.stats / eng.counters next, from other streams
