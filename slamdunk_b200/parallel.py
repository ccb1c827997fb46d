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
