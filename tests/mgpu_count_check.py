"""Run under torchrun on N GPUs: computeTconversions with the reads sliced over the ranks must write the
same files as the committed goldens (rank 0 writes).  Used by the multi-GPU gpurun check:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_count_check.py
"""
import io
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import goldens                                  # noqa: E402
from slamdunk_b200 import parallel               # noqa: E402
from slamdunk_b200.dunks import tcounter         # noqa: E402


def atlas_check(rank, world, local_rank, reads_per_rank=2_000_000):
    """cfg5 in miniature: every rank generates and counts the reads of its own genome region, one NCCL all-reduce
    of the per-UTR counters; rank 0 then counts all regions' reads alone and must get the identical table."""
    import torch
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import CountEngine, make_params
    from slamdunk_b200.synth import SynthWorkload
    eng = CountEngine(local_rank)
    wl = SynthWorkload(eng, seed=1, genome_bp=30_000_000, n_utr=6000)
    wl.install()
    params = make_params(27, 1, L.SD_MISMATCH_CIGAR, filter_on=True, filter_mq=2, filter_min_identity=0.95)
    parts = wl.region_split(world)

    def shard(r):
        first, count = parts[r]
        return wl.generate(wl.params(seed=3, first_read=r * reads_per_rank, n_reads=reads_per_rank, utr_first=first, utr_count=count),
                           coordinate_sorted=True)[0]
    eng.count(shard(rank), params)
    parallel.reduce_engine_outputs(eng, positions=True)
    eng.synchronize()
    reduced = (eng.counters.clone(), eng.pos_cov.clone(), eng.pos_tc.clone(), eng.stats.clone())
    ok = True
    if rank == 0:
        eng.reset_outputs()
        for r in range(world):
            eng.count(shard(r), params)
        eng.synchronize()
        ok = all(torch.equal(a, b) for a, b in zip(reduced, (eng.counters, eng.pos_cov, eng.pos_tc, eng.stats)))
        print("mgpu", world, "atlas regions: reduce(shards) == unsharded:", "OK" if ok else "MISMATCH",
              "reads", int(reduced[3][L.SD_S_COUNTED]))
    eng.close()
    parallel.dist.barrier()
    return ok


def main():
    rank, world, local_rank = parallel.init("nccl")
    ok = True
    for name in goldens.COUNT_CASES:
        c = goldens.CountCase(name)
        d = tempfile.mkdtemp()
        out = os.path.join(d, "o")
        tcounter.computeTconversions(c.ref, c.bed, c.vcf, c.bam, c.max_read_length, c.min_qual, out + ".tsv", out + ".p", out + ".m",
                                     c.conv_threshold, io.StringIO())
        if rank == 0:
            same = (open(out + ".tsv").read() == open(c.tsv).read() and open(out + ".p").read() == open(c.plus).read()
                    and open(out + ".m").read() == open(c.minus).read())
            print("mgpu", world, name, "OK" if same else "MISMATCH")
            ok = ok and same
        parallel.dist.barrier()
    ok = atlas_check(rank, world, local_rank) and ok
    parallel.dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
