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
    parallel.dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
