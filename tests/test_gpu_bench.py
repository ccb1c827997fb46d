"""bench.py end to end at miniature sizes: every workload of BASELINE.json produces its contract line with the parity
checks green (the GPU path on the benchmarked column form vs the oracle; sharded == unsharded on >= 2 GPUs), and the
pieces the workloads are made of -- region shards of one read stream, the sample's own SNPs in the generator."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_bridge as ob
from oracle import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--genome-bp", "6000000", "--utrs", "1200", "--steps", "2", "--warmup", "1", "--cpu-sample", "60000"]


def _free_port():
    """a rendezvous port of its own for every torchrun: a port still in TIME_WAIT from the previous run stalls the static rendezvous"""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _bench(extra, gpus=1, timeout=900):
    env = dict(os.environ)
    if gpus == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py")] + SMALL + extra
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", str(gpus)] + SMALL + extra
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr.decode()[-3000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout.decode()
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,extra,reads", [
    ("cfg2", ["--reads", "300000"], 300_000),
    ("cfg2", ["--reads", "200000", "--mismatch", "verify"], 200_000),
    ("cfg2", ["--reads", "200000", "--mismatch", "mp"], 200_000),
    ("cfg3", ["--reads", "40000"], 12 * 40_000),
    ("cfg4", ["--reads", "100000"], 100_000),
    ("cfg5", ["--reads", "500000"], 500_000)])
def test_bench_line_of_every_workload(workload, extra, reads):
    d = _bench(["--workload", workload] + extra)
    assert d["config"]["workload"].startswith(workload) and d["n_gpus"] == 1
    assert d["reads_per_step"] == reads
    assert d["parity"]["oracle"].startswith("bit-exact vs oracle"), d["parity"]
    assert d["value"] > 0 and d["roofline"]["frac"] > 0 and d["roofline"]["kernel_ms"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < d["e2e"]["counted_reads"] <= reads and d["counted_reads_last_step"] == d["e2e"]["counted_reads"]
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] > 0 and d["slow_path_reads"] == 0
    if workload == "cfg4":
        assert "no MAPQ threshold" in d["config"]["filter"]


def test_reference_arm_line():
    d = _bench(["--impl", "reference", "--reads", "100000"])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_sharded_bench_equals_unsharded_on_two_gpus():
    """cfg5 / cfg2 under torchrun on 2 GPUs: the run itself compares the region-sharded count (NCCL all-reduce) with the
    unsharded one on a subsample of the same read stream."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for workload, reads in (("cfg5", "600000"), ("cfg2", "200000"), ("cfg3", "30000")):
        d = _bench(["--workload", workload, "--reads", reads, "--no-cpu-baseline"], gpus=2)
        assert d["n_gpus"] == 2
        if workload != "cfg3":
            assert "== unsharded" in d["parity"]["sharding"], d["parity"]
        if workload == "cfg5":
            assert d["reads_per_step"] == 600_000


def _setup(genome_bp=3_000_000, n_utr=600, seed=5):
    from slamdunk_b200.engine import CountEngine
    from slamdunk_b200.synth import SynthWorkload
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=seed, genome_bp=genome_bp, n_utr=n_utr, n_frac=0.01)
    wl.install()
    return eng, wl


def test_region_shards_of_one_stream_add_up_and_own_disjoint_rows():
    """BASELINE.json config 5 on one GPU: the reads of one stream split by chromosome region (bench.Bench.region_bounds) are
    counted shard by shard; the sum equals the unsharded count, every read has one owner, and no two shards touch the same
    counter row or per-position slot (what lets a rank write its slice of the outputs without a reduction of the arrays)."""
    import argparse
    import bench
    from slamdunk_b200.engine import make_params
    eng, wl = _setup(genome_bp=4_000_000, n_utr=900)
    b = bench.Bench.__new__(bench.Bench)
    b.wl, b.args = wl, argparse.Namespace(read_len=100)
    n = 150_000
    params = make_params(27, 1, 0, filter_on=True, filter_mq=2, filter_min_identity=0.95)
    p = wl.params(seed=3, first_read=0, n_reads=n)
    whole, _ = wl.generate(p, coordinate_sorted=True)
    eng.count(whole, params)
    eng.finalize()
    want = eng.results()
    for n_parts in (2, 5):
        bounds = b.region_bounds(n_parts)
        assert bounds[0][0] == 0 and all(bounds[k][1] == bounds[k + 1][0] for k in range(n_parts - 1))
        total = {k: np.zeros_like(want[k]) for k in ("counters", "pos_cov", "pos_tc", "stats")}
        touched_rows = np.zeros(len(wl.ann), dtype=np.int32)
        touched_pos = np.zeros(eng.n_positions, dtype=np.int32)
        reads = 0
        for lo, hi in bounds:
            db, _ = wl.generate(wl.params(seed=3, first_read=0, n_reads=n), coordinate_sorted=True, gpos_range=(lo, hi))
            assert db is not None
            reads += int(db.batch.n_reads)
            eng.reset_outputs()
            eng.count(db, params)
            eng.synchronize()
            r = {"counters": eng.counters.cpu().numpy().reshape(-1, 5), "pos_cov": eng.pos_cov.cpu().numpy()[:eng.n_positions],
                 "pos_tc": eng.pos_tc.cpu().numpy()[:eng.n_positions], "stats": eng.stats.cpu().numpy()}
            touched_rows += (r["counters"][:, 0] > 0)
            touched_pos += (r["pos_cov"] != 0)
            for k in total:
                total[k] = total[k] + r[k]
        assert reads == n
        assert touched_rows.max() == 1 and touched_pos.max() == 1
        # the unsharded arrays went through the coverage scan; do the same with the summed difference arrays
        import torch
        eng.reset_outputs()
        with torch.cuda.stream(eng.stream):
            eng.counters.copy_(torch.from_numpy(total["counters"].reshape(-1)))
            eng.pos_cov.copy_(torch.from_numpy(total["pos_cov"]))
            eng.pos_tc.copy_(torch.from_numpy(total["pos_tc"]))
        eng.finalize()
        got = eng.results()
        for k in ("counters", "pos_cov", "pos_tc"):
            assert np.array_equal(want[k], got[k]), (n_parts, k)
        assert np.array_equal(want["stats"][:7], total["stats"][:7])
    assert wl.generate(wl.params(seed=3, first_read=0, n_reads=1000), gpos_range=(1 << 33, 1 << 34)) == (None, None)
    eng.close()


def test_sample_snps_reach_the_reads_and_are_masked():
    """BASELINE.json config 4 in miniature: the generator's reads carry the alternative allele at the sample's SNP sites, the
    VCF built from those sites masks them (utils/SNPtools.py:30-60 semantics incl. multi-allelic records that never match),
    and the masked count equals the oracle's."""
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.utils.snpmask import SNPMasks
    eng, wl = _setup(genome_bp=2_000_000, n_utr=500, seed=8)
    params = make_params(27, 1, 0, filter_on=True, filter_mq=0, filter_min_identity=0.95)
    n = 60_000
    p = dict(seed=30, first_read=0, n_reads=n, read_len=250, frac_multimap=0.15)
    db0, _ = wl.generate(wl.params(**p), want_triples=False, coordinate_sorted=True)
    eng.count(db0, params)
    eng.finalize()
    plain = eng.results()
    records = wl.set_sample_snps(spacing=500, seed=30, p_allele=0.9)
    assert abs(len(records) - 2_000_000 // 500) < 400
    assert any("," in r[3] for r in records) and any(r[2] == "T" and r[3] == "C" for r in records) and any(r[2] == "A" and r[3] == "G" for r in records)
    db, triples = wl.generate(wl.params(**p), want_triples=True, coordinate_sorted=True)
    eng.reset_outputs()
    eng.count(db, params)
    eng.finalize()
    unmasked = eng.results()
    sm = SNPMasks(wl.names, wl.chrom_off, wl.chrom_len, wl.n_words)
    for r in records:
        sm.add(*r)
    eng.set_snps(*sm.masks())
    eng.reset_outputs()
    eng.count(db, params)
    eng.finalize()
    masked = eng.results()
    conv = lambda r: int(r["counters"][:, L.SD_C_CONV_ON_T].sum())      # noqa: E731
    # the alleles look like conversions (about 0.056 per 250 bp read against 1.5 real ones); the mask takes them out again,
    # except at the multi-allelic sites
    assert conv(unmasked) - conv(masked) > 0.01 * conv(plain)
    assert 0.97 * conv(plain) < conv(masked) < 1.03 * conv(plain)
    assert int(masked["counters"][:, L.SD_C_MULTIMAP].sum()) > 0.1 * n
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, wl.lo.cpu().numpy().view(np.uint32),
                              wl.hi.cpu().numpy().view(np.uint32), wl.nmask.cpu().numpy().view(np.uint32))
    reads, qual, mp = ob.batch_to_oracle(ob.device_batch_to_numpy(db), triples.cpu().numpy())
    assert len(reads) == n
    rec = ob.device_batch_to_numpy(db)["rec"][:n]
    xi_ok = rec[:, 2].view(np.float32).astype(np.float64) >= 0.95
    h = oracle.make_snps(records)
    try:
        rows, bg = oracle.count_arrays(g, reads[xi_ok], qual, ob.annotation_to_oracle(wl.ann, len(wl.names)), 27, 1, h, mp_triples=mp, threads=4)
    finally:
        oracle.free_snps(h)
    assert ob.compare(masked, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
    eng.close()


def test_file_path_on_two_gpus_writes_the_golden_files():
    """computeTconversions under torchrun on 2 GPUs (rank 0 decodes, the wire slices are scattered, NCCL all-reduce of the outputs,
    rank 0 writes): every golden case byte for byte, and the region-sharded atlas check (tests/mgpu_count_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_count_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, cwd=ROOT)
    text = out.stdout.decode()
    assert out.returncode == 0 and "MGPU_CHECK PASS" in text, text[-3000:]
