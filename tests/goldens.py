"""Index of the committed golden vectors (tests/golden, made by tests/golden/make_golden.py)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class CountCase:
    def __init__(self, name):
        d = os.path.join(GOLDEN, name)
        self.name = name
        self.dir = d
        if name == "cfg1":
            self.ref = os.path.join(d, "cfg1.fa")
            self.bed = os.path.join(d, "actb.bed")
            self.vcf = None
            self.bam = os.path.join(d, "reads_slamdunk_mapped_filtered.bam")
            self.max_read_length, self.min_qual, self.conv_threshold = 100, 27, 1
        else:
            p = json.load(open(os.path.join(d, "params.json")))["count"]
            self.ref = os.path.join(d, name + ".fa")
            self.bed = os.path.join(d, name + ".bed")
            self.vcf = os.path.join(d, name + ".vcf") if p["vcf"] else None
            self.bam = os.path.join(d, name + ".bam")
            self.max_read_length, self.min_qual, self.conv_threshold = p["max_read_length"], p["min_qual"], p["conv_threshold"]
        self.tsv = os.path.join(d, "expected_tcount.tsv")
        self.plus = os.path.join(d, "expected_tcount_plus.bedgraph")
        self.minus = os.path.join(d, "expected_tcount_mins.bedgraph")


class FilterCase:
    def __init__(self, name):
        d = os.path.join(GOLDEN, name)
        self.name = name
        self.dir = d
        self.expected = json.load(open(os.path.join(d, "expected.json")))
        self.bam = os.path.join(d, name + ".bam")
        self.bed = os.path.join(d, name + ".bed") if self.expected["filter"]["bed"] else None
        self.mq = self.expected["filter"]["mq"]
        self.min_identity = self.expected["filter"]["min_identity"]
        self.nm = self.expected["filter"]["nm"]


COUNT_CASES = ["cfg1", "case_a", "case_b", "case_c", "case_d", "case_e"]
FILTER_CASES = ["filt_default", "filt_nm", "filt_multimap", "filt_multimap_nm"]


class PositionalCase:
    """alleyoop positional-tracks on the inputs of a count case (tests/golden/<case>/positional)."""
    def __init__(self, name):
        c = CountCase(name)
        self.name, self.ref, self.vcf, self.bam = name, c.ref, c.vcf, c.bam
        self.dir = os.path.join(c.dir, "positional")
        p = json.load(open(os.path.join(self.dir, "params.json")))
        self.min_qual, self.conv_threshold, self.coverage_cutoff, self.stdout = p["min_qual"], p["conv_threshold"], p["coverage_cutoff"], p["stdout"]
        self.prefix_name = name + "_slamdunk_mapped_filtered_positional_rates"

    def expected(self, suffix):
        return os.path.join(self.dir, self.prefix_name + suffix)


POSITIONAL_CASES = ["case_a", "case_b", "case_d", "case_e"]


class ReadSepCase:
    """alleyoop read-separator on a count case's BAM with extra same-name secondary alignments (tests/golden/<case>/readsep)."""
    def __init__(self, name):
        c = CountCase(name)
        self.name, self.ref, self.vcf = name, c.ref, c.vcf
        self.dir = os.path.join(c.dir, "readsep")
        self.bam = os.path.join(self.dir, name + "_multi.bam")
        self.expected = json.load(open(os.path.join(self.dir, "expected.json")))
        self.min_qual, self.conv_threshold = self.expected["min_qual"], self.expected["conv_threshold"]


READSEP_CASES = ["case_a", "case_b", "case_e"]


class StatsCase:
    """alleyoop rates / utrrates / tcperreadpos / tcperutrpos / snpeval and the mle per-read lists on a count case's inputs
    (tests/golden/<case>/stats, written by the unmodified reference)."""
    RUNS = [("rates", False), ("tccontext", False), ("utrrates", False), ("utrrates", True), ("tcperreadpos", False), ("tcperutrpos", False),
            ("snpeval", False), ("snpeval", True)]

    def __init__(self, name):
        c = CountCase(name)
        self.name, self.ref, self.vcf, self.bam, self.bed = name, c.ref, c.vcf, c.bam, c.bed
        self.max_read_length, self.min_qual, self.conv_threshold = c.max_read_length, c.min_qual, c.conv_threshold
        self.dir = os.path.join(c.dir, "stats")

    def expected(self, tool, strict=False):
        return os.path.join(self.dir, tool + ("_strict" if strict else "") + ".csv")

    @property
    def perread(self):
        return os.path.join(self.dir, "tcount_perread.tsv")


STATS_CASES = ["case_a", "case_b", "case_d", "case_e"]
