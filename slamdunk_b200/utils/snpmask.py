"""VCF -> SNP bit masks over the linear genome, with the reference's key semantics.

The reference (slamdunk/utils/SNPtools.py:23-60) stores T>C and A>G SNPs in dicts keyed by the
STRING `CHROM + POS` and queries with `chromosome + str(position + 1)`.  A record therefore masks
every (chromosome c, 1-based position p) with c + str(p) == CHROM + POS -- normally just its own
position, but e.g. 'chr12'+'345' also masks chr1:2345.  The masks built here reproduce exactly
that set, so the device only ever tests one bit.
"""
import os

import numpy as np


def read_vcf_records(path, warn=None):
    """Yield (CHROM, POS text, REF, ALT) like pybedtools iteration (SNPtools.py:43-49): header and
    blank lines skipped, columns split on TAB."""
    if path is None:
        return
    if not os.path.exists(path):
        if warn is not None:
            warn("Warning: SNP file " + path + " not found.")
        return
    with open(path, "rb") as probe:                  # pybedtools reads gzip / bgzip VCFs transparently (SNPtools.py:43)
        gz = probe.read(2) == b"\x1f\x8b"
    if gz:
        import gzip
        fh = gzip.open(path, "rt")
    else:
        fh = open(path, "r")
    with fh:
        for line in fh:
            if not line or line[0] == "#" or not line.strip():
                continue
            f = line.rstrip("\n").rstrip("\r").split("\t")
            yield f[0], f[1], f[3], f[4]


def _canonical_pos(text):
    """int(text) if text is what str(int) would print for a positive int, else None."""
    if text and text.isascii() and text.isdigit() and text[0] != "0":
        return int(text)
    return None


class SNPMasks(object):
    def __init__(self, names, chrom_off, chrom_len, n_words):
        self.names = list(names)
        self.chrom_off = [int(x) for x in chrom_off]
        self.chrom_len = [int(x) for x in chrom_len]
        self.n_words = int(n_words)
        self._tc = []
        self._ag = []
        self._cand = {}
        self.n_tc = 0
        self.n_ag = 0

    def _candidates(self, chrom):
        """FASTA chromosomes c for which chrom + POS can equal c + digits:
        (index, strip, need): key must start with `need` digits removed/added accordingly."""
        c = self._cand.get(chrom)
        if c is None:
            c = []
            for i, name in enumerate(self.names):
                if chrom.startswith(name) and (chrom[len(name):].isdigit() or name == chrom):
                    c.append((i, chrom[len(name):], ""))       # rest = extra + POS
                elif name.startswith(chrom) and name[len(chrom):].isdigit():
                    c.append((i, "", name[len(chrom):]))        # POS must start with `need`
            self._cand[chrom] = c
        return c

    def add(self, chrom, pos_text, ref, alt):
        """SNPDictionary._addSNP, SNPtools.py:30-38"""
        r, a = ref.upper(), alt.upper()
        if r == "T" and a == "C":
            dest = self._tc
        elif r == "A" and a == "G":
            dest = self._ag
        else:
            return
        for idx, extra, need in self._candidates(chrom):
            if need:
                if not pos_text.startswith(need):
                    continue
                p = _canonical_pos(pos_text[len(need):])
            else:
                p = _canonical_pos(extra + pos_text)
            if p is not None and p <= self.chrom_len[idx]:
                dest.append(self.chrom_off[idx] + p - 1)

    def _mask(self, positions):
        m = np.zeros(self.n_words, dtype=np.uint32)
        if positions:
            g = np.unique(np.asarray(positions, dtype=np.int64))
            np.bitwise_or.at(m, g >> 5, (np.uint32(1) << (g & 31).astype(np.uint32)))
        return m

    def masks(self):
        self.n_tc, self.n_ag = len(set(self._tc)), len(set(self._ag))
        return self._mask(self._tc), self._mask(self._ag)


def build_snp_masks(vcf_path, names, chrom_off, chrom_len, n_words, warn=None):
    """-> (tc_mask, ag_mask) uint32[n_words], or (None, None) when there is nothing to mask."""
    if vcf_path is None:
        return None, None
    sm = SNPMasks(names, chrom_off, chrom_len, n_words)
    n = 0
    for chrom, pos, ref, alt in read_vcf_records(vcf_path, warn):
        sm.add(chrom, pos, ref, alt)
        n += 1
    if not sm._tc and not sm._ag:
        return None, None
    return sm.masks()
