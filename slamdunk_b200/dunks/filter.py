"""`slamdunk filter` with the decisions taken on the GPU: drop-in for the reference's
filter.Filter (slamdunk/dunks/filter.py:213-315), same signature, same log lines, same @RG DS /
@PG header rewrite, coordinate-sorted output.

Device: per-read predicates (MAPQ / identity / NM) or, when a BED file is given, the
multimapper-retention logic (sd_filter).  Host: BAM decode (C++), and writing the surviving records
with the plain Python BAM codec -- the native BAM writer / BAI index is the next row of the build
plan, so `pysam.index` has no equivalent here yet and `slamdunk_b200 count` reads
the whole file instead of using an index.
"""
from __future__ import print_function

import os

import numpy as np

from .. import _lib as L
from ..engine import CountEngine, DeviceBatch, HostBatch, make_params
from ..io import bam as bamio
from ..utils.bed import BedIterator
from ..utils.samheader import SlamSeqInfo, checkStep, md5
from ..version import __bam_version__, __version__

OPTIONS = {"device": 0, "threads": 0, "stats": None}


def _rd_tags(records, bam_names, keep, state, state_hits):
    """RD:Z text of every record written by dumpBufferToBam (filter.py:56-65): the reference's
    `multimapList` string, including its quirk of not being reset by a unique read that finds the
    buffer empty (filter.py:176-187)."""
    tags = {}
    mlist = []
    prev = ""
    open_retained = None
    buffer_nonempty = False
    for i, r in enumerate(records):
        st = state[i]
        if st == 0:
            continue
        if st == 2:
            if r.qname != prev and prev != "":
                if open_retained is not None:
                    tags[open_retained] = " ".join(mlist)
                open_retained = None
                buffer_nonempty = False
                mlist = []
            mlist.append("%s:%d-%d" % (bam_names[r.ref_id], r.pos, r.reference_end))
            prev = r.qname
        else:
            if buffer_nonempty:
                if open_retained is not None:
                    tags[open_retained] = " ".join(mlist)
                open_retained = None
                buffer_nonempty = False
                mlist = []
            prev = r.qname
        # group bookkeeping: the device tells us which record of the group is retained; whether the
        # buffer is non-empty equals "some alignment of the open group hit a UTR"
        if st == 2 and state_hits[i]:
            buffer_nonempty = True
        if keep[i] == 2:
            open_retained = i
    if open_retained is not None:
        tags[open_retained] = " ".join(mlist)
    return tags



def Filter(inputBAM, outputBAM, log, bed, MQ=2, minIdentity=0.8, NM=-1, printOnly=False, verbose=True, force=False):
    if not (printOnly or checkStep([inputBAM], [outputBAM], force)):
        print("Skipped filtering for " + inputBAM, file=log)
        return
    hb = HostBatch(inputBAM, [], want_mp=False, threads=OPTIONS["threads"])
    eng = CountEngine(OPTIONS["device"])
    try:
        db = DeviceBatch.from_host(hb.batch, eng.device)
        params = make_params(filter_on=True, filter_mq=MQ, filter_min_identity=minIdentity, filter_max_nm=NM)
        if bed is None:
            print("#No bed-file supplied. Running default filtering on " + inputBAM + ".", file=log)
            keep, stats = eng.filter(db, params)
            utr_rows = None
        else:
            print("#Bed-file supplied. Running multimap retention filtering strategy on " + inputBAM + ".", file=log)
            utr_rows = list(BedIterator(bed))
            chrom_of = {n: i for i, n in reversed(list(enumerate(hb.ref_names)))}
            name_ids = {}
            ci = np.array([chrom_of.get(u.chromosome, L.SD_REF_NONE) for u in utr_rows], dtype=np.uint32)
            st = np.array([u.start for u in utr_rows], dtype=np.int64)
            sp = np.array([u.stop for u in utr_rows], dtype=np.int64)
            ni = np.array([name_ids.setdefault(u.name, len(name_ids)) for u in utr_rows], dtype=np.uint32)
            import torch
            nh = torch.from_numpy(hb.name_hash().view(np.int64).copy()).to(eng.device)
            keep, stats = eng.filter(db, params, nh, (ci, st, sp, ni))
            if stats[L.SD_S_SLOWPATH] > 0:
                raise RuntimeError("multimapper group exceeds the device limits (more than 32 UTR names / hits per alignment)")
    finally:
        eng.close()
    mapped, unmapped = int(stats[L.SD_S_MAPPED]), int(stats[L.SD_S_UNMAPPED])
    filtered = int(stats[L.SD_S_FILTERED])
    mqf, idf, nmf = int(stats[L.SD_S_MQFILTERED]), int(stats[L.SD_S_IDFILTERED]), int(stats[L.SD_S_NMFILTERED])
    multimapper = 0
    print("Criterion\tFiltered reads", file=log)
    if bed is None:
        print("MQ < " + str(MQ) + "\t" + str(mqf), file=log)
    else:
        multimapper = mapped - filtered - idf - nmf                   # filter.py:201
        print("MQ < 0\t0", file=log)
    print("ID < " + str(minIdentity) + "\t" + str(idf), file=log)
    print("NM > " + str(NM) + "\t" + str(nmf), file=log)
    print("MM\t" + str(multimapper), file=log)
    OPTIONS["stats"] = dict(mapped=mapped, unmapped=unmapped, filtered=filtered, mqfiltered=mqf, idfiltered=idf,
                            nmfiltered=nmf, multimapper=multimapper)

    # ---- write the output BAM (host; plain codec)
    text, refs, records = bamio.read_bam(inputBAM)
    assert len(records) == len(keep)
    bam_names = [n for n, _l in refs]
    written = []
    rd = {}
    if bed is not None and (keep == 2).any():
        rec = hb.rec_array()[:len(records)]
        flags = rec[:, 1] >> 24
        mapq = (rec[:, 1] >> 16) & 0xFF
        xi = rec[:, 2].view(np.float32).astype(np.float64)
        nm = rec[:, 4] & 0xFFFF
        surv = ((flags & L.SD_F_UNMAPPED) == 0) & ~(xi < minIdentity) & ~((NM > -1) & (nm > NM))
        state = np.where(surv, np.where(mapq == 0, 2, 1), 0)
        hits = np.zeros(len(records), dtype=bool)
        by_chrom = {}
        for u in utr_rows:
            by_chrom.setdefault(u.chromosome, []).append((u.start, u.stop + 1))
        for i in np.nonzero(state == 2)[0]:
            r = records[int(i)]
            e = r.reference_end
            for a, b in by_chrom.get(bam_names[r.ref_id], ()):
                if a < e and b > r.pos:
                    hits[i] = True
                    break
        rd = _rd_tags(records, bam_names, keep, state, hits)
    for i, r in enumerate(records):
        if keep[i] == 1:
            written.append(r)
        elif keep[i] == 2:
            r.set_tag("RD", rd.get(i, ""), "Z")
            r.flag &= ~(bamio.FLAG_SECONDARY | bamio.FLAG_SUPPLEMENTARY)
            written.append(r)
    hb.close()
    header = bamio.parse_header_text(text)
    if "RG" in header and len(header["RG"]) > 0:                     # filter.py:276-295
        info = SlamSeqInfo()
        info.SequencedReads = mapped + unmapped
        info.MappedReads = mapped
        info.FilteredReads = filtered
        info.MQFilteredReads = mqf
        info.IdFilteredReads = idf
        info.NmFilteredReads = nmf
        info.MultimapperReads = multimapper
        if bed is not None:
            info.AnnotationName = os.path.basename(bed)
            info.AnnotationMD5 = md5(bed)
        else:
            info.AnnotationName = ""
            info.AnnotationMD5 = ""
        header["RG"][0]["DS"] = str(info)
    pg = {"ID": "slamdunk", "PN": "slamdunk filter v" + __version__, "VN": __bam_version__}   # filter.py:298
    header.setdefault("PG", []).append(pg)
    bamio.write_bam(outputBAM, bamio.header_dict_to_text(header), refs, bamio.sort_records(written))
