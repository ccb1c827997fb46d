"""`slamdunk filter` with the decisions taken on the GPU: drop-in for the reference's
filter.Filter (slamdunk/dunks/filter.py:213-315), same signature, same log lines, same @RG DS /
@PG header rewrite, coordinate-sorted output.

Device: per-read predicates (MAPQ / identity / NM) or, when a BED file is given, the
multimapper-retention logic (sd_filter).  Host (C++, all threads): BAM decode (sd_decode_bam) and the
sorted, BGZF-compressed output with the RD tags and its BAI index (sd_rewrite_bam_indexed; `pysam.index`,
filter.py:313).
"""
from __future__ import print_function

import ctypes
import os

import numpy as np

from .. import _lib as L
from ..engine import CountEngine, DeviceBatch, HostBatch, make_params
from ..io import bam as bamio
from ..utils.bed import BedIterator
from ..utils.samheader import SlamSeqInfo, checkStep, md5
from ..version import __bam_version__, __version__

OPTIONS = {"device": 0, "threads": 0, "stats": None}


def Filter(inputBAM, outputBAM, log, bed, MQ=2, minIdentity=0.8, NM=-1, printOnly=False, verbose=True, force=False):
    if not (printOnly or checkStep([inputBAM], [outputBAM], force)):
        print("Skipped filtering for " + inputBAM, file=log)
        return
    hb = HostBatch(inputBAM, [], want_mp=False, threads=OPTIONS["threads"])
    # read.get_tag("XI") / get_tag("NM") raise KeyError on a record without the tag (filter.py:246,249; :130,133 in the
    # retention branch, which has no MQ test before them): a BAM that NextGenMap did not write must not pass silently
    if hb.n_mapped_no_xi:
        rec = hb.rec_array()[:hb.n_reads]
        reached = ((rec[:, 1] >> 24) & L.SD_F_UNMAPPED) == 0
        if bed is None:
            reached &= ((rec[:, 1] >> 16) & 0xFF) >= MQ
        if np.any(reached & np.isnan(rec[:, 2].view(np.float32))):
            hb.close()
            raise KeyError("tag 'XI' not present")
    if NM > -1 and hb.n_mapped_no_nm:
        hb.close()
        raise KeyError("tag 'NM' not present")
    eng = CountEngine(OPTIONS["device"])
    try:
        db = DeviceBatch.from_host(hb.batch, eng.device, eng)
        params = make_params(filter_on=True, filter_mq=MQ, filter_min_identity=minIdentity, filter_max_nm=NM)
        if bed is None:
            print("#No bed-file supplied. Running default filtering on " + inputBAM + ".", file=log)
            keep, stats, state = eng.filter(db, params)
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
            keep, stats, state = eng.filter(db, params, nh, (ci, st, sp, ni))
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

    text = hb.header_text
    n_records = hb.n_reads
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
    # ---- write the output BAM: surviving records byte for byte, RD:Z + cleared secondary/supplementary on retained
    # multimappers (filter.py:56-65), coordinate sorted (filter.py:304-310), BGZF-compressed on the host threads
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    state = np.ascontiguousarray(state, dtype=np.uint8)
    assert len(keep) == n_records
    lib = L.load()
    err = ctypes.create_string_buffer(512)
    written = ctypes.c_uint64(0)
    rc = lib.sd_rewrite_bam_indexed(os.fsencode(inputBAM), os.fsencode(outputBAM), os.fsencode(outputBAM + ".bai"),
                                    bamio.header_dict_to_text(header).encode(), keep.ctypes.data, state.ctypes.data, n_records,
                                    1, 6, OPTIONS["threads"], ctypes.byref(written), err, len(err))
    if rc != L.SD_OK:
        raise IOError("sd_rewrite_bam_indexed: " + err.value.decode("utf-8", "replace"))
    OPTIONS["stats"]["written"] = int(written.value)
