"""Host-only table tools downstream of `count`: `alleyoop summary` and `alleyoop merge`.

Nothing here touches the GPU: they read BAM headers and tcount files.  Kept so that pipelines (MultiQC, the
slamdunk nf-core workflow) that call them right after `count` keep working.

  readSummary  slamdunk/dunks/stats.py:531-589 (the PCA plot behind -t is an R script and is not run)
  mergeRates   slamdunk/dunks/stats.py:854-858 -> slamdunk/plot/merge_rate_files.R, restated in Python
"""
from __future__ import print_function

import ast
import math
import os
import struct
import zlib

from ..utils.samheader import SlamSeqInfo, getSampleInfo, parse_sam_header, replaceExtension
from ..version import __version__


def bam_header_text(path):
    """SAM header text of a BAM file, inflating only the BGZF blocks that hold it."""
    buf = b""
    with open(path, "rb") as fh:
        while True:
            head = fh.read(18)
            if len(head) < 18:
                break
            if head[:4] != b"\x1f\x8b\x08\x04":
                raise ValueError("not a BGZF file: " + path)
            xlen = struct.unpack_from("<H", head, 10)[0]
            extra = head[12:] + fh.read(xlen - 6)
            bsize, x = None, 0
            while x + 4 <= len(extra):
                si1, si2, slen = struct.unpack_from("<BBH", extra, x)
                if si1 == 66 and si2 == 67:
                    bsize = struct.unpack_from("<H", extra, x + 4)[0]
                x += 4 + slen
            if bsize is None:
                raise ValueError("not a BGZF file: " + path)
            rest = fh.read(bsize + 1 - 12 - xlen)
            buf += zlib.decompress(rest[:-8], -15)
            if len(buf) >= 8:
                if buf[:4] != b"BAM\1":
                    raise ValueError("not a BAM file: " + path)
                l_text = struct.unpack_from("<i", buf, 4)[0]
                if len(buf) >= 8 + l_text:
                    return buf[8:8 + l_text].split(b"\0", 1)[0].decode("utf-8", "replace")
    raise ValueError("truncated BAM header: " + path)


def sumCounts(countFile, column="ReadCount"):
    """stats.py:42-71"""
    columnId = -1
    total = 0
    with open(countFile) as f:
        for line in f:
            if line.startswith("#"):
                continue
            columns = line.rstrip().split("\t")
            if line.startswith("Chromosome"):
                for i, col in enumerate(columns):
                    if col == column:
                        columnId = i
                if columnId < 0:
                    return 0
            else:
                total += int(columns[columnId])
    return total


def readSummary(filteredFiles, countDirectory, outputFile, log, printOnly=False, verbose=True, force=False):
    contentDict = {}
    with open(outputFile, "w") as tsvFile:
        print("# slamdunk summary v" + __version__, file=tsvFile)
        for bam in filteredFiles:
            header = parse_sam_header(bam_header_text(bam))
            slamseqInfo = SlamSeqInfo(header)
            sampleInfo = getSampleInfo(header)
            fields = [bam, sampleInfo.Name, sampleInfo.Type, sampleInfo.Time, str(slamseqInfo.SequencedReads),
                      str(slamseqInfo.MappedReads), str(slamseqInfo.DedupReads), str(slamseqInfo.MQFilteredReads),
                      str(slamseqInfo.IdFilteredReads), str(slamseqInfo.NmFilteredReads), str(slamseqInfo.MultimapperReads),
                      str(slamseqInfo.FilteredReads)]
            if countDirectory is not None:
                countedReads = 0
                countFile = os.path.join(countDirectory, replaceExtension(os.path.basename(bam), ".tsv", "_tcount"))
                if not os.path.exists(countFile):
                    print("TCount directory does not seem to contain tcount file for:\t" + countFile)
                else:
                    countedReads = sumCounts(countFile)
                fields.append(str(countedReads))
            fields.append(slamseqInfo.AnnotationName)
            ID = len(contentDict) + 1 if int(sampleInfo.ID) in contentDict else sampleInfo.ID   # stats.py:559-562
            contentDict[int(ID)] = "\t".join(fields)
        cols = ["FileName", "SampleName", "SampleType", "SampleTime", "Sequenced", "Mapped", "Deduplicated", "MQ-Filtered",
                "Identity-Filtered", "NM-Filtered", "Multimap-Filtered", "Retained"]
        if countDirectory is not None:
            cols.append("Counted")
            print("R plotters are not part of slamdunk_b200: the PCA plot of the tcount files was not made.", file=log)
        print(*(cols + ["Annotation"]), sep="\t", file=tsvFile)
        for key in sorted(contentDict):
            print(contentDict[key], file=tsvFile)


# ---------------------------------------------------------------------------------------------------------------
_ALLOWED = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow, ast.USub, ast.UAdd,
            ast.Name, ast.Load, ast.Constant)


def _compile_column(expr):
    """The -c expression of `merge` (R: eval(parse(text = ...)) over the tcount columns): arithmetic on column names."""
    tree = ast.parse(expr.replace("^", "**"), mode="eval")
    for node in ast.walk(tree):
        if not isinstance(node, _ALLOWED):
            raise ValueError("merge: unsupported element in column expression: " + expr)
    return compile(tree, "<column>", "eval")


def _r_num(x):
    """write.table prints doubles with 15 significant digits; NaN / Inf as R spells them."""
    if isinstance(x, str):
        return x
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if float(x).is_integer() and abs(x) < 1e15:
        return str(int(x))
    return "%.15g" % x


def mergeRates(bams, outputCSV, column, columnName, log, printOnly=False, verbose=True, force=False):
    """merge_rate_files.R: one row per UTR (first six tcount columns), the averages of ReadsCPM / multimapCount / Tcontent /
    CoverageOnTs over the samples, then one column per sample = `column` evaluated on that sample's row (0 where
    ReadCount is 0), sample columns ordered by sample ID and named by meta field `columnName`
    (1 ID, 2 name, 3 type, 4 time)."""
    files = sorted(bams.split(","))
    code = _compile_column(column)
    merged, names, ids = None, [], []
    version = annotation = ""
    for i, path in enumerate(files):
        with open(path) as fh:
            l1 = fh.readline().rstrip("\n").split()
            l2 = fh.readline().rstrip("\n").split()
            head = fh.readline().rstrip("\n").split("\t")
            rows = [ln.rstrip("\n").split("\t") for ln in fh if ln.strip()]
        meta = [l1[6], l1[5], l1[7], _r_num(float(l1[8])), l2[1], l2[2], "\t".join(l1[0:3])]
        if i == 0:
            version, annotation = meta[6], "#Annotation:\t" + meta[4] + "\t" + meta[5]
            merged = [{"key": r[:6], "cpm": 0.0, "mm": 0.0, "tcontent": 0.0, "cov": 0.0, "vals": []} for r in rows]
        names.append(meta[int(columnName) - 1])
        ids.append(float(meta[0]))
        ix = {c: k for k, c in enumerate(head)}
        for m, r in zip(merged, rows):
            env = {c: float(r[k]) for c, k in ix.items() if k >= 6 or c in ("Start", "End", "Length")}
            m["cpm"] += env["ReadsCPM"]; m["mm"] += env["multimapCount"]; m["tcontent"] += env["Tcontent"]; m["cov"] += env["CoverageOnTs"]
            if env["ReadCount"] == 0:
                v = 0.0
            else:
                try:
                    v = float(eval(code, {"__builtins__": {}}, env))
                except ZeroDivisionError:
                    v = float("nan")
            m["vals"].append(v)
    n = float(len(files))
    order = sorted(range(len(files)), key=lambda k: ids[k])
    with open(outputCSV, "w") as out:
        print(version, file=out)
        print(annotation, file=out)
        print("#Expression:\t" + column, file=out)
        print(*(["Chromosome", "Start", "End", "Name", "Length", "Strand", "avgReadsCPM", "avgMultimapper", "avgTcontent",
                 "avgCoverageOnTs"] + [names[k] for k in order]), sep="\t", file=out)
        for m in merged or []:
            print(*(m["key"] + [_r_num(m["cpm"] / n), _r_num(m["mm"] / n), _r_num(m["tcontent"] / n), _r_num(m["cov"] / n)] +
                    [_r_num(m["vals"][k]) for k in order]), sep="\t", file=out)
