"""Host helpers of the count path with the reference's behaviour (slamdunk/utils/misc.py)."""
import ast
import collections
import hashlib
import os

SampleInfo = collections.namedtuple("SampleInfo", "ID Name Type Time")


def md5(fname):
    """utils/misc.py:84-89"""
    h = hashlib.md5()
    with open(fname, "rb") as f:
        for chunk in iter(lambda: f.read(4096), b""):
            h.update(chunk)
    return h.hexdigest()


def replaceExtension(inFile, newExtension, suffix=""):
    """utils/misc.py:111-112"""
    return os.path.splitext(inFile)[0] + suffix + newExtension


def parse_sam_header(text):
    """SAM header text -> {'RG': [dict], 'PG': [dict], 'SQ': [dict], ...} (pysam header.to_dict() layout)."""
    hdr = {}
    for line in text.split("\n"):
        if len(line) < 3 or line[0] != "@":
            continue
        kind = line[1:3]
        if kind == "CO":
            hdr.setdefault("CO", []).append(line[4:])
            continue
        fields = collections.OrderedDict()
        for f in line.split("\t")[1:]:
            if len(f) >= 3 and f[2] == ":":
                fields[f[:2]] = f[3:]
        if kind == "HD":
            hdr["HD"] = fields
        else:
            hdr.setdefault(kind, []).append(fields)
    return hdr


def getReadGroup(header):
    """utils/misc.py:229-236 on an already parsed header."""
    if "RG" in header and len(header["RG"]) > 0:
        return header["RG"][0]
    raise RuntimeError("Could not get mapped/unmapped/filtered read counts from BAM file. RG is missing. Please rerun slamdunk filter.")


def getSampleInfo(header):
    """utils/misc.py:238-241"""
    rg = getReadGroup(header)
    parts = rg["SM"].split(":")
    return SampleInfo(ID=rg["ID"], Name=parts[0], Type=parts[1], Time=parts[2])


class SlamSeqInfo(object):
    """Read-count bookkeeping kept in the BAM header's @RG DS field as a Python dict literal
    (utils/misc.py:33-82)."""
    _KEYS = (("SequencedReads", "sequenced"), ("MappedReads", "mapped"), ("FilteredReads", "filtered"),
             ("MQFilteredReads", "mqfiltered"), ("IdFilteredReads", "idfiltered"), ("NmFilteredReads", "nmfiltered"),
             ("MultimapperReads", "multimapper"), ("DedupReads", "dedup"), ("SNPs", "snps"),
             ("AnnotationName", "annotation"), ("AnnotationMD5", "annotationmd5"))

    def __init__(self, header=None):
        if header is None:
            for attr, _k in self._KEYS:
                setattr(self, attr, 0)
            self.AnnotationName = "NA"
            self.AnnotationMD5 = "NA"
        else:
            ds = ast.literal_eval(getReadGroup(header)["DS"])
            for attr, k in self._KEYS:
                setattr(self, attr, ds[k] if k in ds else "NA")

    def __repr__(self):
        return ("{'sequenced':%s,'mapped':%s,'filtered':%s,'mqfiltered':%s,'idfiltered':%s,'nmfiltered':%s,"
                "'multimapper':%s,'dedup':%s,'snps':%s,'annotation':'%s','annotationmd5':'%s'}"
                % (self.SequencedReads, self.MappedReads, self.FilteredReads, self.MQFilteredReads,
                   self.IdFilteredReads, self.NmFilteredReads, self.MultimapperReads, self.DedupReads, self.SNPs,
                   self.AnnotationName, self.AnnotationMD5))


def files_exist(files):
    """utils/misc.py:125-133"""
    if type(files) is list:
        return all(os.path.exists(f) for f in files)
    return os.path.exists(files)


def checkStep(inFiles, outFiles, force=False):
    """make-style skipping, utils/misc.py:147-164"""
    if not files_exist(inFiles):
        raise RuntimeError("One or more input files don't exist: " + str(inFiles))
    inFileDate = os.path.getmtime(inFiles[0])
    for x in inFiles[1:]:
        inFileDate = max(inFileDate, os.path.getmtime(x))
    if len(outFiles) > 0 and files_exist(outFiles):
        outFileDate = os.path.getmtime(outFiles[0])
        for x in outFiles[1:]:
            outFileDate = min(outFileDate, os.path.getmtime(x))
        if outFileDate > inFileDate:
            return True if force == True else False   # noqa: E712 (kept literal)
    return True
