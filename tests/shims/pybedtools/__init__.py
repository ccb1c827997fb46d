"""I/O-only stand-in for pybedtools.BedTool over a VCF file (oracle shim).
The reference (utils/SNPtools.py:40-51) only iterates records and indexes columns 0,1,3,4."""


class _Fields(list):
    pass


class BedTool(object):
    def __init__(self, fn):
        self.fn = fn
        self._rows = []
        first = None
        with open(fn, "r") as fh:
            for line in fh:
                if first is None and line.strip():
                    first = line
                if line.startswith("#") or not line.strip():
                    continue
                self._rows.append(_Fields(line.rstrip("\n").rstrip("\r").split("\t")))
        if first is None:
            self.file_type = "empty"
        elif first.startswith("##fileformat=VCF") or first.startswith("#CHROM") or fn.endswith(".vcf"):
            self.file_type = "vcf"
        else:
            self.file_type = "bed"

    def __iter__(self):
        return iter(self._rows)

    def __len__(self):
        return len(self._rows)
