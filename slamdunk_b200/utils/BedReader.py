"""BED6 reader with the reference's behaviour (slamdunk/utils/BedReader.py:33-84): lines are
rstrip()ped and split on TAB, columns 0-3 are mandatory (int() on start/stop), score and strand
optional, no comment / track / blank-line handling, rows kept in file order, duplicates kept."""


class BedEntry(object):
    __slots__ = ("chromosome", "start", "stop", "name", "score", "strand")

    def __init__(self):
        self.chromosome = ""
        self.start = 0
        self.stop = 0
        self.name = ""
        self.score = "."
        self.strand = "."

    def __repr__(self):
        return self.chromosome + "\t" + str(self.start) + "\t" + str(self.stop) + "\t" + self.name

    def getLength(self):
        return self.stop - self.start

    def hasStrand(self):
        return self.strand == "+" or self.strand == "-"

    def hasNonEmptyName(self):
        return self.name != ""


def _toBED(line):
    cols = line.rstrip().split("\t")
    e = BedEntry()
    e.chromosome = cols[0]
    e.start = int(cols[1])
    e.stop = int(cols[2])
    e.name = cols[3]
    if len(cols) > 4:
        e.score = cols[4]
    if len(cols) > 5:
        e.strand = cols[5]
    return e


class BedIterator(object):
    def __init__(self, filename):
        self._bedFile = open(filename, "r")

    def __iter__(self):
        return self

    def __next__(self):
        try:
            return _toBED(self._bedFile.__next__())
        except StopIteration:
            self._bedFile.close()
            raise StopIteration
