"""BED6 annotation reader with the reference's parsing rules (slamdunk/utils/BedReader.py:63-77):
a line is rstrip()ped and split on TAB; chromosome, start, stop and name are mandatory
(start/stop through int()); score and strand are optional and default to "."; there is no
comment, track or blank-line handling (such lines fail like in the reference); rows keep file
order and duplicates are kept."""
import collections

_Row = collections.namedtuple("_Row", "chromosome start stop name score strand")


class Utr(_Row):
    """One annotation row.  Field names follow the reference's BedEntry so downstream code reads the same."""
    __slots__ = ()

    def getLength(self):
        return self.stop - self.start

    def hasStrand(self):
        return self.strand in ("+", "-")

    def hasNonEmptyName(self):
        return self.name != ""

    def __repr__(self):
        return "\t".join((self.chromosome, str(self.start), str(self.stop), self.name))


def parse_bed_line(line):
    f = line.rstrip().split("\t")
    return Utr(f[0], int(f[1]), int(f[2]), f[3], f[4] if len(f) > 4 else ".", f[5] if len(f) > 5 else ".")


def BedIterator(filename):
    """Rows of a BED file in file order (a generator; the reference exposes a class of the same name)."""
    with open(filename, "r") as fh:
        for line in fh:
            yield parse_bed_line(line)
