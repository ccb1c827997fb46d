"""Minimal stand-in for intervaltree.IntervalTree (oracle shim).
Semantics used by the reference (utils/BedReader.py:20-30, dunks/filter.py:136-143):
tree[a:b] = data inserts the half-open interval [a, b); tree[a:b] returns the set of
Interval objects overlapping [a, b) (begin < b and end > a).

The real library returns a `set` of Intervals whose iteration order depends on the (per-process
randomized) string hash of the UTR name, so the reference's choice of the retained multimapper
alignment is arbitrary whenever ONE alignment overlaps two differently named UTRs
(dunks/filter.py:147-160, dumpBufferToBam :56-65).  The reference only takes len() of the
result and iterates it; this shim returns the hits as a list in BED (insertion) order, which is
one of the orders the real library can produce, and makes the oracle deterministic."""
import collections

Interval = collections.namedtuple("Interval", ["begin", "end", "data"])


class IntervalTree(object):
    def __init__(self):
        self._ivs = []
        self._sorted = True

    def __setitem__(self, sl, data):
        if sl.start >= sl.stop:
            raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree")
        self._ivs.append(Interval(sl.start, sl.stop, data))

    def __getitem__(self, sl):
        if isinstance(sl, slice):
            a, b = sl.start, sl.stop
            if a >= b:
                return []
            return [iv for iv in self._ivs if iv.begin < b and iv.end > a]
        return [iv for iv in self._ivs if iv.begin <= sl < iv.end]

    def __len__(self):
        return len(self._ivs)

    def __iter__(self):
        return iter(self._ivs)
