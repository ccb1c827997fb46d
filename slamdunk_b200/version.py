"""Format / tool versions written into the outputs.  The values are those of slamdunk 0.4.3
(slamdunk/version.py:19-25): they appear in the tcount header line and gate the filtered BAM
(@PG ID:slamdunk VN), so files are interchangeable with the reference's."""
VERSIONS = {
    "slamdunk": "0.4.3",   # tcount header "#slamdunk v..."
    "bam": "3",            # @PG VN of a filtered BAM that `count` accepts (tcounter.py:156)
    "count": "3",          # count file format
    "ngm": "0.5.5",        # NextGenMap release the MP/XI/NM tag semantics come from
    "b200": "0.1.0",
}
__version__ = VERSIONS["slamdunk"]
__bam_version__ = VERSIONS["bam"]
__count_version__ = VERSIONS["count"]
__ngm_version__ = VERSIONS["ngm"]
__b200_version__ = VERSIONS["b200"]
