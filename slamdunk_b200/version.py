"""Version constants that appear in the output files; same values as the reference's
slamdunk/version.py:19-25 so the tcount header and the BAM @PG gate are interchangeable."""
__version__ = "0.4.3"
__bam_version__ = "3"
__count_version__ = "3"
__ngm_version__ = "0.5.5"
__b200_version__ = "0.1.0"
