"""Stand-in for matplotlib (absent from this image): the reference's inference driver imports
``matplotlib.pyplot`` (/root/reference/test_region_grow.py:16) but never calls it."""
