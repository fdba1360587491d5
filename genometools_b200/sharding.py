"""Bucket-code ranges for -parts and for multi-GPU sharding.

Mirror of gt_suftabparts_new (/root/reference/src/match/sfx-partssuf.c:172-347):
the codes 0..K^pl-1 are cut into `numofparts` contiguous ranges holding about the
same number of suffixes each; the cut points are found on the bucket table with
gt_bcktab_findfirstlarger (/root/reference/src/match/bcktab.c:1322-1381): the
first code whose right border reaches the running target.  Because buckets are
independent sorting problems the concatenation of the parts is the global
suffix table whatever the number of parts (verified against the reference with
-parts 3).
"""
import numpy as np


def suftab_parts(leftborder, numofparts):
    """leftborder: uint array [K^pl + 1], entry c = first SA index of bucket c, last
    entry = number of non-special suffixes.  Returns a list of
    (mincode, maxcode, sa_offset, width) with non-empty width, at most numofparts long."""
    lb = np.asarray(leftborder, dtype=np.int64)
    numofallcodes = lb.shape[0] - 1
    total = int(lb[-1])
    if numofallcodes <= 0:
        return []
    if numofparts <= 1 or total <= numofparts or numofallcodes == 1:
        return [(0, numofallcodes - 1, 0, total)]       # sfx-partssuf.c:211-218
    widthofpart = total // numofparts
    remainder = total % numofparts
    parts = []
    mincode = 0
    target = 0
    for part in range(numofparts):
        target += widthofpart + (1 if part < remainder else 0)
        if part == numofparts - 1:
            maxcode = numofallcodes - 1
        else:
            # first code whose right border (= leftborder[code+1]) is >= target
            maxcode = int(np.searchsorted(lb[1:], target, side="left"))
            maxcode = min(max(maxcode, mincode), numofallcodes - 1)
        if mincode > numofallcodes - 1:
            break
        width = int(lb[maxcode + 1] - lb[mincode])
        if width > 0 or part == numofparts - 1:
            parts.append((mincode, maxcode, int(lb[mincode]), width))
        mincode = maxcode + 1
        if mincode >= numofallcodes:
            break
    # make sure the last listed part reaches the last code (its tail may be empty)
    mn, mx, off, w = parts[-1]
    if mx != numofallcodes - 1:
        parts[-1] = (mn, numofallcodes - 1, off, int(lb[numofallcodes] - lb[mn]))
    return [p for p in parts if p[3] > 0] or [(0, numofallcodes - 1, 0, total)]
