"""Multi-GPU sharding of independent units (strips / frames / tiles): host-side only, no collective.

Every unit of the MIC path is self-contained, so ranks take contiguous index ranges balanced by
compressed bytes using the offset tables the containers already carry (PICS parallelstrips.go:115-122,
MIC2 multiframe.go:72-78, MIC3 wsiformat.go:145-155) -- SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def partition_by_bytes(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split units [0, n) into `world` contiguous ranges [lo, hi) with near-equal sums of `lengths`.

    Deterministic, identical on every rank, every unit owned exactly once; trailing ranks may be empty
    when there are fewer units than ranks."""
    n = len(lengths)
    total = sum(int(x) for x in lengths)
    out, lo, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        hi = lo
        while hi < n and (acc + int(lengths[hi]) <= target or hi == lo and n - hi > world - 1 - r):
            acc += int(lengths[hi])
            hi += 1
        if r == world - 1:
            while hi < n:
                acc += int(lengths[hi])
                hi += 1
        out.append((lo, hi))
        lo = hi
    return out


def pics_strip_table(blob: bytes):
    """(width, height, strip_h, [(offset, length)]) of a PICS container (parallelstrips.go:270-290)."""
    import struct

    if len(blob) < 20 or blob[:4] != b"PICS":
        raise ValueError("parallelstrips: invalid magic")
    w, h, n, sh = struct.unpack_from("<4I", blob, 4)
    hdr = 20 + 8 * n
    if len(blob) < hdr:
        raise ValueError("parallelstrips: truncated header")
    tab = [struct.unpack_from("<2I", blob, 20 + 8 * i) for i in range(n)]
    return w, h, sh, [(hdr + o, l) for o, l in tab]
