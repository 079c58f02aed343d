"""Multi-GPU sharding of independent units (strips / frames / tiles): host-side only, no collective -- except the one
exchange step the path has, the carry frame of a temporal MIC2 stack cut across ranks (mic2_temporal_carry below).

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


def mic2_frame_table(blob: bytes):
    """(width, height, frames, temporal, [(offset, length)]) of a MIC2 container (multiframe.go:96-142)."""
    import struct

    if len(blob) < 20 or blob[:4] != b"MIC2":
        raise ValueError("MIC2: invalid magic")
    w, h, n = struct.unpack_from("<3I", blob, 4)
    temporal = bool(blob[16] & 2)
    hdr = 20 + 8 * n
    if len(blob) < hdr:
        raise ValueError("MIC2: truncated header")
    tab = [struct.unpack_from("<2I", blob, 20 + 8 * i) for i in range(n)]
    return w, h, n, temporal, [(hdr + o, l) for o, l in tab]


def mic2_temporal_carry(last_frames, rank: int):
    """Carry frame of `rank` from the LAST frame of every rank's locally decoded range (a stacked array-like
    [world, frame_px], uint16 values; numpy array or torch tensor).

    Rank 0 decodes absolute pixels; every later range decodes running sums relative to a zero carry, so its last frame
    is the sum of its residuals.  out_f = out_(f-1) + UnZigZag(res_f) mod 2^16 is associative, hence
    carry(r) = sum over q < r of last_frames[q]  (mod 2^16) -- an exclusive scan of ONE frame per rank boundary
    (SURVEY.md section 8(e)).  Ranks with an empty range contribute a zero frame.  Returns None for rank 0."""
    if rank == 0:
        return None
    acc = last_frames[0]
    try:  # torch: int16 arithmetic wraps, which is exactly mod 2^16
        acc = acc.clone()
        for q in range(1, rank):
            acc += last_frames[q]
        return acc
    except AttributeError:
        import numpy as np

        acc = np.array(last_frames[0], dtype=np.uint16, copy=True)
        for q in range(1, rank):
            acc = (acc + np.asarray(last_frames[q], dtype=np.uint16)).astype(np.uint16)
        return acc


def all_gather_last_frames(dist, last):
    """all_gather of one frame per rank for mic2_temporal_carry.  `last` is an int16/uint16 torch tensor (host for
    gloo, device for NCCL); it travels as bytes because neither backend moves 16-bit integers."""
    import torch

    world = dist.get_world_size()
    raw = last.contiguous().view(torch.uint8)
    out = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(out, raw)
    return torch.stack([o.view(torch.int16) for o in out])
