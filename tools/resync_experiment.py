"""CPU-only experiment: can an 8-state MIC strip be decoded speculatively from the middle?

tANS decoders are known to self-synchronise: started in a wrong state at a right bit position they usually fall back
onto the true trajectory after a while, which is what segment-parallel ANS decoders exploit.  Here the eight states of a
frame share ONE bitstream (fse8state.go:230-380), so a speculative decoder that starts at round r0 has to recover the
bit position AND all eight states.  The experiment decodes a bench strip (2577x256, 12 bit, 8-state) with the true
decoder, then restarts it at many rounds r0 with
  (a) the true bit position but unknown states (the most favourable case: the encoder could have stored positions),
  (b) a bit position that is off by a few bits,
and counts the rounds until (P, states) coincide with the true trajectory.

Run:  python tools/resync_experiment.py        (about 8 minutes: big-integer bit extraction; no GPU)
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_dtable(norm, table_log):
    """buildDtable (fsedecompressu16.go:198-263) -> arrays (nbBits, newState) per cell."""
    S = 1 << table_log
    sym_of = np.zeros(S, np.int64)
    high = S - 1
    symnext = {}
    for s, c in enumerate(norm):
        if c == -1:
            sym_of[high] = s
            high -= 1
            symnext[s] = 1
        elif c > 0:
            symnext[s] = int(c)
    step = (S >> 1) + (S >> 3) + 3
    pos = 0
    for s, c in enumerate(norm):
        for _ in range(max(int(c), 0)):
            sym_of[pos] = s
            pos = (pos + step) & (S - 1)
            while pos > high:
                pos = (pos + step) & (S - 1)
    nb = np.zeros(S, np.int64)
    ns = np.zeros(S, np.int64)
    for u in range(S):
        s = int(sym_of[u])
        nxt = symnext[s]
        symnext[s] = nxt + 1
        nbits = table_log - (nxt.bit_length() - 1)
        nb[u] = nbits
        ns[u] = (nxt << nbits) - S
    return nb, ns


def main():
    synth = importlib.import_module("medical-image-codec_b200.synth") if False else None
    import importlib.util

    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "medical-image-codec_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    from oracle.oracle import Oracle

    o = Oracle()
    W, Hs = 2577, 256
    img = synth.xr_image(1, W, 2048)[:Hs].ravel()
    sym = o.delta_rle_compress(img, W, Hs, int(img.max()))
    tlog, symlen, norm = o.fse_table_info(sym)
    frame = o.fse_compress(sym, 8)
    assert frame[0] == 0xFF and frame[1] == 0x84
    count = int.from_bytes(frame[2:6], "little")
    nb, ns = build_dtable(norm[:symlen], tlog)
    nbl, nsl = nb.tolist(), ns.tolist()
    big = int.from_bytes(frame, "little")          # the stream as one little-endian integer (k_ans.cu:26-29)
    P0 = 8 * (len(frame) - 1) + (frame[-1].bit_length() - 1)
    rounds = count // 8

    def run(P, st, r_from, r_to, record=None, truth=None):
        """decode rounds [r_from, r_to); with `truth`, stop at the first round whose (P, states) equal the truth's"""
        for r in range(r_from, r_to):
            if record is not None:
                record.append((P, tuple(st)))
            if truth is not None and truth[r] == (P, tuple(st)):
                return r
            for k in range(8):
                n = nbl[st[k]]
                P -= n
                bits = (big >> P) & ((1 << n) - 1) if P >= 0 else 0
                st[k] = nsl[st[k]] + bits
        return None

    # true trajectory
    P = P0
    st = []
    for k in range(8):
        P -= tlog
        st.append((big >> P) & ((1 << tlog) - 1))
    truth = []
    run(P, list(st), 0, rounds, record=truth)
    print(f"strip {W}x{Hs}: {count} symbols, tableLog {tlog}, {rounds} rounds, {len(frame)} bytes")

    rng = np.random.default_rng(1)
    horizon = 4000
    for label, dp_choices in (("true bit position, unknown states", [0]), ("bit position off by 1..7 bits, unknown states", [1, 2, 3, 5, 7, -1, -3])):
        hits = []
        trials = 60
        for t in range(trials):
            r0 = int(rng.integers(100, rounds - horizon - 10))
            Pt, _ = truth[r0]
            dp = dp_choices[t % len(dp_choices)]
            st0 = [int(x) for x in rng.integers(0, 1 << tlog, 8)]
            hit = run(Pt + dp, st0, r0, r0 + horizon, truth=truth)
            hits.append(None if hit is None else hit - r0)
        ok = sorted(h for h in hits if h is not None)
        print(f"{label}: {len(ok)}/{trials} restarts met the true trajectory within {horizon} rounds"
              + (f"; rounds to meet: min {ok[0]}, median {ok[len(ok) // 2]}, max {ok[-1]}" if ok else ""))
    # how often does ONE state alone recover when the others and the position are right? (the single-state property)
    hits = []
    for t in range(60):
        r0 = int(rng.integers(100, rounds - horizon - 10))
        Pt, stt = truth[r0]
        st0 = list(stt)
        st0[t % 8] = int(rng.integers(0, 1 << tlog))
        hit = run(Pt, st0, r0, r0 + horizon, truth=truth)
        hits.append(None if hit is None else hit - r0)
    ok = sorted(h for h in hits if h is not None)
    print(f"one wrong state of eight (position and the other seven right): {len(ok)}/60 recover"
          + (f"; rounds: min {ok[0]}, median {ok[len(ok) // 2]}, max {ok[-1]}" if ok else ""))


if __name__ == "__main__":
    main()
