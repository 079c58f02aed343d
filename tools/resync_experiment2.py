"""CPU-only experiment: do N-state MIC frames self-synchronise when the criterion is the right one?

tools/resync_experiment.py compared a restarted decoder with the true one at the same ROUND index.  A decoder that starts
in wrong states consumes a different number of bits per symbol, so it reaches a given bit position after a different
number of symbols; what a segment-parallel decoder needs is weaker: at some point the restarted decoder must be at a bit
position P, about to decode with state k, holding exactly the states the true decoder held when IT was at P about to
decode with state k.  From then on both produce the same symbols (the symbol index offset is recovered afterwards by
counting).  This script records the true trajectory per SYMBOL keyed by (P, k) and measures how many symbols a decoder
restarted at a true (P, k = 0) with random states needs to land on it, for N = 1, 2, 4, 8.

Run:  python tools/resync_experiment2.py   (a few minutes, no GPU)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from resync_experiment import build_dtable  # noqa: E402


def main():
    import importlib.util

    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "medical-image-codec_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    from oracle.oracle import Oracle

    o = Oracle()
    W, Hs = 2577, 64
    img = synth.xr_image(1, W, 2048)[300:300 + Hs].ravel()
    sym = o.delta_rle_compress(img, W, Hs, 4095)
    tlog, symlen, norm = o.fse_table_info(sym)
    nb, ns = build_dtable(norm[:symlen], tlog)
    nbl, nsl = nb.tolist(), ns.tolist()
    rng = np.random.default_rng(3)
    for N in (1, 2, 4, 8):
        frame = o.fse_compress(sym, N)
        if N == 1:
            body, count = frame, None
        else:
            assert frame[0] == 0xFF
            body, count = frame, int.from_bytes(frame[2:6], "little")
        big = int.from_bytes(body, "little")
        P = 8 * (len(body) - 1) + (body[-1].bit_length() - 1)
        st = []
        for k in range(N):
            P -= tlog
            st.append((big >> P) & ((1 << tlog) - 1))
        nsym = count if count is not None else len(sym)
        # true trajectory per symbol
        truth = {}
        order = []
        for i in range(nsym - N):
            k = i % N
            truth[(P, k)] = tuple(st)
            order.append((P, k))
            n = nbl[st[k]]
            P -= n
            st[k] = nsl[st[k]] + ((big >> P) & ((1 << n) - 1))
        horizon = 40000
        res = []
        trials = 30
        for t in range(trials):
            i0 = int(rng.integers(1000, max(1001, len(order) - horizon - 10))) // N * N
            P0, k0 = order[i0]
            assert k0 == 0
            s = [int(x) for x in rng.integers(0, 1 << tlog, N)]
            P, hit = P0, None
            for j in range(min(horizon, len(order) - i0 - 1)):
                k = j % N
                if truth.get((P, k)) == tuple(s):
                    hit = j
                    break
                n = nbl[s[k]]
                P -= n
                if P < 0:
                    break
                s[k] = nsl[s[k]] + ((big >> P) & ((1 << n) - 1))
            res.append(hit)
        ok = sorted(h for h in res if h is not None)
        msg = f"{N}-state, tableLog {tlog}, {nsym} symbols: {len(ok)}/{trials} restarts (true P, random states) land on the true trajectory"
        if ok:
            msg += f"; symbols needed: min {ok[0]}, median {ok[len(ok) // 2]}, p90 {ok[int(len(ok) * 0.9)]}, max {ok[-1]}"
        print(msg, flush=True)


if __name__ == "__main__":
    main()
