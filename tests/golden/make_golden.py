"""Regenerates tests/golden/streams.json: length + sha256 of the compressed streams the CPU oracle and the
compiled reference C twin (oracle/_ref, built from /root/reference/ojph/*.c) produce for fixed inputs.

Run in the build container (needs /root/reference for the twin):  python tests/golden/make_golden.py
The reference holds no golden compressed vectors of its own; these pin the oracle to
the reference *binary* for the 2/4/8-state Delta+RLE+FSE path, and pin everything else against regressions."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, RefTwin  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    o, r = Oracle(), RefTwin()
    synth = importlib.import_module("medical-image-codec_b200.synth")
    out = {}
    inputs = {
        "CT_512_512": (np.fromfile(os.path.join(HERE, "CT_512_512_image.bin"), np.uint16), 512, 512),
        "MR_256_256": (np.fromfile(os.path.join(HERE, "MR_256_256_image.bin"), np.uint16), 256, 256),
        "xr_seed1_611x403": (synth.xr_image(1, 611, 403).ravel(), 611, 403),
    }
    for name, (img, w, h) in inputs.items():
        mx = int(img.max())
        ent = {"width": w, "height": h, "max": mx, "pixels_sha256": sha(img.tobytes())}
        for ns in (2, 4, 8):
            a = o.compress_single_frame(img, w, h, mx, ns)
            b = r.compress(img, w, h, ns)
            assert a == b, (name, ns)
            ent[f"frame_{ns}state"] = {"len": len(a), "sha256": sha(a), "source": "reference C twin == oracle"}
        a = o.compress_single_frame(img, w, h, mx, 1)
        ent["frame_1state"] = {"len": len(a), "sha256": sha(a), "source": "oracle (restatement of fsecompressu16.go)"}
        sym = o.delta_rle_compress(img, w, h, mx)
        ent["delta_rle_symbols"] = {"len": int(sym.size), "sha256": sha(sym.tobytes())}
        a = o.fse_compress(sym, 108)
        ent["frame_rans8"] = {"len": len(a), "sha256": sha(a), "source": "oracle (restatement of rans8state.go)"}
        a = o.pics_compress(img, w, h, mx, 8, 2)
        ent["pics8_2state"] = {"len": len(a), "sha256": sha(a), "source": "oracle; decoded by reference mic_decompress_parallel"}
        assert np.array_equal(r.decompress_parallel(a, w, h, 8), img)
        a = o.wavelet_v2_compress(img, h, w, mx, 5)
        ent["wavelet_v2_5lv"] = {"len": len(a), "sha256": sha(a), "source": "oracle (restatement of waveletfsecompressu16.go)"}
        out[name] = ent
    st = synth.tomo_stack(7, 5, 128, 128)
    for temporal in (False, True):
        a = o.mic2_compress(st.ravel(), 128, 128, 1023, temporal)
        out[f"mic2_tomo_seed7_5x128x128_temporal{int(temporal)}"] = {"len": len(a), "sha256": sha(a), "source": "oracle"}
    rgb = synth.wsi_region(11, 300, 200, 512, 384, 1024, 768)
    a = o.wsi_compress(rgb.ravel(), 512, 384, 3, 8, 256, 256, 0)
    out["mic3_wsi_seed11_512x384"] = {"len": len(a), "sha256": sha(a), "pixels_sha256": sha(rgb.tobytes()), "source": "oracle"}
    json.dump(out, open(os.path.join(HERE, "streams.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
