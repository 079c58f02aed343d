"""Per-kernel times of a MIC2 stack decode (temporal and independent), device resident."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
W, H, NF = 1996, 2457, int(os.environ.get("FRAMES", "24"))
st = synth.tomo_stack(7, NF, W, H)
for temporal in (False, True):
    blob = np.frombuffer(mic.CompressMultiFrame(st.reshape(NF, -1), W, H, 1023, temporal), np.uint8)
    d_comp = torch.zeros(blob.size + 256, dtype=torch.uint8, device="cuda"); d_comp[: blob.size] = torch.from_numpy(blob.copy()).cuda()
    d_out = torch.zeros(NF * W * H, dtype=torch.int16, device="cuda")
    dec = mic.Decoder(0); dec.begin(); print(dec.add_mic2(blob, 0, 0)); dec.commit()
    s = torch.cuda.current_stream().cuda_stream
    dec.set_profiling(True)
    for it in range(2):
        dec.run_device(d_comp.data_ptr(), blob.size, d_out.data_ptr(), NF * W * H, s)
        kt = dec.kernel_times()
    print("temporal" if temporal else "independent", "ratio %.2f" % (NF * W * H * 2 / blob.size), [(k, round(v, 2)) for k, v in kt])
    ok = np.array_equal(d_out.cpu().numpy().view(np.uint16), st.ravel()); print("exact", ok)
    # frame headers: state count / tableLog of the first frames
    import struct
    n = struct.unpack_from("<I", blob, 12)[0]; hdr = 20 + 8 * n
    for i in (0, 1, 2):
        o, l = struct.unpack_from("<2I", blob, 20 + 8 * i); f = bytes(blob[hdr + o: hdr + o + 8])
        print("  frame", i, "len", l, "head", f.hex())
