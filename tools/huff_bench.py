"""Kernel time of the Huffman decode (k_huff_decode) on one long stream and on a mammogram's Delta+RLE symbols; run with
MICGPU_HUFF_SERIAL=1 for the one-thread walk.  The stream comes from the oracle's encoder (test infrastructure)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
from oracle.oracle import Oracle
orc = Oracle()
W, H = 2577, 2048
img = synth.xr_image(1, W, H).ravel()
t0 = time.time(); blob = np.frombuffer(orc.delta_rle_huff_compress(img, W, H, int(img.max())), np.uint8); t_enc = time.time() - t0
t0 = time.time(); ref = orc.delta_rle_huff_decompress(blob, W, H); t_cpu = time.time() - t0
d_comp = torch.zeros(blob.size + 256, dtype=torch.uint8, device="cuda"); d_comp[: blob.size] = torch.from_numpy(blob.copy()).cuda()
d_out = torch.zeros(W * H, dtype=torch.int16, device="cuda")
dec = mic.Decoder(0); dec.begin(); dec.add_huff_unit(blob, 0, 0, W, H, 0); dec.commit()
dec.set_profiling(True)
s = torch.cuda.current_stream().cuda_stream
for it in range(3):
    dec.run_device(d_comp.data_ptr(), blob.size, d_out.data_ptr(), W * H, s)
    kt = dec.kernel_times()
ok = np.array_equal(d_out.cpu().numpy().view(np.uint16), img)
nsym = int.from_bytes(bytes(blob[:4]), "big")
print({"image": f"{W}x{H}", "ratio": round(W * H * 2 / blob.size, 3), "symbols": nsym, "exact": ok, "oracle_decode_ms": round(t_cpu * 1e3, 1),
       "serial": os.environ.get("MICGPU_HUFF_SERIAL", "0"), "kernels_ms": [(k, round(v, 3)) for k, v in kt]})
