"""CPU-side checks of the boundary: libmicgpu.so loads and exports every symbol include/micgpu.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side shard planner agrees across ranks."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "micgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:micgpu|mic)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(mic):
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(mic.lib, n), f"libmicgpu.so does not export {n}"


def test_twin_signatures_match_reference_headers():
    """The 8 drop-in symbols keep the reference prototypes (ojph/mic_decompress_c.h:24-49, mic_parallel.h:49-55)."""
    txt = re.sub(r"\s+", " ", open(os.path.join(ROOT, "include", "micgpu.h")).read())
    for n in ("two", "four", "eight"):
        for s in ("", "_simd"):
            assert f"int mic_decompress_{n}_state{s}(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);" in txt
    assert "int mic_decompress_parallel(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height, int max_threads);" in txt


def test_no_cpu_fallback_without_gpu(mic):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert mic.lib.micgpu_device_count() == 0
    assert not mic.lib.micgpu_decoder_create(0)
    assert b"no CUDA device" in mic.lib.micgpu_last_error()
    out = np.zeros(16, np.uint16)
    blob = np.zeros(64, np.uint8)
    rc = mic.lib.micgpu_decompress_single_frame(blob.ctypes.data, blob.size, out.ctypes.data, 4, 4)
    assert rc == mic.api.E_CUDA
    with pytest.raises(mic.MicGpuError):
        mic.DecompressSingleFrame(bytes(64), 4, 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "medical-image-codec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("touches oracle/", "").lower().replace("the oracle is", ""), f
    out = subprocess.run(["nm", "-D", os.path.join(pkg, "libmicgpu.so")], capture_output=True, text=True).stdout
    assert "orc_" not in out


def test_partition_by_bytes(synth):
    import importlib

    shard = importlib.import_module("medical-image-codec_b200.shard")
    lens = [100, 300, 50, 50, 400, 100, 100, 100]
    for world in (1, 2, 3, 4, 8, 16):
        parts = shard.partition_by_bytes(lens, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(lens)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    p2 = shard.partition_by_bytes(lens, 2)
    assert abs(sum(lens[p2[0][0]:p2[0][1]]) - 600) <= 400
    assert shard.partition_by_bytes([], 4) == [(0, 0)] * 4


_GLOO_WORKER = r'''
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shard = importlib.import_module("medical-image-codec_b200.shard")
synth = importlib.import_module("medical-image-codec_b200.synth")
from oracle.oracle import Oracle
o = Oracle()
# every rank derives the same strip table from the same PICS blob and takes its own range
img = synth.xr_image(5, 211, 160).ravel()
blob = o.pics_compress(img, 211, 160, int(img.max()), 8, 2)
w, h, sh, tab = shard.pics_strip_table(blob)
parts = shard.partition_by_bytes([l for _, l in tab], world)
lo, hi = parts[rank]
mine = np.zeros(w * h, np.int64)
for s in range(lo, hi):
    off, ln = tab[s]
    rows = min(sh, h - s * sh)
    px = o.decompress_single_frame(blob[off:off + ln], w, rows)      # CPU stand-in for the per-rank GPU decode
    mine[s * sh * w:(s * sh + rows) * w] = px
owned = torch.zeros(len(tab), dtype=torch.int64)
owned[lo:hi] = 1
dist.all_reduce(owned)                                              # test-only check; the data path has no collective
t = torch.from_numpy(mine)
dist.all_reduce(t)
ok = bool((owned == 1).all()) and np.array_equal(t.numpy().astype(np.uint16), img)
print("RANK", rank, "OK" if ok else "FAIL", parts)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 path: ranks agree on the partition, own every strip exactly once, and
    the union of their outputs is the image."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(script), ROOT], capture_output=True, text=True, env=env, timeout=170)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


_GLOO_TEMPORAL_WORKER = r'''
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shard = importlib.import_module("medical-image-codec_b200.shard")
synth = importlib.import_module("medical-image-codec_b200.synth")
from oracle.oracle import Oracle
o = Oracle()
# a temporal MIC2 stack; every rank reads the same frame table and takes a contiguous range balanced by bytes
st = synth.tomo_stack(3, 7, 96, 80)
blob = o.mic2_compress(st.ravel(), 96, 80, 1023, True)
w, h, n, temporal, tab = shard.mic2_frame_table(blob)
assert temporal and n == 7
lo, hi = shard.partition_by_bytes([l for _, l in tab], world)[rank]
fpx = w * h
# CPU stand-in for the per-rank GPU decode (micgpu_decoder_add_mic2_range): absolute pixels for the range that starts at
# frame 0, running sums of UnZigZag(residual) relative to a ZERO carry for a later range
full = o.mic2_decompress(blob)[0].reshape(n, fpx).astype(np.uint16)
local = full[lo:hi].copy()
if lo > 0:
    local = (local - full[lo - 1]).astype(np.uint16)
# the one exchange step: all-gather the last local frame of every rank, exclusive scan mod 2^16
last = torch.from_numpy((local[-1] if hi > lo else np.zeros(fpx, np.uint16)).astype(np.int16))
carry = shard.mic2_temporal_carry(shard.all_gather_last_frames(dist, last), rank)
if carry is not None:
    local = (local + carry.numpy().astype(np.uint16)).astype(np.uint16)
ok = np.array_equal(local, full[lo:hi]) and np.array_equal(full.reshape(-1), st.ravel())
print("RANK", rank, "OK" if ok else "FAIL", (lo, hi))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_temporal_carry_two_ranks_gloo(tmp_path):
    """The path's one exchange step: a temporal MIC2 stack cut across two ranks needs the carry frame of the
    earlier range; all_gather of one frame per rank + exclusive scan mod 2^16 reproduces the serial decode."""
    script = tmp_path / "worker_t.py"
    script.write_text(_GLOO_TEMPORAL_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29614")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29614", str(script), ROOT], capture_output=True, text=True, env=env, timeout=170)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


def test_temporal_carry_is_an_exclusive_scan():
    import importlib

    shard = importlib.import_module("medical-image-codec_b200.shard")
    rng = np.random.default_rng(3)
    lasts = rng.integers(0, 65536, size=(4, 50)).astype(np.uint16)
    assert shard.mic2_temporal_carry(lasts, 0) is None
    for r in range(1, 4):
        want = lasts[:r].astype(np.uint64).sum(axis=0).astype(np.uint16)      # mod 2^16
        assert np.array_equal(shard.mic2_temporal_carry(lasts, r), want)


def test_partition_twin_matches_python(tmp_path):
    """micgpu_partition_by_bytes is the C++ twin of shard.partition_by_bytes: same cuts on ragged size lists,
    more parts than units, empty lists.  Pure host code: runs without a GPU."""
    import ctypes as C
    import importlib

    import numpy as np

    shard = importlib.import_module("medical-image-codec_b200.shard")
    lib = C.CDLL(os.path.join(ROOT, "medical-image-codec_b200", "libmicgpu.so"))
    lib.micgpu_partition_by_bytes.argtypes = [C.POINTER(C.c_uint64), C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(3)
    cases = [[], [5], [1, 1, 1], [10, 1, 1, 1, 1, 1, 30, 2], list(rng.integers(1, 10**6, 257)), [7] * 64, list(rng.integers(1, 50, 5))]
    for sizes in cases:
        for parts in (1, 2, 3, 4, 8, 11):
            arr = (C.c_uint64 * max(len(sizes), 1))(*[int(x) for x in sizes])
            cuts = (C.c_uint64 * (parts + 1))()
            assert lib.micgpu_partition_by_bytes(arr, len(sizes), parts, cuts) == 0
            want = shard.partition_by_bytes([int(x) for x in sizes], parts)
            got = [(int(cuts[i]), int(cuts[i + 1])) for i in range(parts)]
            assert got == want, (sizes[:8], parts, got, want)


def test_integration_doc_names_only_exported_symbols():
    """INTEGRATION.md shows the cgo binding a maintainer adds: every `micgpu_*` / `mic_*compress*` name it mentions must be
    declared in include/micgpu.h (and therefore exported: test_every_declared_symbol_is_exported)."""
    import re

    hdr = open(os.path.join(ROOT, "include", "micgpu.h")).read()
    declared = set(re.findall(r"\b(micgpu_[a-z0-9_]+|mic_(?:de)?compress_[a-z_]+)\s*\(", hdr))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    used = set(re.findall(r"\b(micgpu_[a-z0-9_]+)\b", doc))
    # families written with braces or wildcards in the prose (micgpu_wsi_plan_*, micgpu_pica_{de,}compress, ...) and plain words
    skip = {n for n in used if n.endswith("_") or n in ("micgpu_h", "micgpu_go")}
    missing = sorted(n for n in used - skip if n not in declared and not any(d.startswith(n) for d in declared))
    assert not missing, missing
