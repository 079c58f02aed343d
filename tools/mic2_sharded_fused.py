"""torchrun script: like tools/mic2_sharded.py, but the carry exchange is ONE kernel per rank that reads the last frames of
the earlier ranks straight out of their GPUs' memory over NVLink (CUDA IPC handles exchanged once), instead of an NCCL
all-gather followed by adds.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mic2_sharded_fused.py"""
import importlib, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
shard = importlib.import_module("medical-image-codec_b200.shard")
synth = importlib.import_module("medical-image-codec_b200.synth")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, NF = 1996, 2457, int(os.environ.get("FRAMES", "24"))
st = synth.tomo_stack(7, NF, W, H)
blob = np.frombuffer(mic.CompressMultiFrame(st.reshape(NF, -1), W, H, 1023, True), np.uint8)
w, h, n, temporal, tab = shard.mic2_frame_table(blob.tobytes())
lo, hi = shard.partition_by_bytes([l for _, l in tab], world)[rank]
fpx = w * h
d_comp = torch.zeros(blob.size + 256, dtype=torch.uint8, device="cuda")
d_comp[: blob.size] = torch.from_numpy(blob.copy()).cuda()
d_out = torch.zeros(max(hi - lo, 1) * fpx, dtype=torch.int16, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
dec = mic.Decoder(local)
dec.begin()
if hi > lo:
    dec.add_mic2_range(blob, 0, 0, lo, hi - lo)
dec.commit()
# exchange buffer: this rank's last frame, in a whole cudaMalloc allocation so that it can be exported
xbuf = mic.device_alloc(fpx * 2)
handles = [None] * world
dist.all_gather_object(handles, mic.ipc_export(xbuf))
peers = [mic.ipc_open(handles[q]) if q != rank else xbuf for q in range(world)]
from cuda import cudart        # cuda-python: one device-to-device copy of the last frame into the exported buffer


def d2d(dst, src, nbytes, stream):
    (err,) = cudart.cudaMemcpyAsync(dst, src, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, stream)
    assert int(err) == 0, err


times = []
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    if hi > lo:
        dec.run_device(d_comp.data_ptr(), blob.size, d_out.data_ptr(), (hi - lo) * fpx, stream)
        src = d_out.data_ptr() + (hi - lo - 1) * fpx * 2
        d2d(xbuf, src, fpx * 2, stream)                  # last frame -> exchange buffer
    else:
        torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize(); dist.barrier()           # every rank's last frame is in place
    if rank > 0 and hi > lo:
        mic.temporal_add_carry_peers(d_out.data_ptr(), peers[:rank], fpx, hi - lo, stream)   # reads ranks 0..rank-1 over NVLink
    torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
ok = np.array_equal(d_out.cpu().numpy().view(np.uint16)[: (hi - lo) * fpx], st[lo:hi].ravel())
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print({"frames": NF, "ranks": world, "bit_exact_on_all_ranks": bool(flag.item()), "ms": round(min(times) * 1e3, 2),
           "exchange": "one kernel per rank reading peer memory (CUDA IPC over NVLink)", "bytes_read_from_peers_by_last_rank": (world - 1) * fpx * 2})
dist.barrier()
for q in range(world):
    if q != rank:
        mic.ipc_close(peers[q])
mic.device_free(xbuf)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
