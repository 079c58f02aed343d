"""torchrun script: one temporal MIC2 stack decoded by N ranks (one GPU each) with the path's single exchange step.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mic2_sharded.py

Every rank plans its own frame range from the container's frame table (balanced by compressed bytes), decodes it on
its GPU, all-gathers ONE frame per rank over NCCL, adds its carry, and rank 0 checks the union against the source."""
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
blob = np.frombuffer(mic.CompressMultiFrame(st.reshape(NF, -1), W, H, 1023, True), np.uint8)   # product encoder, same bytes on every rank
w, h, n, temporal, tab = shard.mic2_frame_table(blob.tobytes())
lo, hi = shard.partition_by_bytes([l for _, l in tab], world)[rank]
fpx = w * h
d_comp = torch.zeros(blob.size + 256, dtype=torch.uint8, device="cuda")
d_comp[: blob.size] = torch.from_numpy(blob.copy()).cuda()     # a rank only needs the bytes of its own frames
d_out = torch.zeros(max(hi - lo, 1) * fpx, dtype=torch.int16, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
dec = mic.Decoder(local)
dec.begin()
if hi > lo:
    dec.add_mic2_range(blob, 0, 0, lo, hi - lo)
dec.commit()
times = []
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    if hi > lo:
        dec.run_device(d_comp.data_ptr(), blob.size, d_out.data_ptr(), (hi - lo) * fpx, stream)
    last = d_out[(hi - lo - 1) * fpx:(hi - lo) * fpx] if hi > lo else torch.zeros(fpx, dtype=torch.int16, device="cuda")
    carry = shard.mic2_temporal_carry(shard.all_gather_last_frames(dist, last), rank)      # the one collective
    if carry is not None and hi > lo:
        mic.temporal_add_carry(d_out.data_ptr(), carry.contiguous().data_ptr(), fpx, hi - lo, stream)
    torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
ok = np.array_equal(d_out.cpu().numpy().view(np.uint16)[: (hi - lo) * fpx], st[lo:hi].ravel())
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print({"frames": NF, "ranks": world, "bit_exact_on_all_ranks": bool(flag.item()), "ms": round(min(times) * 1e3, 2),
           "GBps_raw": round(NF * fpx * 2 / min(times) / 1e9, 3), "exchange_bytes_per_rank": fpx * 2})
dist.destroy_process_group()
sys.exit(0 if ok else 1)
