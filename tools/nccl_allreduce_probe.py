"""All-reduce of the full LAP-3B gradient (3.353 G fp32 = 13.4 GB) alone, in the 512 MB buckets the trainer uses: raw time and
bus bandwidth for the current NCCL environment (NCCL_MAX_NCHANNELS / NCCL_ALGO are read at communicator creation)."""
import os, sys, json
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 3_353_000_000
G = torch.zeros(n, device="cuda")
step = (512 << 20) // 4
def ar():
    works = [dist.all_reduce(G[o:o + step], async_op=True) for o in range(0, n, step)]
    for w in works: w.wait()
for _ in range(2): ar()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ar()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
if rank == 0:
    print(json.dumps({"world": world, "nch": os.environ.get("NCCL_MAX_NCHANNELS"), "algo": os.environ.get("NCCL_ALGO"),
                      "ms": ms, "algbw_gbs": n * 4 / ms / 1e6, "busbw_gbs": n * 4 / ms / 1e6 * 2 * (world - 1) / world}), flush=True)
dist.destroy_process_group()
