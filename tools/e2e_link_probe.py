"""Dev probe (torchrun on N GPUs): what the box's host link gives when every rank does NOTHING but the per-tick transfers of
bench.py's end-to-end figure -- 12 MB device->host (positions) and, at N = 1, 12 MB host->device (forces) per "tick" through a
4-deep ring of pinned buffers on copy streams -- so that the e2e numbers at N > 1 can be read against the link's own ceiling.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 tools/e2e_link_probe.py"""
import os, time, json
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0")); N = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(lr)
if N > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
MB = 12_000_000
ring = 4
dev_out = [torch.empty(MB, dtype=torch.uint8, device="cuda") for _ in range(ring)]
dev_in = [torch.empty(MB, dtype=torch.uint8, device="cuda") for _ in range(ring)]
host_out = [torch.empty(MB, dtype=torch.uint8).pin_memory() for _ in range(ring)]
host_in = [torch.empty(MB, dtype=torch.uint8).pin_memory() for _ in range(ring)]
s_d2h, s_h2d = torch.cuda.Stream(), torch.cuda.Stream()
def run(ticks, up):
    if N > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(ticks):
        b = k % ring
        with torch.cuda.stream(s_d2h): host_out[b].copy_(dev_out[b], non_blocking=True)
        if up:
            with torch.cuda.stream(s_h2d): dev_in[b].copy_(host_in[b], non_blocking=True)
    torch.cuda.synchronize()
    if N > 1: dist.barrier()
    return (time.perf_counter() - t0) / ticks
run(50, True)
res = {}
for name, up in (("d2h_only", False), ("d2h_and_h2d", True)):
    t = run(400, up)
    tt = torch.tensor([t], device="cuda", dtype=torch.float64)
    if N > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    res[name] = {"ms_per_tick": t * 1e3, "ticks_per_s_per_rank": 1.0 / t, "aggregate_GBps": N * MB * (2 if up else 1) / t / 1e9}
if rank == 0: print(json.dumps({"n_gpus": N, "bytes_per_tick_per_rank_each_way": MB, **res}))
if N > 1: dist.destroy_process_group()
