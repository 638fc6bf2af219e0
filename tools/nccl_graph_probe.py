"""Probe: does a NCCL all-reduce captured in a CUDA graph replay correctly on this box? (torchrun, 2+ ranks)
Prints one line per stage so a hang can be located from the log."""
import os
import sys
import time

import torch
import torch.distributed as dist


def say(rank, msg):
    print(f"[rank {rank} {time.strftime('%H:%M:%S')}] {msg}", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.full((1 << 20,), float(rank + 1), device=dev)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    say(rank, f"eager all_reduce ok: {t[0].item()}")
    mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            t.fill_(rank + 1.0)
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    say(rank, "side-stream warm-up ok")
    g = torch.cuda.CUDAGraph()
    kw = {} if mode == "global" else {"capture_error_mode": mode}
    with torch.cuda.graph(g, **kw):
        t.mul_(2.0)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        t.add_(1.0)
    say(rank, "capture ok")
    for k in range(3):
        t.fill_(rank + 1.0)
        g.replay()
        torch.cuda.synchronize()
        say(rank, f"replay {k}: {t[0].item()} (expected {2.0 * (world + 1) / 2 + 1})")
    dist.barrier()
    say(rank, "barrier ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
