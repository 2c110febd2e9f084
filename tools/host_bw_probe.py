"""Pinned device-to-host bandwidth of the box, no kernels: what bounds the end-to-end Gram leg.

    python tools/host_bw_probe.py                       # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_bw_probe.py

Every rank copies `--gib` GiB from its GPU into its OWN pinned host buffer, all ranks at once
(barrier on both sides), in chunks of `--chunk-mib` like pncov_gram's row blocks.  Prints one JSON
line: per-rank GB/s (min / max) and the aggregate = total bytes / slowest rank's time.  With
`--bind` each rank first binds itself to the CPUs NVML reports as local to its GPU (what bench.py
does), so the pinned buffer is allocated on the GPU's own NUMA node.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--chunk-mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--bind", dest="numa", action="store_true")
    ap.add_argument("--h2d", action="store_true", help="host-to-device instead")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if args.numa:
        import bench  # noqa: PLC0415

        bench._bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("gloo")
    nbytes = int(args.gib * (1 << 30))
    chunk = args.chunk_mib << 20
    t0 = time.perf_counter()
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    t_pin = time.perf_counter() - t0
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.fill_(1)
    torch.cuda.synchronize()
    best = None
    for _ in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for o in range(0, nbytes, chunk):
            if args.h2d:
                dev[o:o + chunk].copy_(host[o:o + chunk], non_blocking=True)
            else:
                host[o:o + chunk].copy_(dev[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t = torch.tensor([best, t_pin], dtype=torch.float64)
    if world > 1:
        gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, t)
    else:
        gathered = [t]
    if rank == 0:
        times = [float(g[0]) for g in gathered]
        pins = [float(g[1]) for g in gathered]
        print(json.dumps({
            "probe": "pinned " + ("H2D" if args.h2d else "D2H"), "n_gpus": world, "gib_per_rank": args.gib,
            "chunk_mib": args.chunk_mib, "numa_bound": bool(args.numa),
            "per_rank_gbs_min": nbytes / max(times) / 1e9, "per_rank_gbs_max": nbytes / min(times) / 1e9,
            "aggregate_gbs": world * nbytes / max(times) / 1e9,
            "pin_seconds_per_gib": max(pins) / args.gib,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
