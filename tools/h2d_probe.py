#!/usr/bin/env python
"""Host->device bandwidth per rank when N ranks copy at the same time (the e2e leg of bench.py is PCIe-bound).

    python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py [--mode default|affinity] [--gb 2]

Prints one JSON line per rank: GPU PCI bus id, its NUMA node, the CPUs the rank ran on, GB/s alone (ranks take turns) and
GB/s with all ranks copying together.  `affinity` pins the rank's threads to the GPU's NUMA node BEFORE the pinned buffer is
allocated (first touch places the pages there).
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def gpu_numa(bus_id):
    # torch: '00000000:1B:00.0' -> sysfs '0000:1b:00.0'
    b = bus_id.lower()
    if len(b.split(":")[0]) == 8:
        b = b[4:]
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % b).read())
    except Exception:
        node = -1
    cpus = None
    if node >= 0:
        try:
            cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        except Exception:
            pass
    return node, cpus


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="default")
    ap.add_argument("--gb", type=float, default=2.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import ctypes
    bus = torch.cuda.get_device_properties(lr).pci_bus_id if hasattr(torch.cuda.get_device_properties(lr), "pci_bus_id") else None
    if bus is None:
        import subprocess
        bus = subprocess.check_output(["nvidia-smi", "-i", str(lr), "--query-gpu=pci.bus_id", "--format=csv,noheader"]).decode().strip()
    elif isinstance(bus, int):
        import subprocess
        bus = subprocess.check_output(["nvidia-smi", "-i", str(lr), "--query-gpu=pci.bus_id", "--format=csv,noheader"]).decode().strip()
    node, cpus = gpu_numa(bus)
    if args.mode == "affinity" and cpus:
        allowed = sorted(set(parse_cpulist(cpus)) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    n = int(args.gb * (1 << 30))
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()

    def once():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            d.copy_(h, non_blocking=True)
            e1.record()
        e1.synchronize()
        return n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    once()
    alone = None
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = max(once() for _ in range(3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 6
    with torch.cuda.stream(s):
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
    s.synchronize()
    together = reps * n / (time.perf_counter() - t0) / 1e9
    print(json.dumps({"rank": rank, "mode": args.mode, "bus": bus, "numa": node, "node_cpus": cpus, "affinity": len(os.sched_getaffinity(0)),
                      "alone_gbs": round(alone, 2), "together_gbs": round(together, 2)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
