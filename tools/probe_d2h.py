#!/usr/bin/env python
"""Host ceiling of the end-to-end step: how fast can N ranks (one per GPU) pull one step's outputs to pinned host memory
at the same time?

    python tools/probe_d2h.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_d2h.py [--bind]

Each rank copies `--mb` MiB (default 15.6 = the 16.4 MB of obs + reward + cost + done of 65,536 point environments)
device -> pinned host, `--reps` times back to back, and the aggregate GB/s over all ranks is printed as one JSON line
(max over ranks of the elapsed time).  `--bind` pins every rank to the cores of its GPU's NUMA node (distinct cores per
rank) BEFORE the pinned allocation, so that the pages are first-touched on that node -- the same binding bench.py applies.
The e2e number of bench.py is reported as a fraction of this ceiling (`e2e.host_ceiling_frac`).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def numa_node_of_gpu(index):
    """NUMA node of a CUDA device from sysfs (None if unknown)"""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        return node if node >= 0 else None
    except Exception:
        return None


def cpus_of_node(node):
    try:
        out = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return []


def bind_rank_to_gpu_node(local_rank, local_world):
    """Restrict this process to its share of the cores of its GPU's NUMA node; returns a description for the JSON line."""
    allowed = sorted(os.sched_getaffinity(0))
    node = numa_node_of_gpu(local_rank)
    cores = [c for c in (cpus_of_node(node) if node is not None else []) if c in allowed] or allowed
    # ranks whose GPUs sit on the same node split its cores
    same = [r for r in range(local_world) if numa_node_of_gpu(r) == node] or [local_rank]
    k, m = same.index(local_rank) if local_rank in same else 0, len(same)
    share = cores[k * len(cores) // m:(k + 1) * len(cores) // m] or cores
    try:
        os.sched_setaffinity(0, share)
    except Exception:
        share = allowed
    torch.set_num_threads(max(1, min(4, len(share))))
    return {"numa_node": node, "cores": [share[0], share[-1]], "n_cores": len(share)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=16384000 / 2**20)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--bind", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    binding = bind_rank_to_gpu_node(local_rank, world) if args.bind else {"numa_node": numa_node_of_gpu(local_rank), "cores": None}
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = int(args.mb * 2**20)
    # the copies are issued by the library itself (cudaHostAlloc + cudaMemcpyAsync, exactly what sag_step_host does)
    import ctypes as C
    from safe_adaptation_gym_b200 import _abi
    L = _abi.load()
    cfg = L.default_config()
    cfg.n_envs = 65536
    h = C.c_void_p()
    L.check(L.L.sag_create(C.byref(cfg), local_rank, C.byref(h)))
    secs = C.c_double()
    L.check(L.L.sag_probe_d2h(h, n, 5, C.byref(secs)))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    L.check(L.L.sag_probe_d2h(h, n, args.reps, C.byref(secs)))
    dt = secs.value
    L.L.sag_destroy(h)
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        infos = [None] * world
        dist.all_gather_object(infos, {"rank": rank, "gbs": n * args.reps / dt / 1e9, **binding})
    else:
        infos = [{"rank": 0, "gbs": n * args.reps / dt / 1e9, **binding}]
    if rank == 0:
        agg = world * n * args.reps / float(t[0]) / 1e9
        print(json.dumps({"probe": "d2h_pinned", "n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps, "bind": args.bind,
                          "aggregate_gbs": agg, "per_gpu_gbs": agg / world, "host_cores": os.cpu_count(), "ranks": infos}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
