#!/usr/bin/env python3
"""Host-side copy bandwidth probe for the end-to-end path: what the box can move between HBM and pinned host memory.

    python tools/host_bw_probe.py                  one process, GPU 0
    torchrun --nproc-per-node N tools/host_bw_probe.py     N ranks copying at the same time (aggregate = what an N-GPU e2e run sees)

Prints one JSON line per rank 0: per-rank and aggregate D2H / H2D GB/s for default pinned memory and, when the box has
several NUMA nodes, for pinned memory bound to the node of the GPU (mbind through libnuma if present)."""
import ctypes
import json
import os
import subprocess
import time

import torch


def topo():
    out = {}
    for name, cmd in (("lscpu_numa", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"), ("nvidia_topo", "nvidia-smi topo -m"),
                      ("nodes", "ls /sys/devices/system/node/ | grep node"), ("affinity", "taskset -p $$"), ("mem", "free -g | head -2")):
        try:
            out[name] = subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
        except Exception as e:  # noqa: BLE001
            out[name] = f"failed: {e}"
    return out


def bw(dst, src, stream, reps=3):
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        dst.copy_(src, non_blocking=True)
        e1.record(stream)
        e1.synchronize()
        best = max(best, src.numel() * src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = (3 << 30) // 8  # 3 GiB, the per-GPU result volume of the 8-GPU C2 run
    d = torch.empty(n, dtype=torch.float64, device=dev).normal_()
    stream = torch.cuda.current_stream(dev)
    res = {}
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.zero_()
    for label, a, b in (("d2h", h, d), ("h2d", d, h)):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        v = bw(a, b, stream)
        if dist is not None:
            dist.barrier()
        wall = time.perf_counter() - t
        vals = torch.tensor([v], device=dev, dtype=torch.float64)
        if dist is not None:
            allv = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(allv, vals)
            per = [float(x.item()) for x in allv]
        else:
            per = [v]
        res[label] = {"per_rank_gbs": per, "sum_of_best_gbs": sum(per), "wall_s_for_3_reps": wall,
                      "aggregate_gbs_by_wall": world * 3 * n * 8 / wall / 1e9}
    # D2H in 78 MB pieces (the size of a streamed batch of the HOST pipeline) on 1, 2 and 4 streams at once
    for n_streams in (1, 2, 4):
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        piece = 78_643_200 // 8
        pieces = [(o, min(o + piece, n)) for o in range(0, n, piece)]
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i, (a, b) in enumerate(pieces):
            with torch.cuda.stream(streams[i % n_streams]):
                h[a:b].copy_(d[a:b], non_blocking=True)
        torch.cuda.synchronize()
        mine = n * 8 / (time.perf_counter() - t) / 1e9
        if dist is not None:
            dist.barrier()
        wall = time.perf_counter() - t
        vals = torch.tensor([mine], device=dev, dtype=torch.float64)
        if dist is not None:
            allv = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(allv, vals)
            per = [float(x.item()) for x in allv]
        else:
            per = [mine]
        res["d2h_78MB_pieces_%d_streams" % n_streams] = {"per_rank_gbs": per, "aggregate_gbs_by_wall": world * n * 8 / wall / 1e9}
    # Does the page size of the pinned buffer matter (IOMMU translations of a virtualised box)?  The same piecewise copy into an
    # anonymous mapping with transparent huge pages requested, registered with cudaHostRegister.
    try:
        import mmap
        size = n * 8
        mm = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        mm.madvise(mmap.MADV_HUGEPAGE)
        base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        off = (-base) % (2 << 20)
        hh = torch.frombuffer(mm, dtype=torch.float64, count=n, offset=off)
        hh.zero_()
        rc = torch.cuda.cudart().cudaHostRegister(base + off, size, 0)
        thp = {"registered_rc": int(rc), "is_pinned": bool(hh.is_pinned())}
        try:
            thp["AnonHugePages_kB_of_process"] = int([l for l in open("/proc/self/smaps_rollup") if l.startswith("AnonHugePages")][0].split()[1])
            thp["thp_enabled"] = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        except Exception as e:  # noqa: BLE001
            thp["thp_query"] = repr(e)
        piece = 78_643_200 // 8
        pieces = [(o, min(o + piece, n)) for o in range(0, n, piece)]
        for label, buf in (("torch_pinned", h), ("thp_registered", hh)):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for a, b in pieces:
                buf[a:b].copy_(d[a:b], non_blocking=True)
            torch.cuda.synchronize()
            mine = n * 8 / (time.perf_counter() - t) / 1e9
            vals = torch.tensor([mine], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.barrier()
                allv = [torch.zeros_like(vals) for _ in range(world)]
                dist.all_gather(allv, vals)
                thp[label + "_per_rank_gbs"] = [round(float(x.item()), 2) for x in allv]
            else:
                thp[label + "_per_rank_gbs"] = [round(mine, 2)]
        res["d2h_page_size"] = thp
    except Exception as e:  # noqa: BLE001
        res["d2h_page_size"] = {"failed": repr(e)}
    # who shares what: only a SUBSET of the ranks copies (78 MB pieces, one stream), the others wait at the barrier
    if dist is not None and world >= 2:
        half = world // 2
        subsets = [[0], [world - 1], list(range(half)), list(range(half, world)), [0, world - 1], list(range(world))]
        piece = 78_643_200 // 8
        pieces = [(o, min(o + piece, n)) for o in range(0, n, piece)]
        res["d2h_subsets_78MB_pieces"] = []
        for sub in subsets:
            dist.barrier()
            torch.cuda.synchronize()
            mine = 0.0
            if rank in sub:
                t = time.perf_counter()
                for a, b in pieces:
                    h[a:b].copy_(d[a:b], non_blocking=True)
                torch.cuda.synchronize()
                mine = n * 8 / (time.perf_counter() - t) / 1e9
            dist.barrier()
            vals = torch.tensor([mine], device=dev, dtype=torch.float64)
            allv = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(allv, vals)
            res["d2h_subsets_78MB_pieces"].append({"ranks_copying": sub, "per_rank_gbs": [round(float(x.item()), 2) for x in allv]})
    if rank == 0:
        print(json.dumps({"world": world, "bytes_per_rank": n * 8, "results": res, "topology": topo()}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
