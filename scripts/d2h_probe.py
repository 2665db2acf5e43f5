#!/usr/bin/env python
"""Device -> pinned-host copy bandwidth of N ranks at once (what bounds bench.py's e2e at N > 1: every rank ships its
slice of the 10 GB matrix over its own PCIe link into host memory).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/d2h_probe.py [--mb 1250]

For every variant (pinned memory as torch allocates it / allocated after pinning the rank to the CPUs nvidia-smi lists
as local to its GPU / write-combined) it times (a) each rank alone, the others idle, and (b) all ranks together, and
rank 0 prints one JSON line: per-rank GB/s alone, per-rank GB/s together, aggregate GB/s together."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def gpu_cpu_affinity(index: int):
    """CPUs local to GPU `index` according to `nvidia-smi topo -m` (None if it cannot be parsed)."""
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    for ln in txt.splitlines():
        f = ln.split()
        if f and f[0] == f"GPU{index}":
            for tok in f[1:]:
                if tok[0].isdigit() and ("-" in tok or "," in tok) and not tok.endswith("X"):
                    cpus = set()
                    try:
                        for part in tok.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                    except ValueError:
                        continue
                    return cpus
    return None


def main():
    mb = int(sys.argv[sys.argv.index("--mb") + 1]) if "--mb" in sys.argv else 1250
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = mb * 1000 * 1000 // 2
    src = torch.empty(n, dtype=torch.int16, device=dev).fill_(7)
    res = {"world": world, "mb_per_rank": mb, "host_cpus": os.cpu_count(), "variants": {}}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(dst, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
        return best

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [x]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o[0]) for o in out]

    aff0 = os.sched_getaffinity(0)
    for variant in ("torch_pinned", "pinned_after_gpu_local_affinity"):
        note = ""
        if variant == "pinned_after_gpu_local_affinity":
            cpus = gpu_cpu_affinity(local)
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus & aff0 or aff0)
                    note = f"rank {rank}: {len(cpus & aff0)} local cpus"
                except OSError as e:
                    note = f"setaffinity failed: {e}"
            else:
                note = "no affinity information"
        dst = torch.empty(n, dtype=torch.int16).pin_memory()
        dst.zero_()            # first touch on the (possibly re-pinned) thread
        timed(dst, 1)          # warm-up
        alone = []
        for r in range(world):
            barrier()
            t = timed(dst) if r == rank else 0.0
            barrier()
            alone.append(t)
        barrier()
        together = timed(dst)
        barrier()
        a = gather(alone[rank])
        tg = gather(together)
        res["variants"][variant] = {
            "alone_gbs": [round(mb / 1e3 / x, 2) for x in a],
            "together_gbs": [round(mb / 1e3 / x, 2) for x in tg],
            "aggregate_together_gbs": round(world * mb / 1e3 / max(tg), 2),
            "note": note,
        }
        del dst
    os.sched_setaffinity(0, aff0)
    if rank == 0:
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:4000]
        except Exception:
            pass
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
