#!/usr/bin/env python
"""Measured dense int8 tensor-core throughput of this GPU (the denominator of the dense variant's roofline; the
driver-written MEASURED_PEAKS.json only has bf16): cuBLASLt int8 x int8 -> int32 GEMMs through torch._int_mm, CUDA-event
timed after warm-up.  Prints one JSON line; bench.py reads profiles/int8_peak.json when it reports the dense variant."""
import json
import sys

import torch


def main():
    dev = torch.device("cuda", 0)
    res = {"device": torch.cuda.get_device_name(0), "unit": "Top/s", "runs": []}
    for n in (4096, 8192, 16384):
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        reps = 20 if n <= 8192 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res["runs"].append({"n": n, "ms": ms, "tops": 2.0 * n ** 3 / (ms * 1e-3) / 1e12})
        del a, b
    res["int8_tops"] = max(r["tops"] for r in res["runs"])
    print(json.dumps(res))


if __name__ == "__main__":
    sys.exit(main())
