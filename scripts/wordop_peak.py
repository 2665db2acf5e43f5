#!/usr/bin/env python
"""Measured integer word-op throughput of this GPU: AND + popcount + accumulate of 64-bit words, the unit SURVEY 8d
quotes the tolerant mode's work in (2 * (2n - 2) * W word-ops per tree pair).  bench.py reads profiles/wordop_peak.json
for cfg5's roofline.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from phybin_b200.rfdistance import RFContext  # noqa: E402

ctx = RFContext(0)
best = 0.0
for _ in range(5):
    v = C.c_double(0)
    ctx._ck(ctx._L.prf_debug_wordop_peak(ctx._h, C.byref(v)))
    best = max(best, v.value)
props = torch.cuda.get_device_properties(0)
print(json.dumps({"device": props.name, "sms": props.multi_processor_count, "unit": "word-ops/s (64-bit AND + popc + add)",
                  "wordops_per_s": best, "wordops_per_clk_per_sm_at_1965MHz": best / props.multi_processor_count / 1.965e9}))
ctx.close()
