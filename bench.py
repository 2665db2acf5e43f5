#!/usr/bin/env python
"""bench.py -- RF tree-pairs/s on the headline workload (BASELINE.json: 100k trees x 200 taxa).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3|cfg2|...]

One "step" = one pass of the hot path over the whole batch: canonical bipartitions resident in HBM
-> dedup -> incidence -> matrix kernel -> packed uint16 upper triangle resident in HBM.  With N > 1
(torchrun, one rank per GPU) the trees are sharded across ranks, the build is hash-partitioned over the
GPUs (prf_team_load: the library's own kernels store keys and list structures into the peers' memory
over NVLink; torch.distributed only carries the 64-byte arena handles once, outside the timed region),
and each rank computes a contiguous, aligned slice of the packed triangle (fixed total work => "strong"
scaling).  Rank 0 prints ONE JSON line.

`e2e` is the same metric through the public host-buffer call prf_hashrf (pinned host input, H2D,
compute, D2H of the whole matrix into pinned host memory).  `cpu_baseline` / `--impl reference` time
the oracle's literal restatement of the reference algorithm (the Haskell binary cannot be built in
this image: no GHC) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (T, n, seed, family, editdist, mode)
    "cfg2": dict(T=10_000, n=100, seed=0x5EED0002, family=0, editdist=-1, mode="hashrf"),
    "cfg3": dict(T=100_000, n=200, seed=0x5EED0003, family=0, editdist=-1, mode="hashrf"),
    # the "dense-universe" family of SURVEY 8d: one base tree + r ~ U{0..8} NNI moves per tree (tensor-core variant)
    "cfg2_nni": dict(T=10_000, n=100, seed=0x5EED0002, family=1, editdist=-1, mode="hashrf"),
    "cfg3_nni": dict(T=100_000, n=200, seed=0x5EED0003, family=1, editdist=-1, mode="hashrf"),
    "cfg4": dict(T=50_000, n=500, seed=0x5EED0004, family=0, editdist=4, mode="hashrf"),
    # cfg4's shape on the similar-trees family: the mask's consumer (components + UPGMA flat clusters) is non-trivial here
    "cfg4_nni": dict(T=50_000, n=500, seed=0x5EED0004, family=1, editdist=4, mode="hashrf"),
    "cfg5": dict(T=20_000, n=150, seed=0x5EED0005, family=0, editdist=-1, mode="tolerant", p_missing=0.10),
}
METRIC = "rf_tree_pairs_per_sec"
UNIT = "tree-pairs/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Samples taken before this call are ignored by stop()."""
        self.first = len(self.lines)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle's literal hashRF / naiveDistMatrix on a bounded sample
# ---------------------------------------------------------------------------------------------------

def cpu_sample(wl: dict, budget_trees: int, all_cores: bool = False):
    """Times the oracle (literal restatement of RFDistance.hs:284-341 / :168-198) on the first `budget_trees` trees of
    the workload: on one thread, like the reference's sequential runST ingest (RFDistance.hs:300-341), or -- hashRF
    only -- the same algorithm parallelised over all host cores (oracle: hashRFAllCores).
    Returns (pairs_per_s, seconds, threads, sample description)."""
    from oracle import oracle as orc
    from phybin_b200.host import Forest
    Ts = min(budget_trees, wl["T"])
    f = Forest.synthetic(Ts, wl["n"], wl["seed"], wl["family"], 8, wl.get("p_missing", 0.0))
    text = f.to_newick()
    of = orc.Forest([text])
    threads = 1
    t0 = time.perf_counter()
    if wl["mode"] == "tolerant":
        of.naive()
    elif all_cores:
        of.hashrf(wl["n"], "literal_all_cores")
        threads = orc.max_threads()
    else:
        of.hashrf(wl["n"], "literal")
    dt = time.perf_counter() - t0
    pairs = Ts * (Ts - 1) // 2
    how = f"literal {wl['mode']} restatement (allBips + table + matrix), " + (f"OpenMP on {threads} threads" if all_cores and wl["mode"] != "tolerant" else "1 thread like the reference")
    return pairs / dt, dt, threads, f"first {Ts} of {wl['T']} trees x {wl['n']} taxa, all {pairs} pairs, {how}"


def cpu_baseline_report(wl: dict, ref_trees: int, full_sweep: bool) -> dict:
    """cpu_baseline of the GPU arm's line: the all-cores variant at the largest sampled T is the
    value; the single-threaded run (what the reference's runtime does) and the T sweep are listed beside it."""
    if wl["mode"] == "tolerant":
        v, s, th, desc = cpu_sample(wl, ref_trees or 150)
        return {"value": v, "unit": UNIT, "cores": th, "kind": "port", "sample": desc, "seconds": s, "host_cores": os.cpu_count()}
    sweep = []
    sizes = [1000, 2000, 5000] + ([10000] if full_sweep else [])
    if ref_trees:
        sizes = [ref_trees]
    for Ts in sizes:
        v, s, th, desc = cpu_sample(wl, Ts, all_cores=True)
        sweep.append({"trees": min(Ts, wl["T"]), "cores": th, "value": v, "seconds": s})
    v1, s1, _, desc1 = cpu_sample(wl, min(sizes[-1], 2000))
    return {"value": v, "unit": UNIT, "cores": th, "kind": "port", "sample": desc, "seconds": s, "host_cores": os.cpu_count(),
            "single_thread": {"value": v1, "unit": UNIT, "cores": 1, "sample": desc1, "seconds": s1},
            "sweep": sweep}


def workload_config(wl: dict, wl_name: str, editdist: int, nnz: int, words: int) -> dict:
    """The `config` object: the same for both arms (what was sampled or measured is reported elsewhere in the line)."""
    T = wl["T"]
    return {"workload": wl_name, "trees": T, "taxa": wl["n"], "mode": wl["mode"], "editdist": editdist,
            "pairs": T * (T - 1) // 2, "nnz": int(nnz), "words": int(words)}


def run_reference(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from phybin_b200.host import Forest
    hashrf = wl["mode"] == "hashrf"
    sample_T = args.ref_trees or (4000 if hashrf else 150)
    f = Forest.synthetic(wl["T"], wl["n"], wl["seed"], wl["family"], 8, wl.get("p_missing", 0.0))
    if hashrf:
        _, off = f.all_bips()
    else:
        _, off, _ = f.raw_clades()
    cfg = workload_config(wl, wl_name, wl["editdist"], int(off[-1]), f.words)
    del f, off
    for _ in range(args.warmup):
        cpu_sample(wl, max(100, sample_T // 8), all_cores=True)
    vals, secs, desc, th = [], [], "", 1
    for _ in range(args.steps):
        v, s, th, desc = cpu_sample(wl, sample_T, all_cores=True)
        vals.append(v); secs.append(s)
    total_pairs = len(vals) * (min(sample_T, wl["T"]) * (min(sample_T, wl["T"]) - 1) // 2)
    value = total_pairs / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "config": cfg,
        "sampled_trees": min(sample_T, wl["T"]),
        "note": "reference arm = CPU oracle (C++ restatement of PhyBin's hashRF; GHC is absent so the Haskell binary cannot be built). "
                "Each step runs the whole path on the first `sampled_trees` trees of the workload and the rate is in tree-pairs/s "
                "(the reference's cost per pair does not fall with T). The reference's own ingest is single-threaded "
                "(runST, RFDistance.hs:300-341); this arm uses every host core (cpu_baseline.cores).",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": desc,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

def run_gpu(args, wl, wl_name):
    import numpy as np
    import torch
    import torch.distributed as dist
    from phybin_b200.host import Forest
    from phybin_b200.rfdistance import MEM_DEVICE, RFContext, shard_range
    from phybin_b200.distributed import KeyGather, shard_rows as key_shard_rows, team_load, team_setup, tree_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; phybin_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    T, n, mode, editdist = wl["T"], wl["n"], wl["mode"], wl["editdist"]
    P = T * (T - 1) // 2
    ctx = RFContext(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- synthetic input: every rank generates only its own shard of trees ----------------------
    t_lo, t_hi = tree_shard(T, rank, world)
    f = Forest.synthetic(T, n, wl["seed"], wl["family"], 8, wl.get("p_missing", 0.0))
    W = f.words
    if mode == "hashrf":
        bips_all, off_all = f.all_bips()
        b0, b1 = int(off_all[t_lo]), int(off_all[t_hi])
        my_bips = np.ascontiguousarray(bips_all[b0:b1])
        counts = np.diff(off_all.astype(np.int64))
    else:
        clades_all, off_all, leaf_all = f.raw_clades()
        b0, b1 = int(off_all[t_lo]), int(off_all[t_hi])
        my_bips = np.ascontiguousarray(clades_all[b0:b1])
        counts = np.diff(off_all.astype(np.int64))
    nnz = int(off_all[-1])
    # the parsed trees as flat arrays (prf_allbips input) for the from-trees end-to-end number, N = 1 only
    tree_arrays = [a.copy() for a in f.tree_arrays()] if (world == 1 and not args.no_e2e) else None
    del f

    with torch.cuda.stream(ext):
        d_my = torch.from_numpy(my_bips.view(np.int64)).to(dev)                  # this rank's tree shard, resident
        d_off = torch.from_numpy(off_all.view(np.int64)).to(dev)                  # offsets are tiny and replicated
        d_leaf = torch.from_numpy(leaf_all.view(np.int64)).to(dev) if mode == "tolerant" else None
        gather = KeyGather(key_shard_rows(off_all, T, world), W, dev)
        d_off_my = torch.from_numpy((off_all[t_lo:t_hi + 1] - off_all[t_lo]).astype(np.int64)).to(dev)
    # N > 1, hashRF: hash-partitioned build over the ranks' GPUs (prf_team.cuh); --build replicated = NCCL all-gather
    # of the keys + the whole build on every rank (kept for comparison)
    build = args.build if args.build != "auto" else "sharded"
    sharded = world > 1 and mode == "hashrf" and build == "sharded"
    if sharded:
        rows_per_rank = key_shard_rows(off_all, T, world)
        team_setup(ctx, T, W, nnz, max(rows_per_rank), int(counts.max()) if len(counts) else 0)
    with torch.cuda.stream(ext):
        p_begin, p_end = shard_range(T, rank, world)
        d_out = torch.empty(max(p_end - p_begin, 8), dtype=torch.int16, device=dev)
        d_mask = torch.empty(max((p_end - p_begin + 31) // 32, 1), dtype=torch.int32, device=dev) if editdist >= 0 else None
    ext.synchronize()

    matrix_ms = []

    def step():
        """dedup + incidence + matrix for this rank's slice; inputs and outputs stay in HBM."""
        if sharded:
            # hash-partitioned build: all-to-all of keys, per-class dedup, all-gather + merge of the structures
            team_load(ctx, d_my, d_off_my, T)
        else:
            with torch.cuda.stream(ext):
                d_full = gather(d_my)     # NCCL all-gather over NVLink when world > 1
        if sharded:
            st = ctx.hashrf_compute(p_begin, p_end, editdist, 0, d_out.data_ptr(),
                                    d_mask.data_ptr() if d_mask is not None else 0)
        elif mode == "hashrf":
            ctx.hashrf_load(d_full.data_ptr(), d_off.data_ptr(), T, W, MEM_DEVICE)
            st = ctx.hashrf_compute(p_begin, p_end, editdist, 0, d_out.data_ptr(),
                                    d_mask.data_ptr() if d_mask is not None else 0)
        else:
            ctx.tolerant_load(d_full.data_ptr(), d_off.data_ptr(), d_leaf.data_ptr(), T, W, MEM_DEVICE)
            st = ctx.tolerant_compute(p_begin, p_end, editdist, d_out.data_ptr(),
                                      d_mask.data_ptr() if d_mask is not None else 0)
        matrix_ms.append(st["ms_matrix"])
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation takes driver locks) can stall
    # kernel launches for milliseconds and must not fall into the timed region; only samples taken after
    # sampler.mark() below are used.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(2):   # setup, not warm-up: lets the stream-ordered memory pool reach its steady size
        step()
    barrier()
    for _ in range(args.warmup):
        st = step()
    barrier()
    matrix_ms.clear()
    launches0 = ctx.launches
    if sampler:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = step()
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    launches = ctx.launches - launches0
    tm = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_s, wall = float(tm[0]), float(tm[1])
    # the timed region can be shorter than nvidia-smi's sampling period: keep the GPU under the same load
    # (identical, untimed steps) until the sampler has seen it for >= 1.5 s.  The count comes from the
    # all-reduced time, so every rank runs the same number of (collective) steps.
    probe_steps = 0
    if wall < 1.5:
        probe_steps = int(min(2000, max(1, (1.5 - wall) / max(wall / args.steps, 1e-4))))
        for _ in range(probe_steps):
            step()
        barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["sampled_over"] = f"timed region + {probe_steps} identical untimed steps (sampler period 200 ms)"
    # the step synchronises with the host a few times (scalar read-backs), so device-event time on the
    # library's stream and wall time between the bracketing synchronisations agree; report the larger.
    step_s = max(dev_s, wall) / args.steps
    value = P / step_s

    # ---- self-check on the resident result (not timed) --------------------------------------------------
    # (1) 256 pairs sampled over this rank's whole slice against a direct set symmetric difference;
    # (2) the checksum identity  sum d = (T-1) * sum_t |B_t| - 2 * sum_lists m(m-1)/2  at N = 1 (when no
    #     majority flip re-centred the sets: n_hits then counts exactly the shared incidences).
    selfcheck = None
    if mode == "hashrf" and not os.environ.get("PRF_DEBUG_NOWALK"):
        rng = np.random.default_rng(1 + rank)
        n_slice = p_end - p_begin
        pos = np.sort(rng.integers(0, n_slice, size=256)) if n_slice else np.zeros(0, dtype=np.int64)
        vals = d_out[torch.from_numpy(pos).to(dev)].cpu().numpy().view(np.uint16) if n_slice else []
        row_starts = np.array([r * T - r * (r + 1) // 2 for r in range(T)], dtype=np.int64)
        row_first = int(np.searchsorted(row_starts, p_begin, side="right") - 1)
        row_last = min(T - 2, int(np.searchsorted(row_starts, max(p_end - 1, p_begin), side="right") - 1))
        for p, v in zip(pos, vals):
            gp = p_begin + int(p)
            a = int(np.searchsorted(row_starts, gp, side="right") - 1)
            b = a + 1 + gp - int(row_starts[a])
            sa = {tuple(x) for x in bips_all[int(off_all[a]):int(off_all[a + 1])].tolist()}
            sb = {tuple(x) for x in bips_all[int(off_all[b]):int(off_all[b + 1])].tolist()}
            assert int(v) == len(sa ^ sb), f"self-check failed at pair ({a},{b}): {int(v)} != {len(sa ^ sb)}"
        selfcheck = {"sampled_pairs": int(len(pos))}
        if not args.quick_check:
            # independent of the library: the multiset of 32-byte keys -> Map bip -> trees on the host (numpy)
            t_chk = time.perf_counter()
            void = np.ascontiguousarray(bips_all).view(np.dtype((np.void, bips_all.shape[1] * 8))).ravel()
            _, inv, mult = np.unique(void, return_inverse=True, return_counts=True)
            inv = inv.ravel()
            tree_of = np.repeat(np.arange(T, dtype=np.int64), counts)
            # a tree's entries are a set: count every (key, tree) once
            pair = np.unique(inv.astype(np.int64) * T + tree_of)
            m = np.bincount(pair // T, minlength=len(mult)).astype(np.int64)
            host_hits = int((m * (m - 1) // 2).sum())
            nb_host = np.bincount(pair % T, minlength=T).astype(np.int64)
            selfcheck["host_n_unique"] = int(len(mult))
            selfcheck["host_n_hits"] = host_hits
            assert int(st["n_unique"]) == len(mult), f"n_unique {st['n_unique']} != host {len(mult)}"
            if not st.get("n_flipped"):
                assert int(st["n_hits"]) == host_hits, f"n_hits {st['n_hits']} != host {host_hits}"
            # 8 full rows of this rank's slice recomputed on the host from the key ids
            key_of, tr_of = pair // T, pair % T
            rows_chk = []
            for a in sorted(set(int(x) for x in rng.integers(row_first, row_last + 1, size=8))):
                mark = np.zeros(len(mult), dtype=bool)
                mark[key_of[tr_of == a]] = True
                shared = np.bincount(tr_of[mark[key_of]], minlength=T).astype(np.int64)
                want = nb_host[a] + nb_host[a + 1:] - 2 * shared[a + 1:]
                g0 = int(row_starts[a]); lo = max(g0, p_begin); hi = min(g0 + (T - 1 - a), p_end)
                if hi <= lo:
                    continue
                got = d_out[lo - p_begin:hi - p_begin].cpu().numpy().view(np.uint16).astype(np.int64)
                assert np.array_equal(got, want[lo - g0:hi - g0]), f"row {a} differs from the host recomputation"
                rows_chk.append(a)
            selfcheck["host_rows_checked"] = rows_chk
            selfcheck["host_check_s"] = round(time.perf_counter() - t_chk, 1)
        if world == 1 and not st.get("n_flipped"):
            total = int(d_out[:P].view(torch.int16).to(torch.int64).bitwise_and(0xFFFF).sum().item()) if P else 0
            hits_ref = selfcheck.get("host_n_hits", int(st["n_hits"]))
            want = (T - 1) * int(counts.sum()) - 2 * hits_ref
            assert total == want, f"checksum identity failed: {total} != {want}"
            selfcheck["checksum"] = total
            selfcheck["checksum_uses"] = "host n_hits" if "host_n_hits" in selfcheck else "library n_hits"
        if world > 1:
            # every rank's slice against the same slice of a single-GPU compute (rank 0 redoes the whole job alone)
            def slice_sums(t):
                v = t.view(torch.int16).to(torch.int64).bitwise_and(0xFFFF)
                w = (torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521) + 1
                return [int(v.sum().item()), int((v * w).sum().item() & 0x7FFFFFFFFFFFFFFF)]
            mine = slice_sums(d_out[:p_end - p_begin]) + [p_begin, p_end]
            allsums = [None] * world
            dist.all_gather_object(allsums, mine)
            if rank == 0:
                with torch.cuda.stream(ext):
                    d_all = torch.from_numpy(bips_all.view(np.int64)).to(dev)
                ext.synchronize()
                ctx1 = RFContext(local_rank)
                ext1 = torch.cuda.ExternalStream(ctx1.stream, device=dev)
                ctx1.hashrf_load(d_all.data_ptr(), d_off.data_ptr(), T, W, MEM_DEVICE)
                ok = []
                for r, (s0, s1, pb, pe) in enumerate(allsums):
                    with torch.cuda.stream(ext1):
                        buf = torch.empty(max(pe - pb, 8), dtype=torch.int16, device=dev)
                    ctx1.hashrf_compute(pb, pe, -1, 1, buf.data_ptr(), 0)
                    ref = slice_sums(buf[:pe - pb])
                    assert ref == [s0, s1], f"slice of rank {r} differs from the single-GPU compute"
                    ok.append(r)
                    del buf
                ctx1.close()
                del d_all
                selfcheck["slices_equal_single_gpu"] = ok

    # ---- what the mask feeds (untimed, N = 1): flat clusters at --editdist, as doCluster / sliceDendro use it ------
    consumer = None
    if world == 1 and editdist >= 0 and d_mask is not None and P:
        t0 = time.perf_counter()
        _, ncomp = ctx.mask_components(d_mask.data_ptr())
        t1 = time.perf_counter()
        _, nclu = ctx.flat_clusters(editdist, "UPGMA", d_mask.data_ptr(), d_out.data_ptr())
        t2 = time.perf_counter()
        lab_c, _ = ctx.mask_components(d_mask.data_ptr())
        consumer = {"what": "connected components of {d <= editdist} on the device (prf_mask_components), then UPGMA flat "
                            "clusters (sliceDendro editdist . C.dendrogram UPGMA): components of >= 64 trees agglomerated on the "
                            "device (prf_agglomerate), smaller ones by the host restatement of the same rule",
                    "components_ms": (t1 - t0) * 1e3, "n_components": int(ncomp),
                    "largest_component": int(np.bincount(lab_c).max()) if len(lab_c) else 0,
                    "upgma_flat_clusters_ms": (t2 - t1) * 1e3, "n_clusters": int(nclu)}

    # ---- roofline of the dominant kernel (the matrix kernel), per launch ---------------------------
    peaks, peak_src = load_peaks()
    my_pairs = p_end - p_begin
    alg_bytes = 2 * my_pairs + (my_pairs // 8 if editdist >= 0 else 0)
    timed_ms = matrix_ms[:args.steps]
    avg_ms = sum(timed_ms) / max(len(timed_ms), 1)
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                "kernel": "rows_kernel" if mode == "hashrf" else "tolerant_kernel",
                "kernel_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_share_of_step": avg_ms * 1e-3 / step_s}
    if mode == "tolerant":
        # HBM is not what binds this kernel: per pair it restricts, re-normalises and hashes both trees' clades --
        # 2 * (2n - 2) * W word-ops (64-bit AND + popcount) by the survey's count.  Reported against the MEASURED word-op peak.
        n_taxa = wl["n"]
        wordops = 2.0 * (2 * n_taxa - 2) * W * my_pairs
        got = wordops / (avg_ms * 1e-3) if avg_ms > 0 else 0.0
        peak_w, w_src = 8.0 * 148 * 1.965e9, "nominal: 16 POPC/clk/SM -> 8 word-ops/clk/SM x 148 SMs x 1965 MHz (profiles/wordop_peak.json missing)"
        try:
            with open(os.path.join(ROOT, "profiles", "wordop_peak.json")) as fh:
                peak_w, w_src = float(json.load(fh)["wordops_per_s"]), "measured: scripts/wordop_peak.py on a B200 of this pool (profiles/wordop_peak.json)"
        except Exception:
            pass
        roofline = {"bound": "alu", "achieved": got / 1e9, "peak": peak_w / 1e9, "unit": "Gword-op/s", "frac": got / peak_w,
                    "traffic": None, "peak_source": w_src, "kernel": "tolerant_kernel", "kernel_ms": avg_ms,
                    "algorithmic_wordops_per_launch": wordops, "wordops_per_pair": 2.0 * (2 * n_taxa - 2) * W,
                    "kernel_share_of_step": avg_ms * 1e-3 / step_s,
                    "hbm": {"achieved_gbs": achieved, "frac": achieved / peaks["hbm_gbs"], "note": "2 B per pair out; not the binding limit"}}
    if mode == "hashrf" and st.get("variant_used") == 2:
        # tensor-core variant: 2 * Kpad int8 ops per pair (Kpad = shared bipartitions padded to the 128-column k-block)
        kpad = (int(st["n_shared"]) + 127) // 128 * 128
        ops = 2.0 * kpad * my_pairs
        tops = ops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        peak_i8, i8_src = 2.0 * peaks["bf16_tflops"], peak_src + " bf16 x2 (nominal ratio; profiles/int8_peak.json missing)"
        try:
            with open(os.path.join(ROOT, "profiles", "int8_peak.json")) as fh:
                peak_i8, i8_src = float(json.load(fh)["int8_tops"]), "measured: cuBLASLt int8 GEMM on a B200 of this pool (scripts/int8_peak.py, profiles/int8_peak.json)"
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": tops, "peak": peak_i8, "unit": "TFLOP/s", "frac": tops / peak_i8,
                    "traffic": None, "peak_source": i8_src,
                    "kernel": "dense_gram_kernel", "kernel_ms": avg_ms, "algorithmic_ops_per_launch": ops,
                    "k_columns": kpad, "output_gbs": alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0,
                    "kernel_share_of_step": avg_ms * 1e-3 / step_s}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                # measured for the sparse matrix kernel at N = 1 (per-rank captures at N > 1 are not taken: ncu must not
                # wrap a multi-rank job)
                roofline["traffic"] = json.load(fh).get(wl_name if world == 1 and roofline["bound"] == "hbm" and mode == "hashrf" else "", None)
        except Exception:
            pass

    # ---- e2e through the host-buffer C ABI call (rank-local slice of work: whole job at N=1) ------
    e2e = None
    if world == 1 and not args.no_e2e:
        e2e = run_e2e(ctx, wl, mode, bips_all if mode == "hashrf" else clades_all, off_all,
                      None if mode == "hashrf" else leaf_all, P, editdist, max(1, min(args.steps, 2)), tree_arrays)
    elif world > 1 and not args.no_e2e:
        # every rank: its tree shard from pinned host memory -> HBM, all-gather, build, its slice of the matrix,
        # slice -> pinned host memory.  Each GPU uses its own PCIe link, so the copies run in parallel.
        pin_in = torch.from_numpy(my_bips.view(np.int64)).pin_memory()
        pin_out = torch.empty(max(p_end - p_begin, 8), dtype=torch.int16).pin_memory()
        pin_mask = torch.empty(max((p_end - p_begin + 31) // 32, 1), dtype=torch.int32).pin_memory() if editdist >= 0 else None
        n_e2e = max(1, min(args.steps, 2))

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def e2e_step():
            with torch.cuda.stream(ext):
                ev[0].record(ext)
                d_my.copy_(pin_in, non_blocking=True)
                ev[1].record(ext)
            step()
            with torch.cuda.stream(ext):
                ev[2].record(ext)
                pin_out[: p_end - p_begin].copy_(d_out[: p_end - p_begin], non_blocking=True)
                if pin_mask is not None:
                    pin_mask.copy_(d_mask, non_blocking=True)
                ev[3].record(ext)
            ext.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nbytes = torch.tensor([float(my_bips.nbytes), float((p_end - p_begin) * 2 + (pin_mask.numel() * 4 if pin_mask is not None else 0))],
                              dtype=torch.float64, device=dev)
        dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
        del pin_in, pin_out, pin_mask  # before the library's stream goes away (the host allocator records events on it)
        e2e = {"value": P / float(dt[0]), "unit": UNIT, "h2d_bytes_per_step": int(nbytes[0]), "d2h_bytes_per_step": int(nbytes[1]),
               "ms_per_step": float(dt[0]) * 1e3, "steps": n_e2e,
               "phases_ms": {"h2d": ev[0].elapsed_time(ev[1]), "build_and_matrix": ev[1].elapsed_time(ev[2]), "d2h": ev[2].elapsed_time(ev[3]),
                             "note": "rank 0, last step (CUDA events on the library's stream)"},
               "d2h_aggregate_gbs": float(nbytes[1]) / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9,
               "host_ceiling": "the step is the device -> pinned-host copy of the matrix: d2h_aggregate_gbs = all ranks' slice bytes / rank 0's "
                               "copy time in this run; scripts/d2h_probe.py measures the same ceiling without the compute (one rank alone: "
                               "~57 GB/s, one PCIe Gen5 x16 link)",
               "note": "per rank: pinned tree shard -> HBM, hash-partitioned build over NVLink, own slice, slice -> pinned host; bytes summed over ranks"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = cpu_baseline_report(wl, args.ref_trees, args.cpu_sweep_full)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16" if mode == "hashrf" else "u64",
            "data": "synthetic",
            "config": workload_config(wl, wl_name, editdist, nnz, W),
            "details": {"n_unique": st.get("n_unique"), "n_hits": st.get("n_hits"),
                       "variant": {1: "sparse", 2: "dense"}.get(st.get("variant_used"), mode), "n_shared": st.get("n_shared"),
                       "parallelism": f"packed-range shard x{world}" + ((" + hash-partitioned build (prf_team_load: keys and per-class list structures stored into the peers' memory over NVLink by the library's kernels)" if sharded else " + NCCL all-gather of bipartition keys, whole build on every rank") if world > 1 else ""),
                       "l2": "output per step (>= 2 B/pair) far exceeds the 126 MB L2; no explicit flush"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "device_s": dev_s, "wall_s": wall, "selfcheck": selfcheck, "consumer": consumer,
            "phases_ms": {("key_exchange" if sharded else "dedup"): st.get("ms_dedup"),
                          ("class_build_and_part_exchange" if sharded else "lists"): st.get("ms_incidence"),
                          "matrix": avg_ms, "note": "CUDA-event times of the last step on rank 0"},
        }
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize(dev)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(ctx, wl, mode, keys, off, leaf, P, editdist, steps, tree_arrays=None):
    """Host buffers in, host buffers out, through prf_hashrf / prf_tolerant; copies inside the timed region.
    With tree_arrays also the same job through prf_hashrf_trees / prf_tolerant_trees: parsed trees in (the
    bipartitions are extracted on the device), matrix out -- reported under e2e["from_trees"]."""
    import numpy as np
    import torch
    T = wl["T"]
    W = keys.shape[1]
    pin_in = torch.from_numpy(keys.view(np.int64)).pin_memory()
    pin_out = torch.empty(P, dtype=torch.int16).pin_memory()
    pin_mask = torch.empty((P + 31) // 32, dtype=torch.int32).pin_memory() if editdist >= 0 else None
    L = ctx._L
    import ctypes as C
    from phybin_b200._native import prf_stats
    st = prf_stats()
    offc = np.ascontiguousarray(off, dtype=np.uint64)
    leafc = np.ascontiguousarray(leaf, dtype=np.uint64) if leaf is not None else None

    def call():
        if mode == "hashrf":
            rc = L.prf_hashrf(ctx._h, pin_in.data_ptr(), offc.ctypes.data, T, W, editdist, pin_out.data_ptr(),
                              pin_mask.data_ptr() if pin_mask is not None else None, C.byref(st))
        else:
            rc = L.prf_tolerant(ctx._h, pin_in.data_ptr(), offc.ctypes.data, leafc.ctypes.data, T, W, editdist,
                                pin_out.data_ptr(), pin_mask.data_ptr() if pin_mask is not None else None, C.byref(st))
        ctx._ck(rc)

    for _ in range(2):  # warm-up: the first two calls grow the stream-ordered pool (10 GB result) and fault in the pinned buffers
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    h2d = int(keys.nbytes + offc.nbytes + (leafc.nbytes if leafc is not None else 0))
    d2h = int(P * 2 + ((P + 31) // 32 * 4 if editdist >= 0 else 0))
    res = {"value": P / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": dt * 1e3, "steps": steps,
           "phases_ms": {"h2d": st.ms_h2d, "dedup": st.ms_dedup, "incidence": st.ms_incidence,
                         "matrix": st.ms_matrix, "d2h": st.ms_d2h}}
    if tree_arrays is not None:
        from phybin_b200._native import prf_extract_info
        info = prf_extract_info()
        pins = [torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else (a.view(np.int64) if a.dtype == np.uint64 else a)).pin_memory()
                for a in tree_arrays]
        fn = L.prf_hashrf_trees if mode == "hashrf" else L.prf_tolerant_trees

        def call_trees():
            ctx._ck(fn(ctx._h, pins[0].data_ptr(), pins[1].data_ptr(), pins[2].data_ptr(), pins[3].data_ptr(), T, W, editdist,
                       pin_out.data_ptr(), pin_mask.data_ptr() if pin_mask is not None else None, C.byref(st), C.byref(info)))

        call_trees()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            call_trees()
        torch.cuda.synchronize()
        dt2 = (time.perf_counter() - t0) / steps
        res["from_trees"] = {"value": P / dt2, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                             "h2d_bytes_per_step": int(sum(a.nbytes for a in tree_arrays)), "d2h_bytes_per_step": d2h,
                             "extract_ms": info.ms_kernels, "h2d_ms": info.ms_h2d,
                             "api": "prf_hashrf_trees" if mode == "hashrf" else "prf_tolerant_trees"}
        del pins
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=0, help="override the number of trees (debug)")
    ap.add_argument("--ref-trees", type=int, default=0, help="CPU sample size in trees")
    ap.add_argument("--build", default="auto", choices=["auto", "sharded", "replicated"],
                    help="N > 1: hash-partitioned build over the GPUs (auto = sharded), or NCCL all-gather of the keys + the whole build on every rank")
    ap.add_argument("--quick-check", action="store_true", help="skip the host-side recomputation of n_hits and of 8 full rows")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sweep-full", action="store_true", help="cpu_baseline: add T = 10000 to the all-cores sweep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = dict(WORKLOADS[args.workload])
    if args.trees:
        wl["T"] = args.trees
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)
    return run_gpu(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
