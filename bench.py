#!/usr/bin/env python
"""bench.py -- the retrieval hot path on N B200s (BASELINE.json metric: bank rows scored + top-k'd / s).

    python bench.py [--gpus N --steps K --warmup W]            # our arm
    python bench.py --impl reference [...]                      # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): semi-aves, C = Q = 200 class prompts, full T2T500+T2I0.25
pipeline over a synthetic 10 M x 512 bf16 caption + image bank PER GPU (weak scaling; rows are
sharded over the ranks and the per-class candidates merged after one NCCL all-gather).
One step = one pass of the whole pipeline (scan + select + T2I walk [+ gather + merge]) over the bank.

`value`  : rows/s with the banks resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : rows/s through the C-ABI host entry point (swat_topk_host): banks in pinned host memory,
           H2D of the caption bank and of the candidates' image rows and D2H of the result inside
           the timed region.  For N > 1 every rank does that for its own host shard, then the [C,k]
           results are all-gathered and merged, all inside the timed region.
`roofline`: scan kernel, HBM bound for Q <= 209: 1 KB/row over the measured copy
           bandwidth in MEASURED_PEAKS.json.
`configs` : the other BASELINE.json configs, driver-timed in the same run (each entry: CUDA-event ms per whole step,
           scan-kernel ms, rows/s and both roofline fractions): at every N the config-4 shape (imagenet C = Q = 1000,
           T2T top-500, 50 M rows per GPU, NCCL merge), the strong-scaling point of config 5 (100 M rows split over
           the N GPUs, Q = 200) and `cold_call_ms` (first call on a fresh context, allocation and escalation included);
           at N = 1 also the query-count sweep Q in {64, 200, 400, 1000} over 50 M rows, the nine-dataset sweep of config 3
           over the same bank (class-mean prompts and synonym groups with MAX) and config 1 (fp32 banks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "bank_rows_scored_topk_per_sec"
UNIT = "rows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="bank rows per GPU")
    ap.add_argument("--classes", type=int, default=200)
    ap.add_argument("--k", type=int, default=500)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--t2t-only", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--overfetch", type=int, default=0, help="first k_fetch of the T2I walk (0 = library default)")
    ap.add_argument("--no-extras", action="store_true", help="skip the `configs` object (other BASELINE configs, cold call)")
    ap.add_argument("--extra-rows", type=int, default=50_000_000, help="rows per GPU of the config-4 / Q-sweep bank")
    ap.add_argument("--strong-rows", type=int, default=100_000_000, help="total rows of the strong-scaling point")
    return ap.parse_args()


def config_dict(a, world):
    """Identical for both arms (the driver compares them)."""
    return {"workload": workload_name(a, world), "rows_per_gpu": a.rows, "classes": a.classes, "k": a.k,
            "l2": f"inputs ({a.rows * 1024 / 1e9:.1f} GB per bank per GPU) far larger than the 126 MB L2; no flush needed"}


def workload_name(a, world):
    ds = {200: "semi-aves", 1000: "imagenet-shaped"}.get(a.classes, "synthetic")
    banks = "caption" if a.t2t_only else "caption+image"
    return (f"{ds} C={a.classes} Q={a.classes} {'T2T' if a.t2t_only else 'T2T+T2I0.25'} top-{a.k}, "
            f"{a.rows} x 512 bf16 {banks} rows per GPU, {world} GPU(s)")


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region is tens of
    milliseconds, shorter than nvidia-smi's polling period, so NVML is polled in-process from a
    thread (ctypes calls release the GIL); nvidia-smi is the fallback."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th, self.nvml = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
        self.th = threading.Thread(target=self._poll if self.nvml else self._smi, daemon=True)
        self.th.start()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.rows.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM), n.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                  n.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.max_sm = float(o[1]); self.rows.append((float(o[0]), int(o[2].strip(), 16), 0.0))
            except Exception:
                time.sleep(0.05)

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= int(r[1])
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(getattr(self, "max_sm", 0) or 0),
                "reasons": sorted(v for k, v in names.items() if bits & k), "samples": len(sm),
                "power_w_max": max(r[2] for r in self.rows)}


# ----------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_sample(a, n_rows, seed_shift=0):
    import torch
    from swat_b200 import synth
    qc, queries, _ = synth.make_queries(a.classes, 1, seed=a.seed, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(n_rows, qc, seed=a.seed + seed_shift, dtype=torch.bfloat16, chunk=1 << 16,
                                  tie_block=min(1000, n_rows // 8), with_images=not a.t2t_only)
    return queries.float().numpy(), cap.float().numpy(), None if img is None else img.float().numpy()


def cpu_vectorised(a, budget_s=12.0):
    """fp32 GEMM + exact tie-aware selection on all host cores (oracle.topk_walk)."""
    from oracle import swat_oracle as so
    n = 500_000
    q, cap, img = cpu_sample(a, n)
    so.topk_walk(cap[:50_000], q, a.k, 0.0, t2i_bank=None if img is None else img[:50_000])      # warm BLAS threads
    reps, t0 = 0, time.perf_counter()
    while reps < 1 or time.perf_counter() - t0 < budget_s:
        so.topk_walk(cap, q, a.k, 0.0, t2i_bank=None if img is None else img)
        reps += 1
    dt = time.perf_counter() - t0
    return n * reps / dt, f"{reps} x {n} rows x {a.classes} classes, vectorised fp32 GEMM + exact selection, {dt:.1f} s of CPU work"


def cpu_verbatim(a, n_rows, n_cls=None):
    """The reference's algorithm as written (per class: GEMV, Python sorted on zipped tuples, accept
    walk; sample_retrieval.py:774-825), unpartitioned: every class scans the whole sample."""
    from oracle import swat_oracle as so
    q, cap, img = cpu_sample(a, n_rows)
    if img is None:
        img = cap
    C = a.classes if n_cls is None else n_cls
    paths = [f"/s/{i % C}/{i}.jpg" for i in range(n_rows)]
    feats = {str(c): {"file_paths": paths, "feats": img, "caption_feats": cap} for c in range(C)}
    prompts = {str(c): {"mean": q[c]} for c in range(C)}
    fn = so.verbatim_t2t_ranked_sampler if a.t2t_only else so.verbatim_t2t_ranked_t2i_tshd_sampler
    t0 = time.perf_counter()
    fn(prompts, a.k, 0.0, feats)
    dt = time.perf_counter() - t0
    return dt * (a.classes / C), dt


REF_SAMPLE_ROWS = 50_000      # fixed: every run of the reference arm scores the same 50 000-row sample of the workload


def run_reference(a, rank, world):
    """--impl reference: the reference's own CPU path for this workload, timed on the host cores.
    /root/reference is Python + needs open_clip stubs and does not exist on the GPU box, so the arm
    runs the oracle's verbatim port (kind = "port") of sample_retrieval.py:774-825: per class a GEMV over the whole
    sample, Python sorted on the zipped tuples, the accept walk.  One step = one pass over a FIXED 50 000-row sample
    of the workload's bank (same generator, same 200 classes, unpartitioned like the GPU arm); the per-row rate does
    not depend on the sample size beyond the O(log n) of the sort."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    n_rows = int(os.environ.get("SWAT_BENCH_REF_ROWS", REF_SAMPLE_ROWS))
    for _ in range(a.warmup):
        cpu_verbatim(a, n_rows)
    dt = 0.0
    for _ in range(a.steps):
        dt += cpu_verbatim(a, n_rows)[1]            # the sampler call only, not the generation of the sample
    dt /= max(a.steps, 1)
    value = n_rows / dt
    vec, vec_sample = (None, None) if a.no_cpu else cpu_vectorised(a)
    sample = (f"{n_rows} rows x {a.classes} classes per step (fixed sample of the {a.rows}-row bank), unpartitioned, verbatim port of "
              f"sample_retrieval.py:774-825")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "vectorised_value": vec, "vectorised_sample": vec_sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------
def roof_rows_per_s(q_cols, bytes_per_row, hbm, tf):
    """the slower of bytes_per_row at HBM bandwidth and 2*512*Q flop/row at dense bf16."""
    h, t = hbm * 1e9 / bytes_per_row, tf * 1e12 / (1024.0 * q_cols)
    return (h, "hbm") if h <= t else (t, "tensor")


def run_extras(a, rank, world, dev, main):
    """The `configs` object: the other BASELINE.json configs and the cold call, timed like the main workload
    (warm-up, barrier + synchronize on both sides, CUDA events, max over ranks).  `main` = (cap, img, queries) of the
    main workload, still resident; they are released here."""
    import gc
    import torch
    import torch.distributed as dist
    from swat_b200 import _lib, synth
    from swat_b200 import dist as sdist
    hbm, tf, src = peaks()
    out = {"peaks": f"{src}: hbm {hbm} GB/s, bf16 sustained {tf} TFLOP/s"}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(ctx, step, n_local, q_cols, bytes_per_row, warm=3, reps=3):
        for _ in range(warm):
            step()
        scans = ctx.__dict__["_time_scans"] = []
        sampler = ClockSampler(dev.index) if rank == 0 else None
        if sampler:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern = []
        e0.record()
        for _ in range(reps):
            step()
            if world == 1:
                tm = ctx.last_timing()
                kern.append(tm["scan_ms"] / max(tm["scan_launches"], 1.0))      # one pass over the bank (all of its launches)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ctx.__dict__["_time_scans"] = None
        ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        if world > 1:
            kern = [x.elapsed_time(y) for x, y in scans]
        kern.sort()
        kms = max_over_ranks(kern[len(kern) // 2]) if kern else None
        roof, bound = roof_rows_per_s(q_cols, bytes_per_row, hbm, tf)
        return {"rows_per_gpu": n_local, "Q": q_cols, "ms_per_step": ms, "scan_kernel_ms": kms,
                "rows_per_s": n_local * world / (ms * 1e-3), "clocks": clocks,
                "roofline": {"bound": bound, "roof_rows_per_s_per_gpu": roof, "step_frac": n_local / (ms * 1e-3) / roof,
                             "kernel_frac": None if not kms else n_local / (kms * 1e-3) / roof}}

    def t2t_step(ctx, qs, cap, row_offset):
        if world == 1:
            return lambda: _lib.topk(ctx, qs, cap, a.k, 0.0, row_offset=row_offset)
        return lambda: sdist.topk_sharded(ctx, qs, cap, a.k, 0.0, row_offset=row_offset, world=world)

    # ---- cold call: fresh context + query set, first call of the main workload (allocation, threshold bootstrap and
    # any over-fetch escalation included; the banks are resident)
    cap, img, queries = main
    main.clear()
    barrier()
    t0 = time.perf_counter()
    ctx_c = _lib.Context(dev.index)
    qs_c = _lib.Queries(ctx_c, queries.float())
    t1 = time.perf_counter()
    if world == 1:
        _lib.topk(ctx_c, qs_c, cap, a.k, 0.0, t2i_bank=img, t2i_threshold=0.25)
    else:
        sdist.topk_sharded(ctx_c, qs_c, cap, a.k, 0.0, t2i_bank=img, t2i_threshold=0.25, row_offset=rank * a.rows, world=world)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    out["cold_call_ms"] = max_over_ranks((t2 - t0) * 1e3)
    out["cold_call_breakdown_ms"] = {"ctx_and_queries_create": (t1 - t0) * 1e3, "first_pipeline_call": (t2 - t1) * 1e3,
                                     "first_call_device_ms": ctx_c.last_timing()["total_ms"] if world == 1 else None,
                                     "first_call_scans": ctx_c.last_timing()["scan_launches"] if world == 1 else None}
    out["cold_call_note"] = ("wall clock on a fresh context: swat_ctx_create + swat_queries_create + first whole-pipeline call on the main "
                             "workload (buffer allocation, threshold bootstrap and over-fetch escalation included; banks resident)")
    qs_c.close(); ctx_c.close()
    del cap, img
    gc.collect(); torch.cuda.empty_cache()

    ctx = _lib.Context(dev.index)
    # ---- config 4 shape: imagenet C = Q = 1000, T2T top-500, `extra_rows` rows per GPU (+ NCCL merge for N > 1);
    # at N = 1 the same bank serves the query-count sweep (config 5 on one GPU)
    n_x = a.extra_rows
    qc, q1000, _ = synth.make_queries(1000, 1, seed=a.seed + 1, dtype=torch.bfloat16)
    capx, _, _ = synth.make_bank(n_x, qc, seed=a.seed + 1, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False,
                                 row_offset=rank * n_x)
    if world == 1:
        # ---- config 3: the nine datasets of the paper over this one bank, class-mean prompts (Q = C) and synonym groups
        # with a per-class MAX (Q = S); (C, S) from SURVEY.md 8a
        out["cfg3_sweep9"] = []
        for name, C3, S3 in (("flowers102", 102, 343), ("fgvc-aircraft", 100, 271), ("eurosat", 10, 51), ("dtd", 47, 75), ("food101", 101, 413),
                             ("oxford_pets", 37, 114), ("stanford_cars", 196, 1221), ("semi-aves", 200, 400), ("imagenet", 1000, 5191)):
            for mode in ("class-mean", "synonyms-max"):
                if mode == "class-mean":
                    qs = _lib.Queries(ctx, qc[:C3].float())
                    Q3 = C3
                else:
                    sizes = [1] * C3                                  # S3 synonyms over C3 classes, deterministic and uneven
                    left, i = S3 - C3, 0
                    while left > 0:
                        add = min(left, 1 + (i * 7) % max(1, 2 * (S3 // C3)))
                        sizes[(i * 37) % C3] += add
                        left -= add
                        i += 1
                    g = torch.Generator().manual_seed(C3 * 1000 + S3)
                    coq = torch.repeat_interleave(torch.arange(C3, dtype=torch.int32), torch.tensor(sizes))
                    u = torch.nn.functional.normalize(torch.randn(coq.numel(), 512, generator=g), dim=-1)
                    q3 = torch.nn.functional.normalize(qc[:C3][coq.long()].float() + 0.3 * u, dim=-1).to(torch.bfloat16).float()
                    qs = _lib.Queries(ctx, q3, coq, C3, "max")
                    Q3 = S3
                e = timed(ctx, t2t_step(ctx, qs, capx, 0), n_x, Q3, 1024.0, warm=3, reps=2)
                e.pop("clocks", None)
                e["workload"] = f"{name}: C = {C3}, Q = {Q3} ({mode}), T2T top-{a.k}, {n_x} x 512 bf16 rows"
                out["cfg3_sweep9"].append(e)
                qs.close()
    sweep = (64, 200, 400, 1000) if world == 1 else (1000,)
    out["cfg5_qsweep"] = []
    for Q in sweep:
        qs = _lib.Queries(ctx, q1000[:Q].float())
        e = timed(ctx, t2t_step(ctx, qs, capx, rank * n_x), n_x, Q, 1024.0)
        e["workload"] = f"C = Q = {Q}, T2T top-{a.k}, {n_x} x 512 bf16 rows per GPU, {world} GPU(s)"
        if Q == 1000:
            e["workload"] = f"imagenet-shaped (BASELINE config 4): " + e["workload"]
            out["cfg4_imagenet"] = e
        out["cfg5_qsweep"].append(e)
        qs.close()
    del capx
    gc.collect(); torch.cuda.empty_cache()
    # ---- strong scaling (config 5): a fixed 100 M-row bank split over the N GPUs, Q = 200
    n_s = a.strong_rows // world
    qc2, q200, _ = synth.make_queries(200, 1, seed=a.seed + 2, dtype=torch.bfloat16)
    caps, _, _ = synth.make_bank(n_s, qc2, seed=a.seed + 2, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False,
                                 row_offset=rank * n_s)
    qs = _lib.Queries(ctx, q200.float())
    e = timed(ctx, t2t_step(ctx, qs, caps, rank * n_s), n_s, 200, 1024.0)
    e["workload"] = f"strong scaling: {a.strong_rows} rows in total over {world} GPU(s), C = Q = 200, T2T top-{a.k}"
    e["total_rows"] = n_s * world
    out["cfg5_strong_scaling"] = e
    qs.close()
    del caps
    gc.collect(); torch.cuda.empty_cache()
    # ---- config 1 (N = 1): fp32 banks, the reference's dtype, 2 KB/row
    if world == 1:
        out["cfg1_fp32"] = []
        qc3, q3, _ = synth.make_queries(200, 1, seed=a.seed + 3, dtype=torch.float32)
        for n1 in (1_000_000, 10_000_000):
            c32, i32, _ = synth.make_bank(n1, qc3, seed=a.seed + 3, device=dev, dtype=torch.float32, chunk=1 << 18)
            qs = _lib.Queries(ctx, q3)
            for name, kw in (("T2T", {}), ("T2T+T2I0.25", {"t2i_bank": i32})):
                e = timed(ctx, lambda: _lib.topk(ctx, qs, c32, a.k, 0.0, **kw), n1, 200, 2048.0)
                e["workload"] = f"semi-aves C = Q = 200 {name} top-{a.k}, {n1} x 512 fp32 rows (BASELINE config 1 dtype), 1 GPU"
                e["escalations_per_step"] = ctx.last_timing()["escalations"]
                out["cfg1_fp32"].append(e)
            qs.close()
            del c32, i32
            gc.collect(); torch.cuda.empty_cache()
    ctx.close()
    return out


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from swat_b200 import _lib, synth
    from swat_b200 import dist as sdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    all_cpus = os.sched_getaffinity(0)          # restored for the cpu_baseline leg, which uses every host core
    if True:
        # pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated: the e2e leg streams
        # 10 GB per rank out of host memory and should not cross the socket interconnect (at N = 1 too: the same build
        # measured 25-49 M rows/s end to end depending on where the scheduler had put the process)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = f"rank pinned to {len(cpus)} CPUs local to GPU {local_rank}"
        except Exception as e:          # affinity is an optimisation only
            numa = f"affinity not set: {type(e).__name__}"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.Context(local_rank)
    if a.overfetch:
        ctx.set_option("overfetch", a.overfetch)
    n_local = a.rows
    row_offset = rank * n_local
    qc, queries, _ = synth.make_queries(a.classes, 1, seed=a.seed, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(n_local, qc, seed=a.seed, device=dev, dtype=torch.bfloat16, chunk=1 << 20,
                                  with_images=not a.t2t_only, row_offset=row_offset)
    qs = _lib.Queries(ctx, queries.float())
    k = a.k

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the caller owns the result tensors: one set, reused by every step (what a loop over shards does)
    out = (torch.empty(a.classes, k, dtype=torch.float32, device=dev), torch.empty(a.classes, k, dtype=torch.int64, device=dev),
           None if img is None else torch.empty(a.classes, k, dtype=torch.float32, device=dev),
           torch.empty(a.classes, dtype=torch.int32, device=dev))

    def step():
        if world == 1:
            return _lib.topk(ctx, qs, cap, k, 0.0, t2i_bank=img, t2i_threshold=0.25, row_offset=row_offset, out=out)
        return sdist.topk_sharded(ctx, qs, cap, k, 0.0, t2i_bank=img, t2i_threshold=0.25, row_offset=row_offset, world=world)

    for _ in range(max(a.warmup, 3)):
        res = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, scan_launches = [], 0
    scan_events = ctx.__dict__["_time_scans"] = []      # swat_b200.dist records (start, end) events around each scan
    barrier()
    ev0.record()
    for _ in range(a.steps):
        res = step()
        if world == 1:
            tm = ctx.last_timing()
            scan_ms.append(tm["scan_ms"] / max(tm["scan_launches"], 1.0))     # average scan launch of this step
            scan_launches += int(tm["scan_launches"])
    ev1.record()
    barrier()
    ctx.__dict__["_time_scans"] = None
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    escalations = ctx.last_timing()["escalations"] if world == 1 else 0.0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_rows = n_local * world
    value = total_rows * a.steps / (ms * 1e-3)

    # ---- scan-kernel roofline (rank 0's kernel; CUDA events on the launching stream)
    hbm, tf, src = peaks()
    if not scan_ms:                                     # multi-GPU path: events recorded around every scan of the timed steps
        scan_ms = [e0.elapsed_time(e1) for e0, e1 in scan_events]
        scan_launches = len(scan_ms)
    scan_ms.sort()
    scan = scan_ms[len(scan_ms) // 2]
    q_cols = a.classes
    bytes_per_row = 1024.0
    hbm_time = bytes_per_row / (hbm * 1e9)
    tc_time = 2.0 * 512 * q_cols / (tf * 1e12)
    if hbm_time >= tc_time:
        achieved = n_local * bytes_per_row / (scan * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm}
    else:
        achieved = n_local * 2.0 * 512 * q_cols / (scan * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf}
    traffic = None
    tp = os.path.join(REPO, "profiles", "scan_traffic.json")
    if os.path.exists(tp) and a.rows == 10_000_000 and a.classes == 200:      # the ncu capture is of this workload
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roof.update({"traffic": traffic, "kernel": "scan_tc_kernel",
                 "peak_source": f"{src} (MEASURED_PEAKS.json: " + ("burst copy bandwidth)" if roof["bound"] == "hbm" else "sustained cuBLAS bf16)"),
                 "kernel_ms": scan, "launches_in_timed_region": scan_launches,
                 "algorithmic_bytes_per_launch": n_local * bytes_per_row,
                 "note": "every scan launch streams the whole caption bank once (1 KB/row); a step has one launch plus one per "
                         "T2I over-fetch escalation"})

    # ---- end to end through the C-ABI with host buffers
    e2e = None
    if not a.no_e2e:
        h_cap = torch.empty(cap.shape, dtype=cap.dtype, pin_memory=True); h_cap.copy_(cap)
        h_img = None
        if img is not None:
            h_img = torch.empty(img.shape, dtype=img.dtype, pin_memory=True); h_img.copy_(img)
        torch.cuda.synchronize(dev)
        steps_e = max(2, min(a.steps, 5))

        def e2e_step():
            if world == 1:
                r = _lib.topk_host(ctx, qs, h_cap, k, 0.0, t2i_bank=h_img, t2i_threshold=0.25, row_offset=row_offset)
                tm = ctx.last_timing()
                return r, tm["h2d_bytes"], tm["d2h_bytes"]
            # every rank streams ITS shard from pinned host memory through the C-ABI host entry point (exact local
            # walk), then one all-gather of the [C,k] results and an associative merge
            s, r, t, c = _lib.topk_host(ctx, qs, h_cap, k, 0.0, t2i_bank=h_img, t2i_threshold=0.25, row_offset=row_offset)
            tm = ctx.last_timing()
            local = (s.cuda(dev), r.cuda(dev), None if t is None else t.cuda(dev), c.cuda(dev),
                     torch.full((c.numel(),), float("-inf"), dtype=torch.float32, device=dev))      # exact local walks: no limit
            res_m = sdist.gather_merge(local, k, world, ctx=ctx)
            out = [x.cpu() for x in res_m[:4] if x is not None]
            return res_m, tm["h2d_bytes"] + sum(x.numel() * x.element_size() for x in local if x is not None), \
                tm["d2h_bytes"] + sum(x.numel() * x.element_size() for x in out)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            r, h2d, d2h = e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        e2e_tm = ctx.last_timing()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": total_rows * steps_e / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "steps": steps_e, "ms_per_step": dt / steps_e * 1e3,
               "last_step_device_ms": {"caption_stream_and_scan": e2e_tm["scan_ms"], "select": e2e_tm["select_ms"],
                                       "rescore_and_walk": e2e_tm["t2i_ms"], "total": e2e_tm["total_ms"],
                                       "escalations": e2e_tm["escalations"]},
               "numa": numa,
               "api": "swat_topk_host (C-ABI, pinned host banks)" if world == 1 else
                      "swat_topk_host per rank on its pinned host shard + one NCCL all-gather + swat_merge_topk"}
        del h_cap, h_img

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        for tid in os.listdir("/proc/self/task"):          # every thread: OpenMP workers created while pinned keep their mask
            try:
                os.sched_setaffinity(int(tid), all_cpus)
            except OSError:
                pass
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        vec, vec_sample = cpu_vectorised(a)
        est, dt8 = cpu_verbatim(a, 50_000, n_cls=8)
        cpu = {"value": vec, "unit": UNIT, "cores": cores, "kind": "port", "sample": vec_sample,
               "verbatim_value": 50_000 / est,
               "verbatim_sample": f"50000 rows x 8 of {a.classes} classes scaled x{a.classes / 8:.0f}, verbatim port of "
                                  f"sample_retrieval.py:774-825 ({dt8:.1f} s)"}
    counts = res[3]
    accepted = int(counts.sum().item())
    extras = None
    if not a.no_extras:
        main = [cap, img, queries]
        del cap, img, res, counts
        qs.close()
        extras = run_extras(a, rank, world, dev, main)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": config_dict(a, world),
            "stats": {"accepted_rows": accepted, "t2i_escalations_per_step": escalations,
                      "step_frac_of_roofline": (n_local / (ms / a.steps * 1e-3)) / roof_rows_per_s(q_cols, bytes_per_row, hbm, tf)[0]},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "configs": extras}))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        # launched without torchrun: re-exec under it (one rank per GPU over NCCL)
        port = 29500 + os.getpid() % 1000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
