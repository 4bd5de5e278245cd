"""A/B of two builds of the library on the same box: `python tools/gpu_ab_step.py` runs itself once per library,
alternating, and prints the step time of the bench workload (10 M x 512 bf16, C = Q = 200, T2T500+T2I0.25) and of
BASELINE config 1 (1 M x 512 fp32, T2T-500).  The library under test is chosen by patching _lib.LIB_PATH (tool only)."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    import torch
    from swat_b200 import _lib, synth
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    for opt in sys.argv[2:]:                     # e.g. dyn_tiles=0 (options the old build does not know are skipped)
        try:
            ctx.set_option(opt.split("=")[0], int(opt.split("=")[1]))
        except Exception as e:
            print("option skipped:", opt, e)
    def run(name, cap, img, q, reps=30):
        qs = _lib.Queries(ctx, q.float())
        kw = {"t2i_bank": img, "t2i_threshold": 0.25} if img is not None else {}
        for _ in range(4):
            _lib.topk(ctx, qs, cap, 500, 0.0, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scan = tail = 0.0
        e0.record()
        for _ in range(reps):
            _lib.topk(ctx, qs, cap, 500, 0.0, **kw)
            t = ctx.last_timing(); scan += t["scan_ms"]; tail += t["select_ms"] + t["t2i_ms"]
        e1.record(); torch.cuda.synchronize()
        print(f"{os.path.basename(sys.argv[1])} {' '.join(sys.argv[2:])} {name}: {e0.elapsed_time(e1)/reps:.4f} ms/step  scan {scan/reps:.4f}  select+walk {tail/reps:.4f}", flush=True)
        qs.close()
    qc, q, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
    cap, img, _ = synth.make_bank(10_000_000, qc, seed=0, device=dev, dtype=torch.bfloat16, chunk=1 << 20)
    run("cfg2 bf16 10M T2T+T2I", cap, img, q)
    run("cfg2 bf16 10M T2T", cap, None, q)
    del cap, img
    qc, q, _ = synth.make_queries(200, 1, seed=3, dtype=torch.float32)
    cap, img, _ = synth.make_bank(1_000_000, qc, seed=3, device=dev, dtype=torch.float32, chunk=1 << 18)
    run("cfg1 fp32 1M T2T", cap, None, q)
else:
    libs = ["tools/ab/libswat_b200_old.so", "swat_b200/libswat_b200.so"]
    for lib in libs * 2:
        subprocess.run([sys.executable, __file__, lib], check=False)
