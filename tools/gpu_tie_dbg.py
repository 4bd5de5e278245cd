"""First-call escalation depth on a bank whose k-th boundary sits inside the block of 1000 identical rows: old vs new build."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    os.environ["SWAT_DEBUG"] = "1"
    import torch
    from swat_b200 import _lib, synth
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    qc, q, _ = synth.make_queries(40, 1, seed=1, dtype=torch.bfloat16)
    cap, _, lab = synth.make_bank(2_000_000, qc, seed=1, device=dev, dtype=torch.bfloat16, chunk=1 << 20, with_images=False)
    qs = _lib.Queries(ctx, q.float())
    for i in range(3):
        s, r, _, c = _lib.topk(ctx, qs, cap, 500, 0.0); torch.cuda.synchronize()
        print(os.path.basename(sys.argv[1]), "call", i, json.dumps(ctx.last_timing()), "counts", int(c.sum()), flush=True)
else:
    for lib in ("tools/ab/libswat_b200_old.so", "swat_b200/libswat_b200.so"):
        subprocess.run([sys.executable, __file__, lib], check=False)
