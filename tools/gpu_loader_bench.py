"""Loader row: time from a feature file on disk to banks resident in HBM.
  reference route : torch.load of the pickled {caption_features, image_features, labels, filepath} dict, then .cuda()
  flat shard route: FlatShard (mmap) -> to_device (double-buffered pinned staging)
Files are written and read back in the same process, so both routes read from the page cache."""
import os, sys, time, tempfile, torch
sys.path.insert(0, ".")
from swat_b200 import shards, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
qc, _, _ = synth.make_queries(200, 1, seed=0, dtype=torch.bfloat16)
cap, img, labels = synth.make_bank(N, qc, seed=0, device="cuda", dtype=torch.bfloat16, chunk=1 << 20)
cap, img, labels = cap.cpu(), img.cpu(), labels.cpu()
paths, _ = synth.make_paths(labels)
tmp = tempfile.mkdtemp(dir="/tmp")
pth = os.path.join(tmp, "mined.pth")
t0 = time.perf_counter(); shards.save_mined_pth(pth, cap.float(), img.float(), labels, paths); t_save = time.perf_counter() - t0
t0 = time.perf_counter(); shards.convert_pth_to_flat(pth, os.path.join(tmp, "flat"), "bf16"); t_conv = time.perf_counter() - t0
print(f"{N:,} rows: mined.pth {os.path.getsize(pth) / 1e9:.2f} GB (fp32, written in {t_save:.1f} s); converted to a bf16 flat shard in {t_conv:.1f} s")
for rep in range(2):
    t0 = time.perf_counter()
    d = torch.load(pth, map_location="cpu", weights_only=False)
    c = d["caption_features"].cuda(); i = d["image_features"].cuda(); torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0
    del d, c, i
    t0 = time.perf_counter()
    d = shards.load_mined_pth(pth)          # mmap
    c = d["caption_features"].cuda(); i = d["image_features"].cuda(); torch.cuda.synchronize()
    t_mmap = time.perf_counter() - t0
    del d, c, i
    t0 = time.perf_counter()
    fs = shards.FlatShard(os.path.join(tmp, "flat"))
    t_open = time.perf_counter() - t0
    c, i = fs.to_device("cuda:0")          # C-ABI loader: swat_bank_load (GDS or pread + pinned double buffer)
    t_flat = time.perf_counter() - t0
    del c, i
    t0 = time.perf_counter()
    c, i = fs.to_device("cuda:0", native=False, pinned_chunk_rows=1 << 18)
    t_torch = time.perf_counter() - t0
    gb = 2 * N * 1024 / 1e9
    print(f"run {rep}: torch.load + .cuda() {t_ref:.2f} s | torch.load(mmap) + .cuda() {t_mmap:.2f} s | "
          f"FlatShard open {t_open * 1e3:.1f} ms, swat_bank_load to HBM {t_flat:.2f} s ({gb / t_flat:.1f} GB/s for {gb:.1f} GB of bf16, "
          f"GDS={fs.used_gds}) | torch staging {t_torch:.2f} s")
    del c, i, fs
