"""Build libswat_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python swat_b200/csrc/build.py [--force]

Sources are compiled to objects in swat_b200/csrc/build/ and linked into swat_b200/libswat_b200.so
(git-ignored; it travels to the GPU box with the gpurun snapshot).  cudart is linked statically and
the driver API is reached through cudaGetDriverEntryPoint, so the library loads on a box without
libcuda (the CPU-only build container) and fails at swat_ctx_create instead.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libswat_b200.so")
BUILD = os.path.join(HERE, "build")
SOURCES = ["api.cu", "scan_tc.cu", "scan_simt.cu", "select.cu", "loader.cu"]
HEADERS = ["common.cuh", "epilogue.cuh", "scan_tc.h", "tc_ptx.cuh", os.path.join(REPO, "include", "swat_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--threads", "4"]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f) if not os.path.isabs(f) else f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def cc(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
