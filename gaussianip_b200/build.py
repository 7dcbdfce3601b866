"""In-tree build of libgsb.so (sm_100a only).  ``python -m gaussianip_b200.build``.

nvcc cross-compiles without a GPU; the built library sits next to this file (git-ignored,
but shipped to the GPU box with the repo snapshot)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT = PKG / "libgsb.so"
OBJ = PKG / "build"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"),
          "-I", str(CSRC), "--expt-relaxed-constexpr"]
# preprocess.cu must not contract a*b+c into FMA: the CPU oracle rounds every op separately.
PER_FILE = {"preprocess.cu": ["-fmad=false"]}
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "render_fwd.cu", "render_bwd.cu", "preprocess_bwd.cu",
           "standin.cu", "adam.cu", "knn.cu", "compact.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stamp(src: Path, flags) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)

    def compile_one(name: str):
        src = CSRC / name
        flags = ARCH + COMMON + PER_FILE.get(name, []) + (["-Xptxas", "-v"] if ptxas_info else []) + \
            os.environ.get("GSB_NVCC_EXTRA", "").split()          # e.g. -DGSB_FWD_STAGES=2 for tuning sweeps
        obj = OBJ / (name + ".o")
        stamp_file = OBJ / (name + ".stamp")
        stamp = _stamp(src, flags)
        if not force and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
            return obj, ""
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        stamp_file.write_text(stamp)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [str(o) for o, _ in results]
    log = "".join(l for _, l in results)
    if verbose and log:
        print(log, file=sys.stderr)
    newest = max(Path(o).stat().st_mtime for o in objs)
    if force or not OUT.exists() or OUT.stat().st_mtime < newest:
        cmd = [nvcc, *ARCH, "-shared", "-o", str(OUT), *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(p)
