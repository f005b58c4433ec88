"""Builds libxeq_b200.so (sm_100a only) in-tree with nvcc.  No torch headers are involved:
the library is a plain C ABI (include/xeq_b200.h) loaded through ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "build"
LIB = PKG / "libxeq_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-diag-suppress", "170",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libxeq_b200.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


SIMT_LIB = PKG / "libxeq_b200_simt.so"  # TEST-ONLY variant: also carries the round-1 SIMT edge kernels (-DXEQ_WITH_SIMT)


def build(force: bool = False, verbose: bool = False, simt: bool = False) -> Path:
    """Product library (default) or, with simt=True, the test-only variant used by the A/B parity test: the same
    sources plus the SIMT filter contraction of round 1, selected at run time with XEQ_EDGE_SIMT=1."""
    nvcc = _nvcc()
    obj_dir = OBJ / "simt" if simt else OBJ
    lib_path = SIMT_LIB if simt else LIB
    extra = ["-DXEQ_WITH_SIMT"] if simt else []
    obj_dir.mkdir(parents=True, exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "xeq_b200.h"]
    jobs = []
    only = {"edge_message"} if simt else None  # the other objects are shared with the product build
    for src in sources():
        obj = (obj_dir if (only is None or src.stem in only) else OBJ) / (src.stem + ".o")
        if only is not None and src.stem not in only:
            continue
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [(obj_dir if (only is not None and s.stem in only) else OBJ) / (s.stem + ".o") for s in sources()]
    if force or jobs or _stale(lib_path, objs):
        cmd = [nvcc, "-shared", "-o", str(lib_path), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    if "--simt" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, simt=True))
