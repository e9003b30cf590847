"""Build libecloop_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

Fifteen translation units compile in parallel: the host API + small kernels (twice: product and -DECL_EXPERIMENTAL test
library), the GPU blf-gen insert (CUB sort), and one fused add-kernel variant per
address-type / endomorphism combination, each for a filter in shared memory and for a filter in HBM. Objects are cached under build/ keyed by a hash of all sources + flags.
"""
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
OUT = PKG / "libecloop_b200.so"
OUT_EXP = PKG / "libecloop_b200_exp.so"  # the same library with -DECL_EXPERIMENTAL (FP64-pipe field arithmetic): test-only
OBJ = ROOT / "build" / "obj"
ADD_VARIANTS = (1, 2, 3, 5, 6, 7)

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _sources_hash(extra: str = "") -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu*")) + list(CSRC.glob("*.inc")) + [ROOT / "include" / "ecloop_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(extra.encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), variant: str | None = None) -> Path:
    """Compile (if stale) and return the path of libecloop_b200.so. `variant` builds a tuning variant with the
    given -D defines into build/variants/libecloop_b200_<variant>.so (tools/build_variants.py)."""
    global OUT, OBJ, OUT_EXP
    if variant:
        out, obj = ROOT / "build" / "variants" / f"libecloop_b200_{variant}.so", ROOT / "build" / f"obj_{variant}"
        saved = (OUT, OBJ, OUT_EXP)
        OUT, OBJ, OUT_EXP = out, obj, out.with_name(out.stem + "_exp.so")
        try:
            out.parent.mkdir(parents=True, exist_ok=True)
            return build(force, verbose, defines)
        finally:
            OUT, OBJ, OUT_EXP = saved
    tag = _sources_hash(" ".join(defines))
    stamp = OBJ / "stamp.txt"
    if not force and OUT.exists() and OUT_EXP.exists() and stamp.exists() and stamp.read_text() == tag:
        return OUT
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    jobs = [(CSRC / "ecl_api.cu", OBJ / "ecl_api.o", []), (CSRC / "filter_add.cu", OBJ / "filter_add.o", []),
            (CSRC / "ecl_api.cu", OBJ / "ecl_api_exp.o", ["-DECL_EXPERIMENTAL"])]
    for v in ADD_VARIANTS:
        jobs.append((CSRC / "add_inst.cu", OBJ / f"add_inst_{v}.o", [f"-DADD_VARIANT={v}"]))
        jobs.append((CSRC / "add_inst.cu", OBJ / f"add_inst_hbm_{v}.o", [f"-DADD_VARIANT={v}", "-DADD_HBM=1"]))
    dflags = [f"-D{d}" for d in defines]

    def compile_one(job):
        src, obj, extra = job
        cmd = [nvcc, *NVCC_FLAGS, *dflags, *extra, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name} {extra}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for (src, obj, extra), log in zip(jobs, logs):
            sys.stderr.write(f"--- {src.name} {extra}\n{log}\n")
    for out, skip in ((OUT, "ecl_api_exp.o"), (OUT_EXP, "ecl_api.o")):
        link = [nvcc, "-shared", "-o", str(out), *[str(j[1]) for j in jobs if j[1].name != skip]]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(tag)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
