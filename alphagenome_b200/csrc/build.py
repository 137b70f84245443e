"""Build libagb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

Usage: python -m alphagenome_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
REPO = PKG.parent
OUT = PKG / "libagb200.so"
OBJ_DIR = HERE / "build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-I", str(REPO / "include"),
]


def sources() -> list[Path]:
    return sorted(HERE.glob("*.cu"))


def _digest(paths: list[Path]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> Path:
    """trace=True: a separate libagb200_trace.so compiled with -DAGB_TRACE (bring-up timelines; load it with AGB200_LIB)."""
    if trace:
        return _build_variant("trace", ["-DAGB_TRACE"])
    srcs = sources()
    deps = srcs + sorted(HERE.glob("*.cuh")) + sorted(HERE.glob("*.h")) + [REPO / "include" / "agb200.h"]
    stamp = OBJ_DIR / "stamp.txt"
    dig = _digest(deps)
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == dig:
        return OUT
    OBJ_DIR.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", str(OUT), *map(str, objs), "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return OUT


def _build_variant(tag: str, flags: list[str]) -> Path:
    out = PKG / f"libagb200_{tag}.so"
    odir = OBJ_DIR / tag
    odir.mkdir(parents=True, exist_ok=True)
    def compile_one(src: Path) -> Path:
        obj = odir / (src.stem + ".o")
        r = subprocess.run([NVCC, *ARCH_FLAGS, *COMMON, *flags, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([NVCC, *ARCH_FLAGS, "-shared", "-o", str(out), *map(str, objs), "-Xlinker", "--exclude-libs,ALL"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv)
    print(p)
