"""In-tree builds (no JIT cache): the CUDA back end libpb2.so for sm_100a and the C++ host library.

    python -m pupiloptixlab_b200.build [--force]

Outputs go to pupiloptixlab_b200/_build/ (git-ignored; they travel to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
# PB2_BUILD_DIR / PB2_NVCC_EXTRA: side-by-side variant builds for A/B measurements (tools/, profiles/); the product uses _build
BUILD = PKG / os.environ.get("PB2_BUILD_DIR", "_build")
CSRC = PKG / "csrc"
HOST = PKG / "host"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v", *os.environ.get("PB2_NVCC_EXTRA", "").split()]
# Shading arithmetic (wavefront.cu and its known-answer twin kat.cu) is built the way the reference builds its device code
# where that is cheap and harmless: approximate division / square root (<= 2 ulp) and flush-to-zero, three of the switches
# its -use_fast_math implies (CMakeLists.txt:45).  Measured on B200: k_shade -8 % (Cornell) / -23 % (material grid), and
# every parity test keeps its tolerance (profiles/README.md).  The fast transcendental substitutions (__sinf, __powf, ...)
# are NOT taken: +1 % more, at 10x the error.  Camera rays, ray/primitive intersection and the BVH builder spell their
# roundings out with _rn intrinsics or are built IEEE, so hit records do not depend on these flags.
FAST_SHADING_FLAGS = ["--prec-div=false", "--prec-sqrt=false", "--ftz=true"]
FAST_SHADING_FILES = {"wavefront.cu", "kat.cu"}
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, log: Path | None = None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd))}")
    return r


def build_pb2(force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    out = BUILD / "libpb2.so"
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "pb2.h"]
    objs = []

    def compile_one(src: Path):
        obj = BUILD / (src.stem + ".o")
        if force or _newer(obj, [src, *hdrs]):
            extra = FAST_SHADING_FLAGS if src.name in FAST_SHADING_FILES and "PB2_IEEE_SHADING" not in os.environ else []
            _run([NVCC, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)], BUILD / (src.stem + ".ptxas.log"))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _newer(out, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *map(str, objs), "-o", str(out), "-lcudart"])
    return out


def build_kat(force: bool = False) -> Path:
    """libpb2_kat.so: the per-function known-answer hooks of the -m gpu tests (csrc/test_hooks/kat.cu, include/pb2_kat.h).
    Test infrastructure, kept out of the product library; it links against libpb2.so for the record conversions."""
    BUILD.mkdir(exist_ok=True)
    out, src, obj = BUILD / "libpb2_kat.so", CSRC / "test_hooks" / "kat.cu", BUILD / "kat.o"
    hdrs = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "pb2.h", ROOT / "include" / "pb2_kat.h"]
    if force or _newer(obj, [src, *hdrs]):
        extra = FAST_SHADING_FLAGS if "PB2_IEEE_SHADING" not in os.environ else []
        _run([NVCC, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)], BUILD / "kat.ptxas.log")
    if force or _newer(out, [obj, BUILD / "libpb2.so"]):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", str(obj), "-o", str(out), f"-L{BUILD}", "-lpb2", "-lcudart",
              "-Xlinker", "-rpath,$ORIGIN"])
    return out


def build_host(force: bool = False) -> Path:
    """libpupil_host.so (the C++ host surface + its C entry points) and the headless path_tracer executable."""
    BUILD.mkdir(exist_ok=True)
    out, exe = BUILD / "libpupil_host.so", BUILD / "path_tracer"
    lib_srcs = sorted(p for p in HOST.glob("*.cpp") if p.name != "main.cpp")
    hdrs = sorted(HOST.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    common = [CXX, "-O2", "-std=c++20", "-fPIC", "-pthread", "-Wall", "-Wno-unused-function", f"-I{ROOT / 'include'}", f"-I{HOST}"]
    if force or _newer(out, [*lib_srcs, *hdrs]):
        _run([*common, "-shared", *map(str, lib_srcs), "-o", str(out), f"-L{BUILD}", "-lpb2", "-lz", "-Wl,-rpath,$ORIGIN"])
    if force or _newer(exe, [HOST / "main.cpp", out, *hdrs]):
        _run([*common, str(HOST / "main.cpp"), "-o", str(exe), f"-L{BUILD}", "-lpupil_host", "-lpb2", "-Wl,-rpath,$ORIGIN"])
    return out


def build_all(force: bool = False) -> None:
    build_pb2(force)
    build_kat(force)
    build_host(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built:", *sorted(p.name for p in BUILD.glob("*.so")))
