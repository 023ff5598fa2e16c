"""CPU-only: the C-ABI shared libraries load and export every symbol their headers declare
(no compute is called — that needs a GPU and lives in the -m gpu tests)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: Path, prefix: str):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(rf"\b({prefix}\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    from pupiloptixlab_b200 import build
    build.build_pb2()
    build.build_kat()
    build.build_host()
    return build.BUILD


def test_libpb2_exports_header_symbols(built):
    names = _declared(ROOT / "include" / "pb2.h", "pb2_")
    assert len(names) >= 25
    lib = ctypes.CDLL(str(built / "libpb2.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_kat_hooks_live_in_their_own_library(built):
    """the known-answer hooks are test infrastructure: declared in include/pb2_kat.h, exported by libpb2_kat.so, absent from libpb2.so"""
    names = _declared(ROOT / "include" / "pb2_kat.h", "pb2_")
    assert names == ["pb2_kat", "pb2_kat_last_error"]
    ctypes.CDLL(str(built / "libpb2.so"), mode=ctypes.RTLD_GLOBAL)
    hooks = ctypes.CDLL(str(built / "libpb2_kat.so"))
    assert all(hasattr(hooks, n) for n in names)
    import subprocess
    exported = subprocess.run(["nm", "-D", "--defined-only", str(built / "libpb2.so")], capture_output=True, text=True).stdout
    assert "kat" not in exported


def test_pb2_python_binding_matches_header(built):
    from pupiloptixlab_b200 import pb2
    L = pb2.lib()
    for n in _declared(ROOT / "include" / "pb2.h", "pb2_"):
        assert hasattr(L, n)


def test_no_gpu_fails_loudly(built):
    """the product has no CPU path: without a device, init raises instead of falling back"""
    from pupiloptixlab_b200 import pb2
    if pb2.lib().pb2_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb2.Pb2Error):
        pb2.init(0)


def test_product_never_imports_oracle():
    """only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may touch oracle/"""
    pkg = ROOT / "pupiloptixlab_b200"
    offenders = []
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        if "_build" in p.parts:
            continue
        t = p.read_text(errors="ignore")
        if re.search(r"oracle/|liborc_|orc_[a-z_]+\(|import orc\b|from orc\b", t):
            offenders.append(str(p.relative_to(ROOT)))
    assert not offenders, offenders
