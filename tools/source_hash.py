"""sha256 over the CUDA sources of libpb2.so: measured-by-ncu figures committed under profiles/ carry it, and bench.py
reports them only while the kernels they were measured on are the kernels it runs."""
import hashlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernel_source_sha() -> str:
    h = hashlib.sha256()
    for p in sorted((ROOT / "pupiloptixlab_b200" / "csrc").glob("*.cu*")):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(kernel_source_sha())
