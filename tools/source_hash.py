"""sha256 over the CUDA sources that decide what k_shade / k_extend / k_shadow execute and what tree they walk: measured-by-ncu
figures committed under profiles/ carry it, and bench.py reports them only while the kernels they were measured on are the
kernels it runs."""
import hashlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernel_source_sha() -> str:
    h = hashlib.sha256()
    csrc = ROOT / "pupiloptixlab_b200" / "csrc"
    for p in [csrc / n for n in ("bvh_build.cu", "bvh_ploc.cu", "pb2_types.cuh", "pt_math.cuh", "radix_sort.cu", "traverse.cuh", "vecmath.cuh", "wavefront.cu")]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(kernel_source_sha())
