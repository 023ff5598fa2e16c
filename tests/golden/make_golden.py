"""Regenerate tests/golden/*.npz from the REFERENCE build of the oracle (oracle/_ref/liborc_ref.so,
i.e. the reference's own BSDF / emitter / sampling / RNG headers compiled on the host).

Run in the authoring container, where /root/reference exists:
    python tests/golden/make_golden.py
The .npz files are committed; the GPU box has no reference tree and reads only these.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE.parent.parent), str(HERE.parent)]

import kat  # noqa: E402
import orc  # noqa: E402

if __name__ == "__main__":
    lib = orc.ref()
    assert lib is not None and lib.orc_backend_name() == b"reference", "needs the reference tree"
    np.savez_compressed(HERE / "kat_reference.npz", **kat.run(lib))
    np.savez_compressed(HERE / "render_reference.npz", **kat.run_renders(lib))
    for f in ("kat_reference.npz", "render_reference.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")
