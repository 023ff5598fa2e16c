"""Small Python-side helpers (no compute): the built-in shapes' object-space data, needed by callers that
feed pb2 directly.  Vertex order follows framework/resource/shape.cpp:21-68 of the reference because
primitive ids and emitter indices depend on it."""
import numpy as np


def builtin_mesh(kind: str) -> dict:
    if kind == "rectangle":
        pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
        nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
        uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
        idx = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
        return dict(positions=pos, normals=nrm, texcoords=uv, indices=idx)
    if kind == "cube":
        pos = np.array([-1, -1, -1, -1, -1, 1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, -1, 1, -1, 1, 1, -1,
                        1, -1, 1, 1, -1, -1, 1, 1, -1, 1, 1, 1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1, 1, 1,
                        -1, 1, 1, 1, 1, 1, 1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, 1, -1, 1, -1, -1, 1], np.float32).reshape(-1, 3)
        fn = np.array([[-1, 0, 0], [0, 0, -1], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, -1, 0]], np.float32)
        nrm = np.repeat(fn, 4, axis=0)
        uv = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), (6, 1))
        idx = np.concatenate([[[4 * f, 4 * f + 1, 4 * f + 2], [4 * f, 4 * f + 2, 4 * f + 3]] for f in range(6)]).astype(np.uint32)
        return dict(positions=pos, normals=nrm, texcoords=uv, indices=idx)
    raise ValueError(kind)
