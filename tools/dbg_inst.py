import sys
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np
from gpu_util import random_rays, random_soup
from pupiloptixlab_b200 import pb2
pb2.init(0)
rays = random_rays(20000, 9, extent=45.0)
for k_inst, n_tris in ((1000, 2000),):
    mesh = random_soup(n_tris, 4, extent=1.0, size=0.08)
    def make(mode):
        s = pb2.Scene(); s.set_option("instancing", mode)
        mid = s.add_mesh(mesh["positions"], mesh["indices"])
        r = np.random.default_rng(11)
        for k in range(k_inst):
            a = r.uniform(0, 2 * np.pi); c, sn = np.cos(a), np.sin(a); sc = r.uniform(0.5, 1.5)
            m = np.array([[c * sc, 0, sn * sc, 0], [0, sc, 0, 0], [-sn * sc, 0, c * sc, 0]], np.float32)
            m[:, 3] = r.uniform(-40, 40, 3)
            s.add_instance(mid, m)
        return s, s.build()
    s1, st1 = make(1); s1.set_option("coop_prims", 1); h1 = s1.trace_closest(rays)
    s0, st0 = make(0); h0 = s0.trace_closest(rays)
    same = (h1["inst"] == h0["inst"]) & (h1["prim"] == h0["prim"])
    never = set(np.unique(h0["inst"][h0["inst"] >= 0])) - set(np.unique(h1["inst"][h1["inst"] >= 0]))
    print(k_inst, n_tris, 'depth', st1.max_depth, 'nodes', st1.n_nodes, 'differ', np.count_nonzero(~same), 'never hit', len(never), sorted(never)[:10])
