import sys, time; sys.path[:0]=['.','tests']
import numpy as np, orc
from pupiloptixlab_b200 import pupil, scenes
pupil.init(0)
def stats(name, g, r):
    g=g.reshape(-1,g.shape[-1])[:, :3].astype(np.float64); r=r.reshape(-1,r.shape[-1])[:, :3].astype(np.float64)
    d=np.abs(g-r).max(1); tol=1e-4*np.maximum(1.0,np.abs(r).max(1))
    exact=(g.astype(np.float32)==r.astype(np.float32)).all(1)
    print(f"  {name}: bit-exact {exact.mean()*100:.2f}%  within1e-4 {(d<=tol).mean()*100:.3f}%  maxdiff {d.max():.4g}  mean gpu {g.mean():.6f} ref {r.mean():.6f}  nan gpu {np.isnan(g).sum()} ref {np.isnan(r).sum()}")
for desc, frames in [(scenes.cornell_box(128,128,8),1),(scenes.material_grid(160,90,8),1),(scenes.terrain(40,128,72,8),1),(scenes.cornell_box(64,64,8),8),(scenes.material_grid(96,54,8),8)]:
    t0=time.time(); pupil.load_scene(desc); pupil.pass_config(frames_per_run=1); pupil.run(frames); t1=time.time()
    o=orc.OracleScene(orc.port(),desc); ref=o.render(frames); t2=time.time()
    rs=pupil.render_stats()
    print(desc.name, desc.sensor.width, desc.sensor.height, "frames",frames, f"gpu {t1-t0:.2f}s cpu {t2-t1:.2f}s", "rays gpu", rs.closest_rays, rs.shadow_rays, "ref", ref["closest_rays"], ref["shadow_rays"])
    stats("accum", pupil.buffer("pt accum buffer"), ref["accum"])
    stats("frame", pupil.buffer("final result"), ref["frame"])
    stats("albedo", pupil.buffer("albedo"), ref["albedo"])
    stats("normal", pupil.buffer("normal"), ref["normal"])
    t=pupil.buffer("test").reshape(-1); print("  test exact", (t==ref["test"]).mean())
    # batch mode equality
    a1=pupil.buffer("pt accum buffer").copy()
    pupil.pass_config(frames_per_run=frames); pupil.run(1)
    a2=pupil.buffer("pt accum buffer")
    print("  batched == sequential:", np.array_equal(a1,a2,equal_nan=True))
