import sys; sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, orc
from pupiloptixlab_b200 import pupil, scenes
pupil.init(0)
lib=orc.port()
desc = scenes.cornell_box(64, 64, 6); box=6
fresh = scenes.cornell_box(64, 64, 6)
fresh.shapes[box].to_world = scenes.Xf("srt", scale=(0.25, 0.4, 0.25), rotate_axis=(0, 1, 0), rotate_angle=35.0, translate=(0.2, 0.4, 0.1))
xf = orc.OracleScene(lib, fresh).instance_xform(box)
pupil.load_scene(fresh); pupil.pass_config(frames_per_run=4); pupil.run(1)
want = pupil.buffer("pt accum buffer").copy(); wi=pupil.instances()[box]["xform"].copy()
pupil.load_scene(desc); pupil.pass_config(frames_per_run=4); pupil.run(1)
before = pupil.buffer("pt accum buffer").copy()
pupil.set_instance_transform(box, xf); pupil.run(1)
got = pupil.buffer("pt accum buffer")
print("xf equal", np.array_equal(wi, xf), np.array_equal(pupil.instances()[box]["xform"], xf))
d=np.abs(got-want); print("diff pixels", np.count_nonzero(d.max(-1)>0), "max", d.max(), "before-diff", np.count_nonzero(np.abs(got-before).max(-1)>0))
print(pupil.pass_state())
