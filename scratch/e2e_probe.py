import sys, time; sys.path[:0]=['.']
import numpy as np, torch
from pupiloptixlab_b200 import pupil, scenes
pupil.init(0)
desc = scenes.cornell_box(1920,1080,8)
def T(label, f):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); print(f"{label:30s} {1e3*(time.perf_counter()-t0):8.2f} ms"); return r
for it in range(3):
    print("iter", it)
    T("load_scene", lambda: pupil.load_scene(desc))
    T("scene_handle", lambda: pupil.scene_handle())
    T("pass_config", lambda: pupil.pass_config(frames_per_run=64))
    T("run", lambda: pupil.run(1))
    img = T("buffer download", lambda: pupil.buffer("final result"))
    T("isfinite", lambda: np.isfinite(img[..., :3]).all())
