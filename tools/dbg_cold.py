import sys, time
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np
from pupiloptixlab_b200 import pupil, scenes
pif = int(sys.argv[1])
pupil.init(0)
d = scenes.cornell_box(1920, 1080, 8)
pupil.load_scene(d)
if pif: pupil.scene_handle().set_option("paths_in_flight", pif)
pupil.pass_config(frames_per_run=64); pupil.run(1); print('rendered', pupil.render_stats().total_ms)
t = scenes.terrain(3873, 1920, 1080, 8)
t0 = time.perf_counter(); pupil.load_scene(t); print('load s', time.perf_counter() - t0, 'build ms', pupil.build_stats().build_ms)
for _ in range(2):
    pupil.set_bvh_builder(0); print('rebuild ms', pupil.build_stats().build_ms)
pupil.shutdown()
