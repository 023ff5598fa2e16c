import sys; sys.path[:0]=['.','tests']
import numpy as np, orc
from pupiloptixlab_b200 import pupil, scenes
for desc in [scenes.cornell_box(64,48,8), scenes.material_grid(64,36,8)]:
    pupil.load_scene(desc, host_only=True)
    o = orc.OracleScene(orc.port(), desc)
    s2c,c2w,fov = pupil.camera(); os2c,oc2w,ofov = o.camera()
    print(desc.name, "cam s2c", np.array_equal(s2c,os2c), "c2w", np.array_equal(c2w,oc2w), fov==ofov)
    ins = pupil.instances()
    print(" n inst", len(ins), "xforms equal", all(np.array_equal(ins[i]['xform'], o.instance_xform(i)) for i in range(len(ins))))
    ar, env = pupil.emitters(); oar = o.area_emitters()
    print(" emitters", len(ar), len(oar), all(bytes(a)==bytes(b) for a,b in zip(ar,oar)), env is not None)
    print(" flags", [i['flags'] for i in ins][:5], "eo", [i['emitter_offset'] for i in ins if i['emitter_offset']>=0])
