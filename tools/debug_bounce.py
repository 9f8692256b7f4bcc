"""GPU-side debugging aid: find pixels whose radiance differs from the oracle and trace their rays."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import ctx_for
from oracle import pyoracle
from scenes import camera, shading, terrain
from voxelrt_b200 import capi

scene = terrain.terrain_hash(6, 4, 6, seed=77)
ctx = ctx_for(scene)
orc = pyoracle.OracleMap(6, 4); orc.set_palette(scene["palette"]); orc.sync(terrain.scene_records(scene))
bn, _ = shading.load_blue_noise(); desc, tex, _ = shading.load_sky()
ctx.set_blue_noise(bn); ctx.set_sky(desc, tex); orc.set_blue_noise(bn); orc.set_sky(desc, tex)
cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
w, h = 320, 180
for bounces in (1, 2, 3):
    for frame_no in (1, 2, 77):
        proj, inv, wo, frac = cam.matrices(w, h)
        mk = lambda: capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=bounces, flags=capi.VRT_FRAME_LINEAR_OUTPUT)
        g, _ = ctx.render(mk()); c, _, _ = orc.render(mk())
        bad = np.argwhere((g != c).any(axis=0))
        print(f"bounces {bounces} frame {frame_no}: {len(bad)} differing pixels")
        for (y, x) in bad[:4]:
            print(" pixel", x, y, "gpu", [hex(v) for v in g[:, y, x]], "cpu", [hex(v) for v in c[:, y, x]])
            rays, hits, out4 = orc.debug_pixel(mk(), int(x), int(y))
            gh = ctx.trace(rays[:, :3], rays[:, 3:], wo)
            for i in range(len(rays)):
                print("  ray", i, rays[i].tolist(), [hex(v) for v in rays[i].view(np.uint32)])
                print("   cpu", hits[i]); print("   gpu", gh[i])
