import sys, os
sys.path.insert(0, '/root/repo')
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture
from ray_tracer_challenge_rs_b200.render import Renderer
flat, camera = load_scene_fixture("cover")
for w, h in ((320, 180), (1920, 1080)):
    with Renderer(flat) as r:
        n0 = r.launch_count()
        r.render(camera.resized(w, h), family="wavefront", want_rgb8=False)
        print(w, h, "launches of the first frame:", r.launch_count() - n0, flush=True)
