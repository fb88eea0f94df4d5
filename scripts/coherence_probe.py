"""How much cheaper is a coherent wavefront?  Primary rays only (max_depth 1) vs the full mix."""
import sys
sys.path.insert(0, ".")
from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene

for depth in (1, 2, 50):
    sc = demo_builder.book1_end_scene(image_width=1920, samples=32)
    sc.scene_cam.set_max_depth(depth)
    gs = GpuScene(sc.describe(), 0)
    cam = sc.scene_cam.to_abi()
    for _ in range(2):
        gs.render(cam, seed=1, want_rgb=False, want_rgb8=False)
    _, _, st = gs.render(cam, seed=1, time_kernels=True, want_rgb=False, want_rgb8=False)
    print(f"depth {depth}: rays {st['rays']/1e6:.1f}M trace {st['ms_trace']:.2f} ms -> {st['ms_trace']*1e6/st['rays']:.3f} ns/ray; shade {st['ms_shade']:.2f} gen {st['ms_raygen']:.2f} total {st['ms_total']:.2f}")
