"""B200-native camera render pass for przemo199/ray-tracer-challenge-rs.

The package is the host side of ONE hot path — ``Camera::render`` and everything under it
(``ray-tracer/src/composites/camera.rs:79-112``) — re-implemented as hand-written sm_100a CUDA
kernels behind the C ABI of ``include/rtgpu.h``:

  * :mod:`.scene`         the reference's World / Camera / Canvas / shapes / materials / patterns
  * :mod:`.scene_loader`  the YAML scene loader (``ray-tracer-cli/src/scene_loader.rs``)
  * :mod:`.flatten`       scene flattener: trait objects -> SoA buffers
  * :mod:`.render`        ``Camera.render_gpu`` -> ``librtgpu.so`` (no CPU fallback)
  * :mod:`.abi`           ctypes declarations of the C ABI
  * ``csrc/``             the CUDA kernels and the C-ABI implementation
"""
from . import abi, primitives  # noqa: F401
from .flatten import FlatScene, camera_to_c, flatten_world  # noqa: F401
from .scene import (  # noqa: F401
    Camera,
    Canvas,
    CheckerPattern,
    ComplexPattern,
    Cone,
    Cube,
    Cylinder,
    GradientPattern,
    Light,
    Material,
    Plane,
    RingPattern,
    Sphere,
    StripePattern,
    TestPattern,
    Triangle,
    World,
)
from .scene_loader import load_scene_description, load_scene_from_str  # noqa: F401

__all__ = [
    "abi", "primitives", "FlatScene", "camera_to_c", "flatten_world", "Camera", "Canvas", "CheckerPattern",
    "ComplexPattern", "Cone", "Cube", "Cylinder", "GradientPattern", "Light", "Material", "Plane", "RingPattern",
    "Sphere", "StripePattern", "TestPattern", "Triangle", "World", "load_scene_description", "load_scene_from_str",
]
