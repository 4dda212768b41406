"""real-time-opencl-raytracer_b200 -- B200-native hot path (BVH traversal + ray/triangle
intersection) of itmanager85/real-time-opencl-raytracer behind a C ABI (include/rtb200.h).

  csrc/    hand-written sm_100a CUDA kernels + the extern "C" boundary   (librtb200.so)
  host/    C++ host scene layer mirroring the reference's API surface     (librtb200_host.so)
  device.py / hostlib.py   thin ctypes views of the two libraries for tests and bench.py

The directory name contains hyphens, so import it through the `rtb200` shim at the repo root."""
from . import device, hostlib, tiling  # noqa: F401
from .device import ANY, CLOSEST, HIT_DTYPE, RAY_DTYPE, T_INIT, Context, Group, RtError, pack_scene_host  # noqa: F401
from .hostlib import FlatBVH, Mesh, camera_params  # noqa: F401
