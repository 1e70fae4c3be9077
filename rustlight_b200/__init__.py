"""rustlight_b200 -- B200 wavefront path tracer behind rustlight's `path` / `direct` integrators.

Host layer (scene loading, camera) lives in librl_host.so; the device library is librl_b200.so
(C ABI: include/rl_b200.h).  Importing this package does not require a GPU; creating a
`Context` does, and fails loudly when the CUDA library is missing.
"""
from . import _abi  # noqa: F401
from .host import Scene, SceneLoaderManager, SceneError  # noqa: F401
