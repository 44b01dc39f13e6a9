"""cobbletrace_b200 -- B200-native (sm_100a CUDA) renderer for CobbleTrace's ray/scene intersection +
shading path, behind the C ABI of include/ct_gpu.h.  CUDA-only: there is no CPU fallback."""
from .sceneio import FlatScene, load_ctscene, save_ctscene, frame_fnv1a  # noqa: F401
from .api import GpuRenderer, CtError, load_library, device_count  # noqa: F401
from .api import CT_FLAG_WIDE, CT_FLAG_KEEP_HITS, CT_FLAG_COUNT_TESTS, CT_FLAG_STAGE_TIMING, CT_FLAG_SUBSAMPLING, CT_FLAG_SUPERSAMPLING, BACKGROUND, REFERENCE_MAX_DEPTH  # noqa: F401

__version__ = "0.1.0"
