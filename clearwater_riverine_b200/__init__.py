"""B200-native implementation of ClearWater-Riverine's per-timestep implicit
advection-diffusion step (reference: src/clearwater_riverine/transport.py:201-276 and
linalg.py) behind the reference's Python stepping API.  CUDA only -- no CPU fallback."""
from .backend import CwrError, SolverWarning, TransportBackend, load_library  # noqa: F401
from .transport import ClearwaterRiverine, Constituent, ModelMesh  # noqa: F401
from .adapter import attach, extract_model_arrays  # noqa: F401

__version__ = "0.1.0"
