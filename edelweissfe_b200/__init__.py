"""edelweissfe_b200 — B200-native element-loop / CSR-assembly path for EdelweissFE
(hand-written CUDA sm_100a behind a C ABI; PyTorch only owns the buffers)."""
from .assembly import CutbackRequest, ElementAssembly  # noqa: F401
from .boxgen import box_mesh  # noqa: F401

__version__ = "0.1.0"
