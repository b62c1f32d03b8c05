from .attention import CrossAttention, MemoryEfficientCrossAttention  # noqa: F401
from .mlp import SwiGLUFFN, SwiGLUFFNFused  # noqa: F401
