from .vqmodel import VQModel  # noqa: F401
from .quantize import VectorQuantizer  # noqa: F401
