"""paintmind_b200 — B200-native (sm_100a) implementation of PaintMind's image-tokenizer hot path.

Public surface mirrors the reference's paintmind/__init__.py for the path in scope:
    import paintmind_b200 as pm
    model = pm.create_model(arch='vqgan', version='vit-s-vqgan', pretrained=False).cuda().eval()
    z, loss, idx = model.encode(x); rec = model.decode(z)
"""
from .version import __version__  # noqa: F401
from .config import Config, ver2cfg  # noqa: F401
from .factory import create_model, create_pipeline_for_train  # noqa: F401
