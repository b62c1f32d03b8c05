"""Import the real (unmodified) reference.

Lookup order: $PAINTMIND_REF, `baseline/_ref` (the offline `pip install --no-deps --target baseline/_ref
/root/reference` of DESIGN.md §5: git-ignored, travels to the GPU box with the snapshot), /root/reference (the read-only
mount of the build container).  Used by tests/make_golden*.py to generate fixtures and by bench.py's CPU arms
(`--impl reference`, `cpu_baseline`) as the thing being TIMED; nothing on the product path imports it.
Two off-path reference modules (T5/CLIP embedders and the accelerate-based trainers) pull in packages that are absent
here, so they are pre-registered as stubs in sys.modules; every on-path module is the reference's own code.
"""
from __future__ import annotations

import os
import sys
import types


def find_reference():
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("PAINTMIND_REF"), os.path.join(repo_root, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "paintmind")):
            return cand
    return None


def load_reference():
    ref = find_reference()
    if ref is None:
        raise RuntimeError("reference not mounted")
    if "paintmind" in sys.modules and getattr(sys.modules["paintmind"], "__graft_ref__", False):
        return sys.modules["paintmind"]
    sys.dont_write_bytecode = True
    import torch
    from torch import nn

    enc = types.ModuleType("paintmind.modules.encoder")

    class T5TextEmbedder(nn.Module):   # stand-in for the frozen T5 (off path; BASELINE config 5 uses random embeddings)
        def __init__(self, version=None, freeze=True, **kw):
            super().__init__()

        def forward(self, text):
            g = torch.Generator().manual_seed(1234)
            return torch.randn(len(text), 77, 1024, generator=g)

    enc.T5TextEmbedder = T5TextEmbedder
    tr = types.ModuleType("paintmind.utils.trainer")
    tr.VQGANTrainer = None
    tr.PaintMindTrainer = None
    sys.modules["paintmind.modules.encoder"] = enc
    sys.modules["paintmind.utils.trainer"] = tr
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import paintmind as pm
    pm.__graft_ref__ = True
    return pm
