"""create_model — same signature and error behaviour as the reference factory.py:6-21.

    create_model(arch='pipeline', version='paintmindv1', pretrained=True, checkpoint_path=None)

KeyError on an unknown version (ver2cfg lookup), ValueError on an unknown arch.  With
pretrained=True and no checkpoint_path the reference downloads "RootYuan/<version>.pt" from the
HF hub (factory.py:18); this build is offline, so it raises a clear error instead.
"""
from __future__ import annotations

from .config import Config, ver2cfg


def create_model(arch="pipeline", version="paintmindv1", pretrained=True, checkpoint_path=None):
    config = Config(ver2cfg[version])
    if arch == "vqgan":
        from .stage1 import VQModel
        model = VQModel(config)
    elif arch == "pipeline":
        from .generate import Pipeline
        model = Pipeline(config, stage1_pretrained=False)
    else:
        raise ValueError(f"failed to load arch named {arch}")
    if pretrained:
        if checkpoint_path is None:
            try:
                from huggingface_hub import hf_hub_download
                checkpoint_path = hf_hub_download("RootYuan/" + version, f"{version}.pt")
            except Exception as exc:  # offline image: same call as the reference, explicit failure
                raise RuntimeError(
                    f"pretrained=True needs checkpoint_path (hub download of RootYuan/{version} failed: {exc})") from exc
        model.from_pretrained(checkpoint_path)
    return model


def create_pipeline_for_train(version="paintmindv1", stage1_pretrained=True, stage1_checkpoint_path=None):
    from .generate import Pipeline
    return Pipeline(Config(ver2cfg[version]), stage1_pretrained=stage1_pretrained,
                    stage1_checkpoint_path=stage1_checkpoint_path)
