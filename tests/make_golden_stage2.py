"""Stage-2 (MaskGIT) golden fixtures from the unmodified reference; see make_golden.py."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

GOLD = ROOT / "tests" / "golden"

from stage2_inputs import TINY2, full_step_inputs, tiny_inputs  # noqa: E402


def stage2_fixtures(pm):
    from paintmind.stage2 import CondTransformer
    cfg1 = ver2cfg["vit-tiny-test"]
    out = {}
    for name, ctx_dim in (("same", 128), ("proj", 96)):
        sd = synthetic.make_stage2_state_dict(TINY2, cfg1, seed=5, context_dim=ctx_dim)
        tr = CondTransformer(32, TINY2["dim"], 64, TINY2["dim_head"], TINY2["mlp_dim"], TINY2["num_head"], TINY2["depth"],
                             TINY2["dropout"], ctx_dim, cfg1["n_embed"]).eval()
        tsd = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
        tr.load_state_dict(tsd, strict=True)
        tokens, context = tiny_inputs(ctx_dim)
        with torch.no_grad():
            out[f"logits_{name}"] = tr(tokens, context).numpy().astype(np.float32)
            if name == "same":
                out["logits_nocontext"] = tr(tokens, None).numpy().astype(np.float32)
    np.savez_compressed(GOLD / "stage2_tiny.npz", **out)
    print("stage2_tiny:", {k: v.shape for k, v in out.items()})

    step_fixture(pm, "paintmindv1", "stage2_step.npz")
    # the 256-token variant (image_size 128, registered in THIS repo's config.py only): the reference classes take every size from
    # the config, so its registry gets the two entries for the duration of this script (in memory; /root/reference is untouched)
    import paintmind.config as RC
    RC.ver2cfg["vit-s-vqgan-128"] = ver2cfg["vit-s-vqgan-128"]
    RC.ver2cfg["paintmindv1-128"] = ver2cfg["paintmindv1-128"]
    step_fixture(pm, "paintmindv1-128", "stage2_step_256.npz")


def step_fixture(pm, version, out_name):
    """One MaskGIT step of the reference Pipeline with injected noise (generate.py:159-181)."""
    import paintmind.generate as G
    cfg2 = ver2cfg[version]
    cfg1 = ver2cfg[cfg2["stage1"]]
    pipe = pm.create_model(arch="pipeline", version=version, pretrained=False).eval()
    sd = {("vqgan." + k): v for k, v in synthetic.make_vqgan_state_dict(cfg1, seed=0).items()}
    sd.update(synthetic.make_stage2_state_dict(cfg2, cfg1, seed=1, context_dim=1024))
    res = pipe.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    N, V = (cfg1["enc"]["image_size"] // cfg1["enc"]["patch_size"]) ** 2, cfg1["n_embed"]
    text, ids, u = full_step_inputs(N=N)
    orig = G.gumbel_noise
    G.gumbel_noise = lambda t: -G.log(-G.log(u))
    try:
        with torch.no_grad():
            tokens = pipe.ids2tokens(ids)
            logits = pipe.tokens2logits(tokens, text)
            mask_ratio = G.mask_schedule(8 / 12)
            new_ids, img = pipe.sample(ids, mask_ratio, text=text, topk=5, temperature=0.75)
            filtered = G.top_k(logits, 5)
            pred_ids = G.gumbel_sample(filtered, temperature=0.75, dim=-1)
            probs = logits.softmax(-1)
            scores = (1 - probs.gather(2, pred_ids[..., None]))[..., 0].masked_fill(ids != V, -1e5)
    finally:
        G.gumbel_noise = orig
    top6 = logits.topk(6, dim=-1)
    k = max(int((mask_ratio * N).item()), 1)
    np.savez_compressed(
        GOLD / out_name,
        ids_in=ids.numpy().astype(np.int16), mask_ratio=float(mask_ratio), k=k,
        tokens_head=tokens[0, :8].numpy().astype(np.float32),
        logits_sub=logits[0, ::16, ::16].numpy().astype(np.float32),
        lse=torch.logsumexp(logits, -1)[0].numpy().astype(np.float32),
        top6_val=top6.values[0].numpy().astype(np.float32), top6_idx=top6.indices[0].numpy().astype(np.int16),
        pred_ids=pred_ids[0].numpy().astype(np.int16), scores=scores[0].numpy().astype(np.float32),
        new_ids=new_ids[0].numpy().astype(np.int16),
        img_sub=img[0, :, ::4, ::4].numpy().astype(np.float32),
        text_sum=float(text.double().sum()), u_sum=float(u.double().sum()),
    )
    print(f"{out_name}: N={N} k={k} masked_in={(ids == V).sum().item()} masked_out={(new_ids == V).sum().item()} "
          f"logits absmax={logits.abs().max():.3f}")
