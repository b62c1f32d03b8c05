"""Pipeline — the MaskGIT sampling half of the reference's generate.py (:25-46, :49-76, :125-134,
:148-198) on the CUDA kernels.  Same constructor, attributes (vqgan, text_model, transformer,
mask_token, mask_token_id, image_size, patch_size, num_tokens) and method signatures.

The stage-2 training FORWARD (random_masking / loss / forward, generate.py:78-146)
is also here as forward values (no autograd graph).

Out of scope: the T5 text encoder (frozen third-party model, needs network
weights — `text_model` is a pluggable callable, by default seeded random embeddings as in BASELINE
config 5), backward passes, and inpaint / outpaint (broken as shipped in the reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from . import ops
from .config import ver2cfg


def mask_schedule(ratio):
    """cos(pi/2 * ratio) in numpy float64 (generate.py:25-26)."""
    return np.cos(math.pi / 2.0 * ratio)


class RandomTextEmbedder(nn.Module):
    """Stand-in for the frozen T5 embedder (modules/encoder.py:18-42, out of scope): seeded random
    [B, 77, dim] embeddings, the input BASELINE config 5 prescribes.  Replace `Pipeline.text_model`
    with any callable list[str] -> [B, L, context_dim] tensor to plug a real encoder in."""

    def __init__(self, dim=1024, length=77, seed=1234):
        super().__init__()
        self.dim, self.length, self.seed = dim, length, seed

    def forward(self, text):
        g = torch.Generator().manual_seed(self.seed)
        return torch.randn(len(text), self.length, self.dim, generator=g)


class Pipeline(nn.Module):
    def __init__(self, config, stage1_pretrained=True, stage1_checkpoint_path=None):
        super().__init__()
        from .factory import create_model
        from .stage2 import CondTransformer
        t5_txt_dim = {"t5-l": 1024, "t5-xl": 2048}
        self.vqgan = create_model(arch="vqgan", version=config.stage1, pretrained=stage1_pretrained,
                                  checkpoint_path=stage1_checkpoint_path)
        self.vqgan.freeze()
        self.text_model = RandomTextEmbedder(t5_txt_dim[config.t5])
        vq_cfg = ver2cfg[config.stage1]
        self.image_size = vq_cfg["enc"]["image_size"]
        self.patch_size = vq_cfg["enc"]["patch_size"]
        self.num_tokens = (self.image_size // self.patch_size) ** 2
        self.transformer = CondTransformer(
            vq_cfg["embed_dim"], config.dim, self.num_tokens, config.dim_head, config.mlp_dim,
            config.num_head, config.depth, config.dropout, t5_txt_dim[config.t5], vq_cfg["n_embed"])
        self.mask_token = nn.Parameter(torch.zeros(1, vq_cfg["embed_dim"]))
        self.mask_token_id = vq_cfg["n_embed"]
        nn.init.normal_(self.mask_token, std=0.02)
        self._rng_seed = None
        self._rng_calls = 0
        self.cuda_graph = None        # None: replay the transformer forward from a CUDA graph when the batch is small; True / False force it
        self._table = None
        self._table_fp = None

    def invalidate(self):
        """Drop every cache keyed on parameter versions (packed weights, token table, CUDA graphs): call after writes that
        bypass autograd's version counter (`p.data.copy_(...)`, EMA, weight surgery)."""
        self._table_fp = None
        self.__dict__.pop("_step_graph_rec", None)
        self.vqgan.invalidate()
        eng = self.transformer.__dict__.get("_engine")
        if eng is not None:
            eng.invalidate()

    def _advance_rng(self):
        """Philox key / counter of the next sampling call, drawn from torch's default generator: (torch.initial_seed(), one
        31-bit draw).  Results are therefore a function of the torch RNG state — `torch.manual_seed(s)` before generate() /
        random_masking() reproduces a run, as with the reference (whose own random stream differs: it draws 8192 uniforms
        per token on the device)."""
        self._rng_seed = torch.initial_seed()
        self._rng_calls = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())

    def from_pretrained(self, path):
        """generate.py:73-75.  A reference Pipeline checkpoint also carries the frozen T5 encoder (`text_model.transformer.*`,
        a registered submodule there).  The text encoder is out of scope here (`text_model` is a pluggable callable, by
        default parameter-less): its keys are dropped unless the attached text model declares parameters of that name;
        every other key is checked strictly."""
        sd = torch.load(path, map_location="cpu")
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in sd.items() if not (k.startswith("text_model.") and k not in own)}
        return self.load_state_dict(sd, strict=True)

    # ---- stage-2 training FORWARD (generate.py:78-146) -------------------
    # Forward values only: like every other entry point of this package the result carries no autograd graph
    # (the backward pass of the transformer is out of scope).
    @torch.no_grad()
    @ops.on_device_of
    def random_masking(self, x, mask_ratio, _noise=None):
        """generate.py:78-108 -> (x with masked rows replaced by mask_token, mask [N, L] fp32 with 1 = masked).
        `_noise` injects the [N, L] uniforms the reference draws with torch.rand (parity tests); production
        draws Philox4x32-10 keyed on (torch.initial_seed(), sample, token, call counter)."""
        if not x.is_cuda:
            raise RuntimeError("paintmind_b200: CUDA only (no CPU fallback)")
        N, L, D = x.shape
        if D != 32:
            raise RuntimeError("paintmind_b200 token kernels are built for 32-d tokens")
        len_mask = max(int(L * mask_ratio), 1)
        len_keep = L - len_mask
        x2d = x.detach().reshape(N * L, D)
        if x2d.dtype != torch.float32:
            x2d = x2d.float()
        if x2d.stride(1) != 1 or x2d.stride(0) % 4 != 0 or x2d.data_ptr() % 16 != 0:
            x2d = x2d.contiguous()
        mask = torch.empty(N, L, device=x.device, dtype=torch.float32)
        out = torch.empty(N * L, D, device=x.device, dtype=torch.float32)
        self._advance_rng()
        noise = None
        if _noise is not None:
            noise = _noise.to(device=x.device, dtype=torch.float32).contiguous()
            if noise.shape != (N, L):
                raise RuntimeError(f"_noise must be [{N}, {L}]")
        ops.maskgit_random_mask(x2d, self.mask_token.detach().float().contiguous().view(-1), N, L, len_keep,
                                mask=mask, x_out=out, noise=noise, seed=self._rng_seed, offset=self._rng_calls)
        return out.view(N, L, D), mask

    @torch.no_grad()
    @ops.on_device_of
    def loss(self, logit, label, masks):
        """generate.py:110-123: label-smoothed (0.1) cross entropy averaged over the masked positions."""
        if not logit.is_cuda:
            raise RuntimeError("paintmind_b200: CUDA only (no CPU fallback)")
        B, L, V = logit.shape
        lg = logit.detach().reshape(B * L, V)
        if lg.dtype != torch.float32:
            lg = lg.float()
        if lg.stride(1) != 1 or lg.stride(0) % 4 != 0 or lg.data_ptr() % 16 != 0:
            lg = lg.contiguous()
        lab = label.reshape(-1).to(torch.int64).contiguous()
        mk = masks.reshape(-1).to(torch.float32).contiguous()
        row_loss = torch.empty(B * L, device=logit.device, dtype=torch.float32)
        out = torch.empty((), device=logit.device, dtype=torch.float32)
        sums = torch.empty(2, device=logit.device, dtype=torch.float64)
        ops.ce_label_smooth(lg, lab, mk, 0.1, row_loss=row_loss, loss_out=out, sums_out=sums)
        self._last_loss_sums = sums          # (sum of masked row losses, mask count): all-reduce these across ranks
        return out

    @torch.no_grad()
    def forward(self, img, text=None, mask_ratio=0.75, _noise=None):
        """generate.py:136-146: tokenize (frozen VQGAN) -> random masking -> transformer -> masked CE."""
        x, ids, text = self.to_latent(img, text)
        x, mask = self.random_masking(x, mask_ratio, _noise=_noise)
        logits = self.tokens2logits(x, text)
        return self.loss(logits, ids, mask)

    def inpaint(self, *a, **k):
        raise NotImplementedError("inpaint is broken as shipped in the reference (float ids, generate.py:206-210) and out of scope")

    outpaint = inpaint

    # ---- hot path ---------------------------------------------------------------------------
    @torch.no_grad()
    def to_latent(self, img, text=None):
        x, _, indices = self.vqgan.encode(img)
        if text is not None:
            text = self._embed_text(text, img.device)
        return x, indices, text

    def _embed_text(self, text, device):
        if torch.is_tensor(text):
            return text.to(device)
        return self.text_model(text).to(device)

    @torch.no_grad()
    def tokens2logits(self, token, text=None):
        return self.transformer(token, text)

    def _token_table(self):
        """cat(raw codebook, mask_token): [n_embed + 1, 32] — un-normalised, as the reference (generate.py:149-154)."""
        w = self.vqgan.quantize.embedding.weight
        fp = (w.data_ptr(), w._version, self.mask_token.data_ptr(), self.mask_token._version)
        if fp != self._table_fp:
            self._table = torch.cat((w.detach().float(), self.mask_token.detach().float())).contiguous()
            self._table_fp = fp
        return self._table

    @torch.no_grad()
    @ops.on_device_of
    def ids2tokens(self, ids):
        table = self._token_table()
        flat = ids.reshape(-1).to(torch.int64).contiguous()
        out = torch.empty(flat.numel(), table.shape[1], device=ids.device, dtype=torch.float32)
        ops.vq_gather(flat, table, False, out, None)
        return out.reshape(*ids.shape, table.shape[1])

    @torch.no_grad()
    @ops.on_device_of
    def sample(self, ids, mask_ratio, text=None, topk=1, temperature=1, decode=True, _noise=None):
        """One MaskGIT step (generate.py:159-181) -> (ids, img).  `decode=False` skips the per-step
        ViT decode (the reference decodes every step even when generate() discards the image, F9)."""
        if not ids.is_cuda:
            raise RuntimeError("paintmind_b200: CUDA only (no CPU fallback)")
        B, N = ids.shape
        dev = ids.device
        table = self._token_table()
        eng = self.transformer.engine()
        ids = ids.to(torch.int64).contiguous().clone()
        context = self._embed_text(text, dev) if text is not None else None
        # small batches are launch-bound on the host: replay the transformer forward from a CUDA graph (cuda_graph = None: auto)
        use_graph = self.cuda_graph if self.cuda_graph is not None else (B * N <= 16 * 1024)
        fwd = eng.forward_from_ids_graphed if use_graph else eng.forward_from_ids
        logits = fwd(ids, table, context)                                       # fp32 [B, N, V]
        pred_ids = torch.empty(B, N, device=dev, dtype=torch.int64)
        scores = torch.empty(B, N, device=dev, dtype=torch.float32)
        self._advance_rng()
        ops.maskgit_sample(logits.view(B * N, -1), topk=topk, temperature=temperature, ids=ids.view(-1),
                           pred_ids=pred_ids.view(-1), scores=scores.view(-1), mask_id=self.mask_token_id,
                           noise=None if _noise is None else _noise.reshape(B * N, -1),
                           seed=self._rng_seed, offset=self._rng_calls)
        img = self.vqgan.decode_from_indice(pred_ids) if decode else None
        r = mask_ratio.item() if hasattr(mask_ratio, "item") else mask_ratio
        num_token_masked = max(int(r * self.num_tokens), 1)
        ops.maskgit_remask(scores, ids, num_token_masked, self.mask_token_id)
        self._last_pred_ids, self._last_scores, self._last_logits = pred_ids, scores, logits
        return ids, img

    # ---- one CUDA graph for a whole MaskGIT step (small batches are launch-bound on the host) ----
    def _step_graph(self, B, context, topk):
        """Captured sequence ids -> tokens -> transformer -> top-k / gumbel sample -> re-mask, operating in place on static
        buffers.  The per-step scalars (temperature, re-mask count, Philox key) are read on the device from a table indexed by
        a device-side step counter that the last kernel advances, so ONE graph serves every step of every generate() call
        with this (batch, context tensor, topk, weights)."""
        dev = self.mask_token.device
        table = self._token_table()
        eng = self.transformer.engine()
        eng._ensure_packed()
        key = (B, topk, table.data_ptr(), table._version, context.data_ptr(), context._version, tuple(context.shape), eng._fp)
        g = self.__dict__.get("_step_graph_rec")
        if g is not None and g["key"] == key and g["context"] is context and g["table"] is table:
            return g
        N = self.num_tokens
        st = dict(key=key, context=context, table=table,
                  ids=torch.full((B, N), self.mask_token_id, dtype=torch.int64, device=dev),
                  pred_ids=torch.empty(B, N, dtype=torch.int64, device=dev),
                  scores=torch.empty(B, N, dtype=torch.float32, device=dev),
                  step_tab=torch.zeros(64 * 24, dtype=torch.uint8, device=dev),     # up to 64 steps of pm_step_scalars
                  step_idx=torch.zeros(1, dtype=torch.int32, device=dev),
                  ticket=torch.zeros(1, dtype=torch.int32, device=dev))

        def body():
            logits = eng.forward_from_ids(st["ids"], table, context)
            ops.maskgit_sample(logits.view(B * N, -1), topk=topk, temperature=1.0, ids=st["ids"].view(-1),
                               pred_ids=st["pred_ids"].view(-1), scores=st["scores"].view(-1), mask_id=self.mask_token_id,
                               step_tab=st["step_tab"], step_idx=st["step_idx"])
            ops.maskgit_remask_step(st["scores"], st["ids"], self.mask_token_id, st["step_tab"], st["step_idx"], st["ticket"])
            return logits

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):                       # sizes the workspace, fills the context K/V cache
                body()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["logits"] = body()
        st["graph"] = graph
        # raw addresses are baked in: pin everything the captured kernels touch (see engine._graphed_forward_from_ids)
        st["pinned"] = list(eng.ws.bufs.values()) + list(eng._ctx_kv or [])
        self.__dict__["_step_graph_rec"] = st
        return st

    @torch.no_grad()
    def generate(self, text, timesteps=18, temperature=1.0, topk=5, save_interval=2, decode_every_step=False):
        """MaskGIT iterative decoding (generate.py:183-198) -> list of CPU image tensors (one per kept step).

        Two departures from the reference's loop, neither changes a returned value:
          * the ViT decoder only runs on the steps whose image is kept (`step % save_interval == 0`); the reference decodes
            every step and drops the rest (generate.py:193-196).  `decode_every_step=True` restores that (parity tests);
          * kept images are staged on the device and leave through pinned host buffers on a side stream, with ONE
            synchronisation after the last step; the reference's `img.cpu()` inside the loop (generate.py:196) blocks the
            host, and with it the launch-bound small-batch loop, once per kept step."""
        dev = self.mask_token.device
        B = len(text)
        context = self._embed_text(text, dev)
        ids = torch.full((B, self.num_tokens), self.mask_token_id, dtype=torch.long, device=dev)
        main = torch.cuda.current_stream(dev)
        side = self.__dict__.get("_d2h_stream")
        if side is None or side.device != dev:
            side = torch.cuda.Stream(device=dev)
            self.__dict__["_d2h_stream"] = side
        imgs = []
        use_graph = self.cuda_graph if self.cuda_graph is not None else (B * self.num_tokens <= 16 * 1024)
        use_graph = use_graph and 1 <= timesteps <= 64 and 1 <= topk <= 32 and context.is_cuda
        if use_graph:
            with torch.cuda.device(dev):
                g = self._step_graph(B, context, topk)
                # the whole schedule goes to the device once: temperature * (1 - step / T), re-mask counts, one noise key per step
                temps, ks, offs = [], [], []
                for step in range(timesteps):
                    self._advance_rng()
                    temps.append(temperature * (1 - step / timesteps))
                    ks.append(max(int(mask_schedule((step + 1) / timesteps) * self.num_tokens), 1))
                    offs.append(self._rng_calls)
                tab = ops.step_table(temps, ks, self._rng_seed, offs, dev)
                g["step_tab"][:tab.numel()].copy_(tab, non_blocking=True)
                g["step_idx"].zero_()
                g["ids"].fill_(self.mask_token_id)
        for step in range(timesteps):
            progress = (step + 1) / timesteps
            masked_r = mask_schedule(progress)
            cur_temp = temperature * (1 - step / timesteps)
            keep = (step % save_interval == 0)
            if ops.NVTX:
                if step:
                    torch.cuda.nvtx.range_pop()
                torch.cuda.nvtx.range_push(f"pm.maskgit_step{step}")
            if use_graph:
                g["graph"].replay()
                img = self.vqgan.decode_from_indice(g["pred_ids"]) if (keep or decode_every_step) else None
                self._last_pred_ids, self._last_scores, self._last_logits = g["pred_ids"], g["scores"], g["logits"]
            else:
                ids, img = self.sample(ids, mask_ratio=masked_r, text=context, topk=topk, temperature=cur_temp,
                                       decode=keep or decode_every_step)
            if keep:
                host = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    host.copy_(img, non_blocking=True)
                img.record_stream(side)
                imgs.append(host)
        if ops.NVTX and timesteps > 0:
            torch.cuda.nvtx.range_pop()
        side.synchronize()
        return imgs
