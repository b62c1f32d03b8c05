"""VQModel — the drop-in for the reference's stage1/vqmodel.py:7-44.

Same attributes (encoder, decoder, quantize, prev_quant, post_quant), same methods and return
conventions, identical state_dict keys (222 tensors); the arithmetic runs in
hand-written sm_100a kernels through engine.Stage1Engine.  encode / decode / decode_from_indice are
inference entry points (no autograd graph, like the reference's frozen tokenizer, generate.py:55-56);
forward(img) — the call VQGANTrainer differentiates (utils/trainer.py:205-217) — carries gradients for
every parameter when autograd is enabled (train.py: one autograd.Function over the backward kernels)."""
from __future__ import annotations

import torch
from torch import nn

from .layers import Decoder, Encoder
from .quantize import VectorQuantizer


def _as_autocast(img):
    """Return-dtype convention of the reference under `torch.autocast('cuda', ...)` (its trainers call the model
    inside one, utils/trainer.py:187): the last op of decode is an autocast Linear, so the image comes back in the
    autocast dtype; z_q, loss (fp32) and indices (int64) do not change.  The kernels always compute
    in bf16 with fp32 accumulation; this only matches the dtype the caller's code expects."""
    try:
        on = torch.is_autocast_enabled("cuda")
        dt = torch.get_autocast_dtype("cuda") if on else None
    except TypeError:                                       # older torch: no device_type argument
        on = torch.is_autocast_enabled()
        dt = torch.get_autocast_gpu_dtype() if on else None
    return img.to(dt) if on and img.is_floating_point() else img


class VQModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder = Encoder(**config.enc)
        self.decoder = Decoder(**config.dec)
        self.quantize = VectorQuantizer(config.n_embed, config.embed_dim, config.beta)
        self.prev_quant = nn.Linear(config.enc["dim"], config.embed_dim)
        self.post_quant = nn.Linear(config.embed_dim, config.dec["dim"])
        self._engine = None

    # -- reference surface -------------------------------------------------------------------
    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    @torch.no_grad()
    def encode(self, x):
        """(z_q [B,N,32] fp32, loss [] fp32, indices [B,N] int64) — vqmodel.py:21-25."""
        return self.engine().encode(x)

    @torch.no_grad()
    def decode(self, x):
        """[B,N,32] -> image [B,3,H,W] clamped to [-1,1] — vqmodel.py:27-30."""
        return _as_autocast(self.engine().decode(x))

    def forward(self, img):
        """(rec, codebook loss) — vqmodel.py:32-36.  With autograd enabled and trainable parameters this is the
        generator training step: both outputs are differentiable w.r.t. every parameter."""
        if torch.is_grad_enabled() and img.shape[0] > 0 and any(p.requires_grad for p in self.parameters()):
            from ..train import vqgan_forward_with_grad
            rec, loss = vqgan_forward_with_grad(self, img)
            return _as_autocast(rec), loss
        z, loss, _ = self.encode(img)
        return self.decode(z), loss

    @torch.no_grad()
    def decode_from_indice(self, indice):
        return _as_autocast(self.engine().decode_from_indice(indice))

    def from_pretrained(self, path):
        """vqmodel.py:43-44.  `map_location="cpu"`: a checkpoint saved from another device (or a GPU this box does not
        have) must load; load_state_dict copies into the parameters wherever they live."""
        return self.load_state_dict(torch.load(path, map_location="cpu"))

    def invalidate(self):
        """Call after writing parameters through `.data` (EMA / weight surgery): such writes do not bump the version
        counter the packed-weight caches are keyed on."""
        if self._engine is not None:
            self._engine.invalidate()
        self.quantize.invalidate()
        self.__dict__.pop("_train_engine", None)

    # engines hold device buffers and a weak reference to THIS module: a copy / unpickled model builds its own
    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engine", "_train_engine"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new.__dict__["_engine"] = None
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        st.pop("_train_engine", None)
        return st

    # -- pixels in / pixels out (the steps either side of the path) ----------
    @torch.no_grad()
    def encode_pixels(self, img_u8):
        """encode() of decoded pixels: uint8 [B, H, W, 3] (PIL / numpy layout).  Equals
        encode(Normalize(0.5, 0.5)(ToTensor(img))) of the reference's ingest transform
        (utils/transform.py:17-18, reconstruct.py:31-37) — the transform is fused into the patch
        extraction kernel, so the fp32 NCHW image (4x the pixel bytes) never exists."""
        if img_u8.dtype != torch.uint8:
            raise TypeError("encode_pixels expects uint8 [B, H, W, 3]")
        return self.engine().encode(img_u8)

    @torch.no_grad()
    def decode_pixels(self, x):
        """restore(decode(x)) of reconstruct.py:11-16,37-39 as uint8 [B, H, W, 3]: (x+1)*0.5, HWC,
        uint8(255*x) — fused into the un-patchify epilogue of the last projection."""
        return self.engine().decode(x, pixels=True)

    @torch.no_grad()
    def decode_pixels_from_indice(self, indice):
        return self.engine().decode_from_indice(indice, pixels=True)

    # -- small-batch serving: one CUDA graph per input shape ----------------------------------
    def graphed(self, example, decode=True, pixels=False):
        """Capture encode (+ decode) for inputs shaped like `example` into a CUDA graph and return `run(x=None)`.

        A tokenize + detokenize pass is ~90 kernel launches; at batch 1-4 the host side of those launches (ctypes call,
        tensor-map encoding, output allocation: ~24 us each) is 3-4x the device time.  The graph replays the same
        kernels with the tensor maps baked in: `run(x)` copies x into the captured input buffer and launches once.
        Returns (rec or None, loss, indices, z_q) — STATIC tensors, overwritten by the next replay (clone to keep).
        Parameters are checked on every call: after an update the graph is re-captured.
        """
        from ..engine import _fingerprint
        if not example.is_cuda:
            raise RuntimeError("paintmind_b200: input must be a CUDA tensor (no CPU fallback)")
        x_static = example.detach().clone()
        state = {}

        def body():
            z, loss, idx = (self.encode_pixels(x_static) if pixels else self.encode(x_static))
            rec = None
            if decode:
                rec = self.decode_pixels(z) if pixels else self.decode(z)
            return rec, loss, idx, z

        def capture():
            side = torch.cuda.Stream(device=x_static.device)
            side.wait_stream(torch.cuda.current_stream(x_static.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):                      # packs weights, sizes every workspace buffer
                    body()
            torch.cuda.current_stream(x_static.device).wait_stream(side)
            torch.cuda.synchronize(x_static.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                out = body()
            state["graph"], state["out"], state["fp"] = g, out, _fingerprint(self)
            # raw addresses are baked into the graph: pin every workspace buffer and codebook operand it addresses, so that a
            # later eager call at another batch size (which makes the workspace drop same-named buffers) cannot free them
            q = self.quantize
            state["pinned"] = (list(self.engine().ws.bufs.values()) + [t for t in (q._prep or ())[1:]]
                               + [t for t in (q._cand or ())[1:]])

        @torch.no_grad()
        def run(x=None):
            if state.get("fp") != _fingerprint(self):
                capture()
            if x is not None:
                if x.shape != x_static.shape or x.dtype != x_static.dtype:
                    raise RuntimeError(f"graphed(): captured for {tuple(x_static.shape)} {x_static.dtype}, got {tuple(x.shape)} {x.dtype}")
                x_static.copy_(x, non_blocking=True)
            state["graph"].replay()
            return state["out"]

        with torch.cuda.device(x_static.device):
            capture()
        return run

    # -- engine ------------------------------------------------------------------------------
    def train_engine(self):
        if self.__dict__.get("_train_engine") is None:
            from ..train import Stage1TrainEngine
            object.__setattr__(self, "_train_engine", Stage1TrainEngine(self))
        return self._train_engine

    def engine(self):
        if self._engine is None:
            from ..engine import Stage1Engine
            object.__setattr__(self, "_engine", Stage1Engine(self))
        return self._engine
