"""BASELINE configs[3]: batch-sharded tokenize + detokenize of 8192 synthetic images over the ranks of one box, with the
path's only collective — one all-reduce of the codebook-usage histogram and of (sum of squared errors, element count).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 scripts/config4_sharded_tokenize.py

Image content depends only on the GLOBAL batch index (seed = 1000 + global batch index, 256 images per batch), so the
reduced histogram must equal, bit for bit, the histogram a single GPU computes over all 8192 images; rank 0 recomputes
that single-GPU histogram and asserts equality.  Prints one JSON line."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200 import dist as pmdist  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

TOTAL, BATCH = int(os.environ.get("PM_CONFIG4_IMAGES", "8192")), 256


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # the banner goes to stdout whatever NCCL_DEBUG_FILE says
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    cfg = ver2cfg["vit-s-vqgan"]
    model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0), strict=True)
    model = model.to(dev).eval()
    n_batches = TOTAL // BATCH

    def batch_images(gb):
        g = torch.Generator(device=dev).manual_seed(1000 + gb)
        return torch.rand(BATCH, 3, 256, 256, device=dev, generator=g) * 2 - 1

    def run(batches):
        hist = torch.zeros(cfg["n_embed"], device=dev, dtype=torch.int64)
        sums = torch.zeros(2, device=dev, dtype=torch.float64)
        chk = torch.zeros((), device=dev, dtype=torch.float64)
        for gb in batches:
            z, _, _ = model.encode(batch_images(gb))
            rec = model.decode(z)
            hist += model.quantize._last_hist
            sums[0:1] += model.quantize._last_sse
            sums[1] += z.numel()
            chk += rec.double().abs().sum()
        return hist, sums, chk

    lo, hi = pmdist.shard_range(n_batches, rank, world)            # contiguous shards of whole batches
    run(range(lo, min(lo + 1, hi)))                                # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hist, sums, chk = run(range(lo, hi))
    pmdist.allreduce_usage(hist, sums)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref_hist, ref_sums, _ = run(range(n_batches))              # the W = 1 answer
        same = bool(torch.equal(hist, ref_hist))
        print(json.dumps({"config": "BASELINE configs[3]", "images": TOTAL, "n_gpus": world, "ms": float(t), "images_per_s": TOTAL / (float(t) * 1e-3),
                          "histogram_equals_single_gpu": same, "codes_used": int((hist > 0).sum()),
                          "loss": pmdist.global_loss(sums), "loss_single_gpu": pmdist.global_loss(ref_sums)}), flush=True)
        assert same, "sharded histogram differs from the single-GPU histogram"
        assert abs(pmdist.global_loss(sums) - pmdist.global_loss(ref_sums)) < 1e-12
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
