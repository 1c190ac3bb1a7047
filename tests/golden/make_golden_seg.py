"""Golden vectors for the supervised-segmentation edges of the path (loss, metrics, input transport), produced by the
LIVE reference (/root/reference) in the build container:

    python tests/golden/make_golden_seg.py        -> tests/golden/seg_small.npz

  * trainers/train_baseline.py: dice / precision / recall and the BCE loss expression of the training loop (:44-45),
    with TEDM's label repetition (:30-31);
  * dataloaders/JSRT.py: JSRTDataset.load_image / load_labels on 8-bit files written at the target size (the PIL
    resize is then the identity), including an overlapping-lung case.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def main():
    from einops import reduce, repeat
    from PIL import Image
    from torch.nn.functional import binary_cross_entropy_with_logits
    from trainers.train_baseline import dice, precision, recall
    from dataloaders.JSRT import JSRTDataset

    rng = np.random.Generator(np.random.PCG64(7))
    out = {}
    B, S, H = 3, 4, 32
    logits = torch.from_numpy(rng.normal(0, 2.0, (B * S, 1, H, H)).astype(np.float32))
    logits[0, 0, 0, :8] = torch.tensor([0.0, 1e-9, -1e-9, 3e-7, -3e-7, 2e-7, -0.0, 3e-8])   # around the .5 threshold
    # (for 3e-8 < x < 2e-7 torch's own CPU and CUDA sigmoid round differently, so no value is pinned there)
    y = torch.from_numpy((rng.random((B, 1, H, H)) > 0.6).astype(np.float32))
    y[2] = 0.0                                                                              # an empty label -> NaN rows
    logits[8:12] = -5.0 - logits[8:12].abs()                                                # ... and empty predictions
    y_rep = repeat(y, "b c h w -> (b step) c h w", step=S)
    rows = reduce(binary_cross_entropy_with_logits(logits, y_rep, reduction="none"), "b c h w -> b c", "mean")
    y_hat = torch.sigmoid(logits) > .5
    out.update(logits=logits.numpy(), y=y.numpy(), n_steps=np.int64(S), bce_rows=rows.numpy(), bce_loss=rows.mean().numpy(),
               y_hat=y_hat.numpy(), dice=dice(y_hat, y_rep).numpy(), precision=precision(y_hat, y_rep).numpy(),
               recall=recall(y_hat, y_rep).numpy())
    # gradient of the training loss w.r.t. the logits
    lg = logits.clone().requires_grad_(True)
    reduce(binary_cross_entropy_with_logits(lg, y_rep, reduction="none"), "b c h w -> b c", "mean").mean().backward()
    out["bce_grad"] = lg.grad.numpy()

    # ---- loader arithmetic -------------------------------------------------------------------------------------------
    s = 16
    img = rng.integers(0, 256, (s, s), dtype=np.uint8)
    img.flat[:6] = [0, 1, 127, 128, 254, 255]
    right = (rng.random((s, s)) > 0.7).astype(np.uint8) * 255
    left = (rng.random((s, s)) > 0.7).astype(np.uint8) * 255
    right.flat[:4] = [127, 128, 129, 0]                      # grey values on both sides of the .5 threshold
    left_disjoint = left.copy()
    left_disjoint[right >= 128] = 0
    with tempfile.TemporaryDirectory() as d:
        for name, arr in (("img.png", img), ("right.png", right), ("left.png", left), ("left_d.png", left_disjoint)):
            Image.fromarray(arr, mode="L").save(os.path.join(d, name))
        ds = JSRTDataset.__new__(JSRTDataset)
        ds.base_path, ds.img_size = __import__("pathlib").Path(d), s
        out["u8_img"] = img
        out["f32_img"] = ds.load_image("img.png").numpy()
        out["u8_masks_overlap"] = np.stack([right, left])
        out["label_overlap"] = ds.load_labels(["right.png", "left.png"]).float().numpy()
        out["u8_masks_disjoint"] = np.stack([right, left_disjoint])
        out["label_disjoint"] = ds.load_labels(["right.png", "left_d.png"]).float().numpy()
    np.savez_compressed(os.path.join(HERE, "seg_small.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
