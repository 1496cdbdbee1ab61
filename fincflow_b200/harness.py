"""Formats of the reference's training harness (fastflow/train/experiment.py), so that its checkpoints
can be sampled with the B200 kernels and the outputs look the same:

  * checkpoints: the dict `Experiment.save` writes (experiment.py:400-413) --
    {'summary', 'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'config'};
    models trained under nn.DataParallel carry a `module.` key prefix (experiment.py:174-176);
  * sample grids: `torchvision.utils.save_image(x / 256., path, nrow=10, padding=2, normalize=False)`
    (experiment.py:342-345) restated without torchvision: same grid geometry, same rounding, PNG
    written with zlib from the standard library.
"""
from __future__ import annotations

import struct
import zlib

import torch

CHECKPOINT_KEYS = ("summary", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "config")


def save_checkpoint(path, model, optimizer=None, scheduler=None, summary=None, config=None):
    """write the reference's checkpoint dict (train/experiment.py:400-413)"""
    ckpt = {
        "summary": summary if summary is not None else {},
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
        "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else {},
        "config": config if config is not None else {},
    }
    torch.save(ckpt, path)
    return ckpt


def _strip_module_prefix(sd):
    if sd and all(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return sd


# keys the reference's modules own that carry no model state here
IGNORABLE_KEYS = ("preprocess.layers.0.distribution.empty", "base_distribution.empty")


def load_checkpoint(path_or_dict, model, optimizer=None, scheduler=None, map_location="cpu", strict=True):
    """load a checkpoint written by the reference's `Experiment.save` (or by save_checkpoint) into
    `model` (train/experiment.py:415-427).  A bare state dict is accepted too.  Returns
    (summary, config)."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location=map_location,
                                                                          weights_only=False)
    is_experiment = "model_state_dict" in ckpt
    sd = _strip_module_prefix(dict(ckpt["model_state_dict"] if is_experiment else ckpt))
    res = model.load_state_dict(sd, strict=False)
    unexpected = [k for k in res.unexpected_keys if not k.endswith(IGNORABLE_KEYS)]
    if strict and (res.missing_keys or unexpected):
        raise RuntimeError(f"checkpoint does not match the model: missing {res.missing_keys}, unexpected {unexpected}")
    if is_experiment:
        if optimizer is not None and ckpt.get("optimizer_state_dict"):
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if scheduler is not None and ckpt.get("scheduler_state_dict"):
            scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        return ckpt.get("summary", {}), ckpt.get("config", {})
    return {}, {}


# ---------------------------------------------------------------------------------------------
# sample grids
# ---------------------------------------------------------------------------------------------
def make_grid(x, nrow=10, padding=2, pad_value=0.0):
    """[B, C, H, W] (or [B, H, W]) -> [3 or C, gh, gw] with torchvision.utils.make_grid's geometry:
    min(nrow, B) images per row, `padding` pixels around every image, single-channel images
    replicated to three channels."""
    if x.dim() == 3:
        x = x.unsqueeze(1)
    if x.shape[1] == 1:
        x = x.expand(-1, 3, -1, -1)
    B, C, H, W = x.shape
    xmaps = min(nrow, B)
    ymaps = -(-B // xmaps)
    hh, ww = H + padding, W + padding
    grid = x.new_full((C, hh * ymaps + padding, ww * xmaps + padding), pad_value)
    for k in range(B):
        r, c = divmod(k, xmaps)
        grid[:, r * hh + padding:r * hh + padding + H, c * ww + padding:c * ww + padding + W] = x[k]
    return grid


def _png_bytes(img_u8):
    """[H, W, 3] or [H, W, 1] uint8 numpy -> PNG file bytes (8-bit truecolour / greyscale, no filter)"""
    h, w, c = img_u8.shape
    raw = b"".join(b"\x00" + img_u8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


def save_image_grid(x, path, nrow=10, padding=2, scale=1.0 / 256.0):
    """the reference's sample image: save_image(x / 256., path, nrow=10, padding=2, normalize=False)
    (train/experiment.py:342-345; torchvision: grid.mul(255).add_(0.5).clamp_(0, 255).to(uint8))"""
    grid = make_grid(x.detach().float().cpu() * scale, nrow=nrow, padding=padding)
    img = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).contiguous().numpy()
    with open(path, "wb") as f:
        f.write(_png_bytes(img))
    return img


def sample_to_png(model, n_samples, path, nrow=10):
    """`Experiment.sample` (train/experiment.py:297-345): n samples under no_grad -> PNG grid"""
    with torch.no_grad():
        x, _ = model.sample(n_samples)
    return save_image_grid(x, path, nrow=nrow)
