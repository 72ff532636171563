"""Shared helpers for the test-suite (CPU side of the tile-image layout, oracle/model bridges)."""
import numpy as np
import torch

from oracle import moco_oracle as orc


def to_images(x: torch.Tensor) -> np.ndarray:
    """Row-major (rows, cols) float tensor -> the bf16 128B-swizzled tile images the kernels use:
    uint16 array [n_tiles][cols/64 blocks][128 rows][64], with 16-byte chunk c of row r stored at c ^ (r & 7)."""
    rows, cols = x.shape
    T, B = (rows + 127) // 128, (cols + 63) // 64
    pad = torch.zeros(T * 128, B * 64, dtype=torch.float32)
    pad[:rows, :cols] = x.float()
    bits = pad.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    a = bits.reshape(T, 128, B, 8, 8).transpose(0, 2, 1, 3, 4)  # [T][B][r][chunk][8]
    out = np.empty_like(a)
    r = np.arange(128)
    for c in range(8):
        out[:, :, r, c ^ (r & 7), :] = a[:, :, r, c, :]
    return np.ascontiguousarray(out)


def from_images(img: np.ndarray, rows: int, cols: int) -> torch.Tensor:
    """Inverse of to_images -> float32 (rows, cols)."""
    T, B = img.shape[0], img.shape[1]
    a = np.empty_like(img)
    r = np.arange(128)
    for c in range(8):
        a[:, :, r, c, :] = img[:, :, r, c ^ (r & 7), :]
    flat = a.transpose(0, 2, 1, 3, 4).reshape(T * 128, B * 64)
    t = torch.from_numpy(flat.view(np.int16).copy()).view(torch.bfloat16).float()
    return t[:rows, :cols]


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def load_nerf(model, params):
    model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    return model


def pe_module(spec: orc.PESpec, Embedding):
    e = Embedding(spec.in_channels, spec.n_freqs, spec.logscale)
    if spec.weights is not None:
        e.set_weights(list(spec.weights))
    return e
