"""Device-side mirror of the reference's input transform (quick_start/keep_inference.py:88-93, repeated in
WSI_evaluation/zeroshot_*_WSI.py:38-43):

    transforms.Resize(224, BICUBIC) -> CenterCrop(224) -> ToTensor -> Normalize(ImageNet mean/std)

`preprocess(tiles)` performs the first two steps on uint8 RGB tiles that are already on the GPU, bit-identically to
torchvision + Pillow; the last two are fused into the patch gather of `KEEPModel.encode_image` for uint8 input, so

    feats = model.encode_image(preprocess(tiles_u8))          # == encode_image(stack(transform(PIL tile)))
"""
from __future__ import annotations

import torch

from . import _lib


def preprocess(tiles: torch.Tensor, size: int = 224) -> torch.Tensor:
    """uint8 [B,H,W,3] (CUDA) -> uint8 [B,size,size,3]: Resize(size, bicubic, PIL semantics) + CenterCrop(size)."""
    if not isinstance(tiles, torch.Tensor) or tiles.dim() != 4 or tiles.shape[3] != 3 or tiles.dtype != torch.uint8:
        raise ValueError("preprocess expects a uint8 tensor [B,H,W,3]")
    if not tiles.is_cuda:
        raise _lib.KeepB200Error("preprocess: tiles must be on a CUDA device (keep_b200 has no CPU path)")
    tiles = tiles.contiguous()
    B, H, W, _ = tiles.shape
    out = torch.empty(B, size, size, 3, dtype=torch.uint8, device=tiles.device)
    if B == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(tiles.device):
        need = L.keepb200_preprocess_workspace_bytes(B, H, W, size)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=tiles.device)
        base = (ws.data_ptr() + 255) // 256 * 256
        _lib.check(L.keepb200_preprocess_u8(tiles.data_ptr(), B, H, W, size, out.data_ptr(), base, need,
                                            _lib.stream_ptr(tiles.device)), "preprocess_u8")
    return out
