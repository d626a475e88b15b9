"""GPU input pipeline with the tensor contract of `NuscenesDataset.__getitem__`
(src/data/dataloader.py:202-333): ImageNet normalisation of the uint8 camera image into the first three
input channels, inverse-normalised lidar ground truth and its zero-ignoring min-pool pyramid
(dataloader.py:213-222,240-257), nearest-neighbour resize of the segmentation labels (dataloader.py:262-268).
`pack_input_nhwc` writes the whole network input directly in the engine's NHWC bf16 layout, so that
`CamRaDepth.forward_packed` can skip the NCHW fp32 -> NHWC bf16 pack of the nn.Module boundary."""
from __future__ import annotations

import ctypes

import torch

from ._lib import K
from .ops import P, stream

IMAGENET_MEAN = (0.485, 0.456, 0.406)       # dataloader.py:230
IMAGENET_STD = (0.229, 0.224, 0.225)        # dataloader.py:231


def normalize_image(img_u8: torch.Tensor, out: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """img_u8: (B,H,W,3) uint8 CUDA (channel order as read, like the reference); out: (B,C>=3,H,W) fp32."""
    if not img_u8.is_cuda:
        raise RuntimeError("camradepth_b200 preprocessing runs on CUDA devices only")
    B, H, W, three = img_u8.shape
    assert three == 3 and img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] == B and out.shape[2:] == (H, W)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    K.crd_image_normalize(P(img_u8), P(out), B, H, W, out.shape[1], ctypes.cast(m, ctypes.c_void_p),
                          ctypes.cast(s, ctypes.c_void_p), stream())
    return out


def gt_pyramid(lidar_depth_m: torch.Tensor, max_depth: float = 100.0):
    """(B,1,H,W) metric lidar depth (0 = no return) -> (gt_final, gt_stage4 (H/2), gt_stage3 (H/4)) inverse-
    normalised maps; the naming follows runner.py:185 (lidar_depth_partial[0] is the half-resolution map)."""
    d = lidar_depth_m.detach().contiguous().float()
    B, _, H, W = d.shape
    g = torch.empty_like(d)
    K.crd_gt_normalize(P(d), P(g), d.numel(), float(max_depth), stream())
    outs = [g]
    cur, h, w = g, H, W
    for _ in range(2):
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        nxt = torch.empty(B, 1, ho, wo, dtype=torch.float32, device=d.device)
        K.crd_minpool3x3s2(P(cur), P(nxt), B, h, w, stream())
        outs.append(nxt)
        cur, h, w = nxt, ho, wo
    return tuple(outs)


def seg_resize_nearest(labels: torch.Tensor, size) -> torch.Tensor:
    """(B,Hi,Wi) uint8 / int64 label maps -> (B,Ho,Wo) int64 with the sampling rule of
    skimage.transform.resize(order=0, preserve_range=True, anti_aliasing=False) (dataloader.py:262-268)."""
    if not labels.is_cuda:
        raise RuntimeError("camradepth_b200 preprocessing runs on CUDA devices only")
    assert labels.dim() == 3 and labels.dtype in (torch.uint8, torch.int64) and labels.is_contiguous()
    B, Hi, Wi = labels.shape
    Ho, Wo = int(size[0]), int(size[1])
    out = torch.empty(B, Ho, Wo, dtype=torch.int64, device=labels.device)
    K.crd_seg_resize_nearest(P(labels), int(labels.dtype == torch.uint8), P(out), B, Hi, Wi, Ho, Wo, stream())
    return out


def pack_input_nhwc(img_u8: torch.Tensor, extra: torch.Tensor = None, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """img_u8 (B,H,W,3) uint8 + extra (B,Ce<=5,H,W) fp32 radar planes (already scaled like dataloader.py:301-323)
    -> (B,H,W,8) bf16 NHWC network input [normalised RGB | extra | zeros] for `CamRaDepth.forward_packed`."""
    if not img_u8.is_cuda:
        raise RuntimeError("camradepth_b200 preprocessing runs on CUDA devices only")
    B, H, W, three = img_u8.shape
    assert three == 3 and img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    Ce = 0
    if extra is not None:
        assert extra.is_contiguous() and extra.dtype == torch.float32 and extra.shape[0] == B and extra.shape[2:] == (H, W)
        Ce = extra.shape[1]
    out = torch.empty(B, H, W, 8, dtype=torch.bfloat16, device=img_u8.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    K.crd_pack_input_nhwc(P(img_u8), P(extra), P(out), B, H, W, Ce, 8, ctypes.cast(m, ctypes.c_void_p),
                          ctypes.cast(s, ctypes.c_void_p), stream())
    return out
