"""SSIM with the reference's interface (R/metric/pytorch_ssim/__init__.py:7-73): ``SSIM(window_size=11,
size_average=True)(img1, img2)`` and ``ssim(img1, img2)``; one fused CUDA kernel instead of five grouped convolutions.
Used by the reconstruction test as ``1 - SSIM(window_size=11)(recon, images)`` (R/main.py:320-321)."""
import ctypes
from math import exp

import torch

from ..._lib import check, lib, on_device_of, ptr, stream_ptr
from ..common import require_cuda, workspace


def gaussian(window_size, sigma):
    # same construction as the reference (python-float exp -> fp32 tensor -> fp32 normalisation), so the window bits agree
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel=1):
    """[channel, 1, ws, ws] like the reference; the kernel applies one window to every channel (they are all equal)."""
    _1d = gaussian(window_size, 1.5).unsqueeze(1)
    _2d = _1d.mm(_1d.t()).float().unsqueeze(0).unsqueeze(0)
    return _2d.expand(channel, 1, window_size, window_size).contiguous()


@on_device_of
def _ssim(img1, img2, window_size, size_average=True):
    require_cuda(img1, img2)
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise ValueError(f"ssim expects two [N, C, H, W] tensors of one shape, got {tuple(img1.shape)} and {tuple(img2.shape)}")
    if window_size % 2 == 0 or window_size > 15:
        raise ValueError("window_size must be odd and <= 15")
    N, C, H, W = img1.shape
    a, b = img1.contiguous().float(), img2.contiguous().float()
    win = create_window(window_size)[0, 0].contiguous()
    win_host = (ctypes.c_float * (window_size * window_size))(*win.reshape(-1).tolist())
    out = torch.empty((), dtype=torch.float32, device=a.device)
    per_plane = None if size_average else torch.empty(N * C, dtype=torch.float32, device=a.device)
    ws = workspace(a.device)
    check(lib().sd_metric_ssim(ptr(a), ptr(b), N, C, H, W, window_size, ctypes.cast(win_host, ctypes.c_void_p), ptr(out),
                               ptr(per_plane), ptr(ws), stream_ptr()))
    if size_average:
        return out
    return (per_plane / float(H * W)).reshape(N, C).mean(1)        # ssim_map.mean(1).mean(1).mean(1)


class SSIM(torch.nn.Module):
    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1
        self.window = create_window(window_size, self.channel)

    def forward(self, img1, img2):
        return _ssim(img1, img2, self.window_size, self.size_average)


def ssim(img1, img2, window_size=11, size_average=True):
    return _ssim(img1, img2, window_size, size_average)
