"""Quality metrics on the GPU, under the reference's module names (R/metric/).

What is here is the arithmetic that needs no pretrained network: the reconstruction test's MSE and SSIM
(``metric.pytorch_ssim``), the Frechet distance on feature statistics (``metric.Fid_score``), the inception-score
formula on class probabilities (``metric.IS_score``) and the kernel-inception-distance estimator on features
(``metric.kid``).  The Inception-v3 feature extractor the reference downloads (R/metric/Fid_score.py:39,
R/metric/IS_score.py:42) is NOT reproduced: its weights are not available offline, so every function takes the
network's outputs (features [N, 2048] / probabilities [N, 1000]) instead of images.  Every function runs hand-written
CUDA kernels of libsd_b200 (csrc/metrics.cu) on CUDA tensors; there is no CPU path.
"""
from . import Fid_score, IS_score, kid, pytorch_ssim  # noqa: F401
from .common import mse_loss  # noqa: F401
