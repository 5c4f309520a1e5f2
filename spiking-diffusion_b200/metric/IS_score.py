"""Inception score from class probabilities: the part of R/metric/IS_score.py:58-72 after the network's softmax."""
import torch

from .._lib import check, lib, on_device_of, ptr, stream_ptr
from .common import require_cuda, workspace


@on_device_of
def inception_score_from_probs(preds: torch.Tensor, splits: int = 1):
    """preds [N, K] (softmax outputs) -> (mean, std) over ``splits`` consecutive parts of exp(mean_i KL(p_i || p_mean)),
    with scipy.stats.entropy's normalisation of both arguments and numpy's population std.  0-d fp64 CUDA tensors."""
    require_cuda(preds)
    if preds.dim() != 2:
        raise ValueError("expected probabilities of shape [N, K]")
    p = preds.double().contiguous()
    N, K = p.shape
    if splits < 1 or N // splits < 1:
        raise ValueError("need at least one row per split")
    mean = torch.empty((), dtype=torch.float64, device=p.device)
    std = torch.empty((), dtype=torch.float64, device=p.device)
    ws = workspace(p.device, n_elements=K + N // splits + 128)
    check(lib().sd_metric_inception_score(ptr(p), N, K, int(splits), ptr(mean), ptr(std), ptr(ws), stream_ptr()))
    return mean, std
