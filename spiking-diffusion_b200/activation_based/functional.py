"""Net-wide helpers (mirrors SJ/activation_based/functional.py:13-40, 42-106, 109-148, 653-688)."""
import logging

import torch.nn as nn

from . import base


def reset_net(net: nn.Module):
    """Call ``reset()`` on every sub-module that has one (functional.py:35-40)."""
    for m in net.modules():
        if hasattr(m, "reset"):
            if not isinstance(m, base.MemoryModule):
                logging.warning(f"Trying to call `reset()` of {m}, which is not a MemoryModule")
            m.reset()


def set_step_mode(net: nn.Module, step_mode: str):
    """Set ``step_mode`` on every sub-module that has the attribute (functional.py:92-106)."""
    for m in net.modules():
        if hasattr(m, "step_mode"):
            if not isinstance(m, base.StepModule):
                logging.warning(f"Trying to set the step mode for {m}, which is not a StepModule")
            m.step_mode = step_mode


def set_backend(net: nn.Module, backend: str, instance=(nn.Module,)):
    """Set ``backend`` on every matching sub-module that supports it (functional.py:109-148)."""
    for m in net.modules():
        if isinstance(m, instance) and hasattr(m, "backend"):
            if not isinstance(m, base.MemoryModule):
                logging.warning(f"Trying to set the backend for {m}, which is not a MemoryModule")
            if backend in m.supported_backends:
                m.backend = backend
            else:
                logging.warning(f"{m} does not supports for backend={backend}. It will still use backend={m.backend}.")


def detach_net(net: nn.Module):
    for m in net.modules():
        if hasattr(m, "detach"):
            m.detach()


def seq_to_ann_forward(x_seq, stateless_module):
    """Flatten [T, N, ...] -> [T*N, ...], apply, un-flatten (functional.py:680-688)."""
    y_shape = [x_seq.shape[0], x_seq.shape[1]]
    y = x_seq.flatten(0, 1)
    if isinstance(stateless_module, (list, tuple, nn.Sequential)):
        for m in stateless_module:
            y = m(y)
    else:
        y = stateless_module(y)
    y_shape.extend(y.shape[1:])
    return y.view(y_shape)


def convert_sync_batchnorm(net, process_group=None):
    """Make every train-mode ``layer.BatchNorm2d`` of ``net`` compute its batch statistics over all ranks of
    ``process_group`` (default: the world group) - the counterpart of ``torch.nn.SyncBatchNorm.convert_sync_batchnorm``
    for data-parallel training of these models (optional training path).  Eval mode is
    unaffected (running statistics).  Returns ``net``."""
    import torch.distributed as dist
    from . import layer
    group = process_group if process_group is not None else (dist.group.WORLD if dist.is_initialized() else None)
    for m in net.modules():
        if isinstance(m, layer.BatchNorm2d):
            m.sync_group = group
    return net
