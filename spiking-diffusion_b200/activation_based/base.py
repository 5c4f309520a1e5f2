"""Step-mode and stateful-module protocol (mirrors SJ/activation_based/base.py:52-117,153-447).

``functional.reset_net`` / ``set_step_mode`` walk ``net.modules()`` looking for ``reset`` / ``step_mode``
attributes, so stateful modules must expose exactly this protocol: memories registered by name with a reset
value, readable and writable as plain attributes, moved by ``.to()/.cuda()`` and copied by DataParallel.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .._lib import on_device_of as _on_device_of


class StepModule:
    def supported_step_mode(self):
        return ("s", "m")

    @property
    def step_mode(self):
        return self._step_mode

    @step_mode.setter
    def step_mode(self, value: str):
        # same condition and exception type as SJ/activation_based/base.py:115-116
        if value not in self.supported_step_mode():
            raise ValueError(f'step_mode can only be {self.supported_step_mode()}, but got "{value}"!')
        self._step_mode = value


class LazyState:
    """A memory value that is produced on first access (the fused whole-model plans keep LIF states in their planar
    kernel layout; the reference-layout tensor is only built if somebody reads ``node.v`` before the next reset)."""

    def __init__(self, fn):
        self._fn = fn

    def materialize(self):
        return self._fn()


class ConsumedState(LazyState):
    """Marks a LIF state that a fused whole-network kernel chain consumed without writing it back (the denoiser's
    plan keeps membrane potentials in registers / L2 only).  The reference would continue from the carried state on
    the next forward; here that is impossible, so reading the state or calling forward again raises until
    ``reset()`` (``functional.reset_net``) -- which every reference call site issues -- clears the mark."""

    def __init__(self, what: str):
        self._what = what

    def materialize(self):
        raise RuntimeError(f"the LIF state of {self._what} was consumed inside the fused kernels of the last forward; "
                           "call functional.reset_net(model) before reading it or running another forward")


class MemoryModule(nn.Module, StepModule):
    def __init__(self):
        super().__init__()
        self._memories = {}
        self._memories_rv = {}
        self._backend = "torch"
        self.step_mode = "s"

    @property
    def supported_backends(self):
        return ("torch",)

    @property
    def backend(self):
        return self._backend

    @backend.setter
    def backend(self, value: str):
        # SJ/activation_based/base.py:199-208
        if value not in self.supported_backends:
            raise NotImplementedError(f"{value} is not a supported backend of {self._get_name()}!")
        self._backend = value

    def single_step_forward(self, x: torch.Tensor, *args, **kwargs):
        raise NotImplementedError

    def multi_step_forward(self, x_seq: torch.Tensor, *args, **kwargs):
        ys = [self.single_step_forward(x_seq[t], *args, **kwargs).unsqueeze(0) for t in range(x_seq.shape[0])]
        return torch.cat(ys, 0)

    @_on_device_of
    def forward(self, *args, **kwargs):
        if self.step_mode == "s":
            return self.single_step_forward(*args, **kwargs)
        if self.step_mode == "m":
            return self.multi_step_forward(*args, **kwargs)
        raise ValueError(self.step_mode)

    def extra_repr(self):
        return f"step_mode={self.step_mode}, backend={self.backend}"

    def register_memory(self, name: str, value):
        assert not hasattr(self, name), f"{name} has been set as a member variable!"
        self._memories[name] = value
        self.set_reset_value(name, value)

    def reset(self):
        for key in self._memories.keys():
            self._memories[key] = copy.deepcopy(self._memories_rv[key])

    def set_reset_value(self, name: str, value):
        self._memories_rv[name] = copy.deepcopy(value)

    def __getattr__(self, name: str):
        if "_memories" in self.__dict__:
            memories = self.__dict__["_memories"]
            if name in memories:
                value = memories[name]
                if isinstance(value, LazyState):
                    value = memories[name] = value.materialize()
                return value
        return super().__getattr__(name)

    def __setattr__(self, name: str, value) -> None:
        _memories = self.__dict__.get("_memories")
        if _memories is not None and name in _memories:
            _memories[name] = value
        else:
            super().__setattr__(name, value)

    def __delattr__(self, name):
        if name in self._memories:
            del self._memories[name]
            del self._memories_rv[name]
        else:
            super().__delattr__(name)

    def _resolve_lazy(self):
        for key, value in self._memories.items():
            if isinstance(value, LazyState) and not isinstance(value, ConsumedState):
                self._memories[key] = value.materialize()

    def __getstate__(self):
        # pickling / torch.save(module): closures of lazy states do not pickle, tensors do
        self._resolve_lazy()
        state = dict(self.__dict__)
        state["_memories"] = {k: (copy.deepcopy(self._memories_rv[k]) if isinstance(v, ConsumedState) else v)
                              for k, v in self._memories.items()}
        return state

    def memory_is_reset(self, name: str) -> bool:
        """True if the memory still holds its (non-tensor) reset value; does not materialise a lazy state."""
        value = self._memories[name]
        return not isinstance(value, (torch.Tensor, LazyState))

    def memories(self):
        self._resolve_lazy()
        for value in self._memories.values():
            yield value

    def named_memories(self):
        self._resolve_lazy()
        for name, value in self._memories.items():
            yield name, value

    def detach(self):
        for key, value in self._memories.items():
            if isinstance(value, torch.Tensor):
                value.detach_()

    def _apply(self, fn, *args, **kwargs):
        self._resolve_lazy()
        for key, value in self._memories.items():
            if isinstance(value, torch.Tensor):
                self._memories[key] = fn(value)
        return super()._apply(fn, *args, **kwargs)

    def _replicate_for_data_parallel(self):
        self._resolve_lazy()
        replica = super()._replicate_for_data_parallel()
        replica._memories = self._memories.copy()
        return replica
