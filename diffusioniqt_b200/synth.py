"""Deterministic synthetic weights and inputs (no checkpoint or dataset ships with the reference).

Everything is drawn from `numpy.random.RandomState` (the frozen legacy MT19937 stream), keyed
by tensor name, so the same values can be regenerated on any box without torch's RNG and
without storing megabytes of weights next to the golden fixtures.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Mapping, Tuple

import numpy as np
import torch


def _rs(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synthetic_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Weight-like values for a parameter called `name`.

    Matrices / conv kernels: uniform(-g/sqrt(fan_in), +g/sqrt(fan_in)) with a gain that keeps
    activations O(1) through ~40 conv layers.  GroupNorm scale: 1 + small jitter.  Biases and
    GroupNorm shifts: small uniform values.
    """
    rs = _rs(name, seed)
    shape = tuple(int(s) for s in shape)
    if name.endswith("groupnorm.weight") or name == "norm_cond.weight":
        v = 1.0 + 0.2 * rs.uniform(-1, 1, shape)
    elif name.endswith(".bias"):
        v = 0.1 * rs.uniform(-1, 1, shape)
    elif name.endswith("to_time_hiddens.0.weights"):
        v = rs.standard_normal(shape)
    else:
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        gain = 1.0
        v = rs.uniform(-1, 1, shape) * (gain * np.sqrt(3.0 / max(fan_in, 1)))
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synthetic_state_dict(shapes: Mapping[str, Iterable[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synthetic_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def fill_module_(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter/buffer of `module` with `synthetic_tensor(name, shape, seed)`."""
    sd = module.state_dict()
    new = synthetic_state_dict({k: v.shape for k, v in sd.items()}, seed)
    module.load_state_dict(new)


def synthetic_field(shape: Tuple[int, ...], seed: int, smooth: int = 2) -> torch.Tensor:
    """A z-score-like random field (low-pass filtered gaussian noise, unit variance)."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    v = rs.standard_normal(shape).astype(np.float32)
    for ax in range(len(shape) - 3, len(shape)):
        for _ in range(smooth):
            v = (np.roll(v, 1, ax) + 2 * v + np.roll(v, -1, ax)) * 0.25
    v = (v - v.mean()) / (v.std() + 1e-8)
    return torch.from_numpy(np.ascontiguousarray(v))


def synthetic_noise(shape: Tuple[int, ...], count: int, seed: int):
    """`count` standard-normal tensors (the injected sampler noise sequence)."""
    rs = np.random.RandomState((seed * 7919 + 13) & 0x7FFFFFFF)
    return [torch.from_numpy(rs.standard_normal(shape).astype(np.float32)) for _ in range(count)]
