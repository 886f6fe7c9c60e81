"""Host-side parameter sampling.

Parameters of sample i depend only on (seed, global sample index, transform stream), never on
the batch it travels in or the GPU it lands on -- so a batch sharded over N GPUs draws the single-GPU
run's parameters bit for bit (what that means for the results: sharding.py).  The generator is a counter-based splitmix64 hash evaluated
with vectorised NumPy uint64 arithmetic (no per-sample Generator construction).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def counter_uniform(seed: int, sample_index: np.ndarray, stream: int, n_draws: int) -> np.ndarray:
    """U[0,1) doubles of shape [len(sample_index), n_draws], a pure function of its arguments."""
    idx = np.asarray(sample_index, dtype=np.uint64).reshape(-1, 1)
    k = np.arange(n_draws, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ _splitmix64(np.uint64(stream) + np.uint64(0x1234567)))
        h = _splitmix64(key ^ _splitmix64(idx * np.uint64(0xD1342543DE82EF95) + k))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniforms(batch_size: int, n_draws: int, *, seed: int, stream: int, generator=None,
             sample_index=None) -> np.ndarray:
    """[B, n_draws] uniforms.

    generator=None          -> counter-based, keyed by (seed, sample_index or arange(B), stream)
    numpy Generator         -> sequential draws from it
    torch.Generator (CPU)   -> sequential draws through torch.rand
    """
    if generator is None:
        if sample_index is None:
            sample_index = np.arange(batch_size)
        sample_index = np.asarray(sample_index).reshape(-1)
        if sample_index.shape[0] != batch_size:
            raise ValueError("sample_index must have batch_size entries")
        return counter_uniform(seed, sample_index, stream, n_draws)
    if isinstance(generator, np.random.Generator):
        return generator.random((batch_size, n_draws))
    import torch
    if isinstance(generator, torch.Generator):
        return torch.rand((batch_size, n_draws), generator=generator, dtype=torch.float64).numpy()
    raise TypeError("generator must be None, numpy.random.Generator or torch.Generator")
