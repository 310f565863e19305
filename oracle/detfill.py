"""Deterministic, RNG-library-independent tensor fill (test infrastructure).

Golden fixtures cannot carry 40M+ weights, so both the fixture generator (which
drives the real reference) and the tests (which drive the oracle / the CUDA
path) regenerate identical weights and inputs from (name, shape) with a
splitmix64 counter hash written in plain numpy uint64 arithmetic.
"""
import zlib
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + _G).astype(np.uint64)
        x = (x ^ (x >> np.uint64(30))) * _M1
        x = (x ^ (x >> np.uint64(27))) * _M2
        x = x ^ (x >> np.uint64(31))
    return x


def uniform(name, shape, lo=-1.0, hi=1.0, salt=0):
    """float32 array of `shape`, element i = hash(crc32(name), salt, i) mapped to [lo, hi)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(name.encode()) & 0xFFFFFFFF) << np.uint64(32)
    seed = seed | np.uint64(salt & 0xFFFFFFFF)
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(_splitmix64(seed) ^ (idx * _G))
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)  # 24-bit mantissa, exact in fp32
    out = (lo + (hi - lo) * u).astype(np.float32)
    return out.reshape(shape)


def normalish(name, shape, salt=0):
    """Approximately N(0,1): sum of 4 uniforms, variance-normalised (exactly reproducible)."""
    acc = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        acc += uniform(name, shape, -1.0, 1.0, salt=salt * 4 + k + 1000).astype(np.float64)
    return (acc * np.sqrt(3.0 / 4.0)).astype(np.float32)


def randint(name, shape, lo, hi, salt=0):
    """int64 in [lo, hi)."""
    u = uniform(name, shape, 0.0, 1.0, salt=salt + 77).astype(np.float64)
    return np.minimum((lo + np.floor(u * (hi - lo))).astype(np.int64), hi - 1)


def fill_state_dict(sd, tag="w"):
    """Overwrite every float tensor of a torch state_dict in place, keyed by parameter name.

    Scale: weights with >=2 dims get U(-1,1)*sqrt(3/fan_in) (unit-variance preserving),
    BatchNorm weight ~ 1 + 0.1*U, biases 0.05*U, running_mean 0, running_var 1.
    """
    import torch
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue
        shp = tuple(v.shape)
        if k.endswith("running_mean"):
            arr = np.zeros(shp, np.float32)
        elif k.endswith("running_var"):
            arr = np.ones(shp, np.float32)
        elif v.dim() >= 2:
            fan_in = int(np.prod(shp[1:]))
            arr = uniform(tag + ":" + k, shp) * np.float32(np.sqrt(3.0 / fan_in))
        elif k.endswith("bias"):
            arr = uniform(tag + ":" + k, shp) * np.float32(0.05)
        else:  # 1-D weight == BatchNorm gamma
            arr = 1.0 + 0.1 * uniform(tag + ":" + k, shp)
        v.copy_(torch.from_numpy(np.ascontiguousarray(arr)).reshape(v.shape))
    return sd
