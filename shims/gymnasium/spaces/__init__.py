"""gymnasium.spaces stand-in: Box and Dict carriers (shape / bounds / dtype; no sampling machinery beyond uniform)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = (tuple(shape) if shape is not None else None), dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(shape, dtype)
        self.low = np.full(self.shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Dict(Space, dict):
    def __init__(self, spaces=None, **kw):
        dict.__init__(self, spaces or {}, **kw)
        Space.__init__(self, None, None)

    @property
    def spaces(self):
        return self
