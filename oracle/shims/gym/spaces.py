import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            shape = low.shape
        else:
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        Space.__init__(self, shape, dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        Space.__init__(self, (), np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n
