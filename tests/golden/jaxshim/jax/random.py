"""jax.random stand-in: threefry keys / split / normal through oracle/jax_random.py (the repo's restatement of
jax/_src/prng.py, pinned to published JAX values in tests/test_oracle_golden.py).  jax_threefry_partitionable=True,
the default since JAX 0.5.0."""
import os
import sys

import numpy as _np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..")))
from oracle import jax_random as _jr  # noqa: E402

from ._core import FLOAT, wrap  # noqa: E402


class Key(tuple):
    pass


def key(seed):
    return Key(_jr.key(int(seed)))


PRNGKey = key


def split(k, num=2):
    return [Key(x) for x in _jr.split(tuple(k), int(num))]


def normal(k, shape=(), dtype=None):
    return wrap(_jr.normal(tuple(k), tuple(shape)).astype(FLOAT))
