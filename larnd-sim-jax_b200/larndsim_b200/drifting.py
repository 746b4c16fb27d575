"""Mirror of the reference's ``larndsim.drifting_jax``: ``drift(params, tracks, fields)`` (drifting_jax.py:19-58)."""
from .stream_ops import drift  # noqa: F401
