"""Mirror of the reference's ``larndsim.quenching_jax``: ``quench(params, tracks, fields)`` (quenching_jax.py:38-75)."""
from .stream_ops import quench  # noqa: F401
