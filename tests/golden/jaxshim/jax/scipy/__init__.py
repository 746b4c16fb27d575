from . import special, stats  # noqa: F401
