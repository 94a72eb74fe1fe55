"""Limiter ids (pyclaw/src/pyclaw/limiters/tvd.py:74-77)."""
from . import tvd  # noqa: F401
