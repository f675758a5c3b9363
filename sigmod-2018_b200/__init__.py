"""sigmod-2018_b200 — B200-native join hot path of VagelisN/Sigmod-2018.

The product is `lib/libb200join.so` (hand-written CUDA for sm_100a behind a
C-ABI, `include/b200_join.h`).  This package is only the ctypes binding tests
and `bench.py` use; it mirrors the reference's operator API (same names,
argument meaning and NULL/"empty" conventions, see `host.py`).

There is no CPU fallback: importing `host` without the built library raises.
The directory name contains a dash, so load it with
`importlib` (see `tests/conftest.py`: it is registered as `sigmod2018_b200`).
"""
from . import host, sharding  # noqa: F401
from .host import *  # noqa: F401,F403
