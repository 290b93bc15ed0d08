"""graphnics_b200 -- B200-native drop-in for the network-flow hot path of graphnics.

Same flat namespace as graphnics/__init__.py:1-5 (`from graphnics_b200 import *`).
"""
from ._lib import GnxError, LIB_PATH  # noqa: F401
