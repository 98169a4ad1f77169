"""Import shim for the absent rapidfuzz (see oracle/reference_port.py: parity unpinned)."""
from . import fuzz  # noqa: F401
