"""ORACLE shim standing in for the third-party ``dasp-pytorch==0.0.1`` package (see
functional.py / signal.py headers).  Test infrastructure only."""
from . import functional, signal  # noqa: F401
