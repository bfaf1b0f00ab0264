"""ORACLE shim standing in for the third-party ``auraloss==0.4.0`` package (see freq.py
header).  Test infrastructure only."""
from . import freq  # noqa: F401
