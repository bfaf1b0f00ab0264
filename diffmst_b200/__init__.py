"""diffmst_b200 — B200-native (sm_100a) implementation of the Diff-MST mix-console and
loss hot path behind the reference's Python call signatures.  See DESIGN.md."""
from .console import AdvancedMixConsole, BasicMixConsole  # noqa: F401

__all__ = ["AdvancedMixConsole", "BasicMixConsole"]
