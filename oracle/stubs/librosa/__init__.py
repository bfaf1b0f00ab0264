"""Empty stand-in so that the reference's ``mst/loss.py:3`` (``import librosa``) imports
in this container; only ``compute_melspectrum`` (unused on the hot path) touches it."""
