"""diffmst_b200 — B200-native (sm_100a) implementation of the Diff-MST mix-console and
loss hot path behind the reference's Python call signatures.  See DESIGN.md."""
from .console import AdvancedMixConsole, BasicMixConsole  # noqa: F401
from .conv import Cnn14, ConvBlock, SpectrogramEncoder  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .inference import run_diffmst, sliding_window_mix  # noqa: F401
from .mixing import naive_random_mix, random_reference_mix  # noqa: F401
from .losses import (AudioFeatureLoss, MRSTFTLoss, MultiResolutionSTFTLoss,  # noqa: F401
                     batch_stereo_peak_normalize)

__all__ = ["AdvancedMixConsole", "BasicMixConsole", "MultiResolutionSTFTLoss", "MRSTFTLoss",
           "AudioFeatureLoss", "batch_stereo_peak_normalize", "ConvBlock", "Cnn14", "SpectrogramEncoder", "naive_random_mix",
           "random_reference_mix", "GraphedStep", "run_diffmst", "sliding_window_mix"]
