"""basi_b200: B200-native hot path of alisure-ml/Instance-Segment-BASI (click-conditioned PSPNet)."""
from ._lib import BasiError  # noqa: F401

__all__ = ["BAISData", "BAISPSPNet", "BAISRunnerTrain", "BAISRunnerOne", "BAISTools", "engine", "dp"]
