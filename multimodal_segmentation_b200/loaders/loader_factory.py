"""Factory for data loaders (reference: loaders/loader_factory.py:4-10 -- the documented extension
point for new datasets, README.md:23)."""
from .synthetic_chaos import SyntheticChaosLoader


def init_loader(dataset):
    """'chaos' resolves to the synthetic CHAOS-shaped loader: the CHAOS MR DICOM volumes (and the
    dicom / skimage readers of loaders/chaos.py) are not available in this environment."""
    if dataset in ("chaos", "synthetic_chaos"):
        return SyntheticChaosLoader()
    return None
