"""b200taco -- B200-native Tacotron 2 hot path behind the reference's model API.

``from b200taco.model import Tacotron2`` replaces ``from src.model import Tacotron2``
(see INTEGRATION.md).  Importing this package never touches CUDA or the shared library;
the library is loaded on the first forward/inference call and its absence is an error.
"""
from . import config  # noqa: F401

__all__ = ["config"]
