"""vjf_b200: B200-native (sm_100a) implementation of the VJF filter + learning step.

Public surface mirrors the reference package for this path: ``from vjf_b200.model import VJF``.
"""
from .model import VJF, Gaussian  # noqa: F401

__all__ = ["VJF", "Gaussian"]
