"""Data path of the training loop."""
from .loader import DataLoader  # noqa: F401
