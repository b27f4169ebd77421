"""Data package (reference src/data/__init__.py:4)."""
from .dataset import MultiModalDataset

__all__ = ['MultiModalDataset']
