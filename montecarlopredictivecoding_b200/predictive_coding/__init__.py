"""Drop-in ``predictive_coding`` package (reference: predictive_coding/__init__.py:1-2)."""
from .layer import PCLayer
from .trainer import PCTrainer

__all__ = ["PCLayer", "PCTrainer"]
