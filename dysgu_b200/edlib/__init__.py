"""Drop-in for dysgu.edlib (reference: dysgu/edlib/__init__.py:2)."""
from .edlib import align, getNiceAlignment

__all__ = ["align", "getNiceAlignment"]
