"""Drop-in for dysgu.scikitbio (reference: dysgu/scikitbio/__init__.py:10-11)."""
from ._ssw_wrapper import StripedSmithWaterman, AlignmentStructure

__all__ = ["StripedSmithWaterman", "AlignmentStructure"]
