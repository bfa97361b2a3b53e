"""Small shared constants (``stgraph/utils/constants.py``)."""
from .constants import SizeConstants

__all__ = ["SizeConstants"]
