"""botocore stub (reference file_utils.py:34)."""
from . import exceptions  # noqa: F401
