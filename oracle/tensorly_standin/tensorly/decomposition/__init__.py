"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
from ._cp import parafac  # noqa: F401
