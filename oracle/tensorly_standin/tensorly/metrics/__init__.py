"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
