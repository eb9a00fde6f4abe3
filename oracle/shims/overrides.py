"""Shim for the third-party `overrides` package (absent in this image).

The reference uses it only as a decorator (models/diffusion.py:9,158), so an
identity decorator is behaviour-preserving.  Test infrastructure only.
"""


def override(f):
    return f
