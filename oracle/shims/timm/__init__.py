"""Test-infrastructure shim for timm==0.6.12 (absent in this image).

Only the three names the reference imports (simplified_attention.py:9) are restated,
from timm's published semantics. Used ONLY by oracle/make_golden.py to import the
reference in the build container; never by the product path.
"""
