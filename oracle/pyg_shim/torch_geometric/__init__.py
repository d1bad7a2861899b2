"""Minimal stand-in for the PyTorch-Geometric 1.7.2 symbols GLASS imports.

TEST INFRASTRUCTURE ONLY.  The reference (/root/reference, README.md:19) pins
PyG 1.7.2, which is not installable offline.  This package defines exactly the
symbols the GLASS path imports so the *unmodified*
reference can be imported in the build container to generate golden vectors.
Semantics follow the PyG 1.7.2 documentation / GraphNorm paper; the PyG source
is not available offline, so parity at this boundary is "unpinned" (DESIGN.md).
"""
__version__ = "1.7.2-standin"
