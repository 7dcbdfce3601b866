"""Stands where the reference's compiled ``simple_knn._C`` module would be (ext.cpp:15-17 binds distCUDA2)."""
from gaussianip_b200.knn import distCUDA2  # noqa: F401

__all__ = ["distCUDA2"]
