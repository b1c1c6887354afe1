"""Drop-in for DreamScene's `from simple_knn._C import distCUDA2` (/root/reference/gs_renderer.py:9)."""
