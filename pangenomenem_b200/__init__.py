"""pangenomenem_b200 -- B200-native NEM partitioning engine behind the reference's nem() boundary.

``capi``   ctypes mirror of include/nem_b200.h (no CPU fallback: raises without the CUDA library)
``synth``  seeded synthetic pangenomes + the NEM file contract of ppanggolin.py
``build``  in-tree nvcc/gcc build of libnem_b200.so
"""
__all__ = ["capi", "synth", "build"]
