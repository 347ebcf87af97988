"""pdm-b200: B200-native joint image + panoptic-mask diffusion sampling (U-ViT t2i x DPM-Solver++).

Python host layer over libpdm.so (hand-written sm_100a CUDA behind the C ABI in include/pdm.h).
"""
__version__ = "0.1.0"
