"""deepdenoiser_b200 - B200 (sm_100a) native drop-in for the DeepDenoiser conv + kernel-prediction hot path.

Host side mirrors the reference's Python interface (Architecture / RenderPasses / Naming ...); all
arithmetic runs in libdd_b200.so through the C ABI declared in include/dd_b200.h.
"""
__version__ = "0.1.0"
