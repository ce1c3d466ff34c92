"""amuse_b200 -- B200-native engine for AMUSE's gesture-sampling hot path.

Host side (Python, mirrors the reference's Python seam) over a C-ABI CUDA library
(``amuse_b200/lib/libamuse_b200.so``, include/amuse_b200.h).  There is no CPU fallback:
importing the engine without the built library, or creating it without a B200, raises.

    from amuse_b200.infer_ldm import PretrainedLPDM_v1      # drop-in for the reference class
    from amuse_b200.engine import Engine                    # thin tensor-level wrapper
"""
__all__ = ["engine", "infer_ldm", "_lib"]
__version__ = "0.1"
