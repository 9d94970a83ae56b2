"""timm.layers.set_fast_norm (benchmark.py:22, only called with --fast-norm)."""
_FAST_NORM = False


def set_fast_norm(enable=True):
    global _FAST_NORM
    _FAST_NORM = enable
