from .conv import FlashFFTConv  # noqa: F401  (reference flashfftconv/__init__.py:1)
