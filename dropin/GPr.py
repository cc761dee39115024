"""Shim: `import GPr` resolves to the B200 implementation."""
from gptest_b200.GPr import *  # noqa: F401,F403
from gptest_b200 import GPr as _impl
__all__ = [n for n in dir(_impl) if not n.startswith('_')]
