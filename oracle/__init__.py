"""CPU oracles (test infrastructure only).  See the header of each module.

Nothing under ``gptest_b200/`` may import this package: the product path is CUDA-only and
fails loudly when its extension is missing.
"""
