"""locator_b200: B200-native training / prediction hot path of kr-colab/locator.

Importing the package does not touch CUDA; the compute modules (`genotypes`, `model`,
`locator`) load liblocator_b200.so on first import and fail loudly if it is missing.
"""
__version__ = "0.1.0"
