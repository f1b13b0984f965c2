"""Reward-function descriptors; the arithmetic runs in the fused CUDA kernels."""
