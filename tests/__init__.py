"""Test suite: `-m "not gpu"` runs on CPU, `-m gpu` on a B200."""
