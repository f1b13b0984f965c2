"""TEST INFRASTRUCTURE: CPU oracle, reference shim, NumPy port (never imported by mbt_gym_b200)."""
