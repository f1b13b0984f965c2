"""Rollout helpers: generate_trajectory (step by step and fused) and the results table."""
