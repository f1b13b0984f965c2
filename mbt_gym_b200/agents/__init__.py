"""Closed-form baseline agents (callers of the hot path) and their on-device policy forms."""
