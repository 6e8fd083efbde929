"""Stand-in for the reference package `environments/` (hot-path modules only)."""
