"""Stand-in for hydra (test infrastructure only)."""
