"""Stand-in for torch_geometric (test infrastructure only; see ../README.md)."""
__version__ = "2.6.1"
