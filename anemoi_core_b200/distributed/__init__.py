"""Sharding metadata and the integer (index) path of the hot path: balanced dst-range partitions, dst-sorting,
chunk materialisation.  Mirrors ``anemoi.models.distributed`` (shapes, balanced_partition, khop_edges)."""
