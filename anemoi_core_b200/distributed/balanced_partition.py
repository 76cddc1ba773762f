"""Balanced contiguous partitions (reference: distributed/balanced_partition.py:16-41, 44-73, 76-93)."""

from __future__ import annotations


def get_balanced_partition_sizes(total_size: int, n_partitions: int) -> list[int]:
    """Sizes of ``n_partitions`` contiguous parts of ``total_size`` items; the first ``total_size % n`` parts
    carry one extra item (balanced_partition.py:16-41)."""
    if n_partitions <= 0:
        raise ValueError(f"n_partitions must be positive, got {n_partitions}")
    if total_size < 0:
        raise ValueError(f"total_size must be non-negative, got {total_size}")
    q, r = divmod(total_size, n_partitions)
    return [q + (1 if i < r else 0) for i in range(n_partitions)]


def get_partition_range(partition_sizes: list[int], partition_id: int) -> tuple[int, int]:
    """[start, end) of part ``partition_id`` given the part sizes."""
    if not 0 <= partition_id < len(partition_sizes):
        raise IndexError(f"partition_id {partition_id} out of range for {len(partition_sizes)} partitions")
    start = sum(partition_sizes[:partition_id])
    return start, start + partition_sizes[partition_id]


def get_balanced_partition_range(total_size: int, n_partitions: int, partition_id: int) -> tuple[int, int]:
    return get_partition_range(get_balanced_partition_sizes(total_size, n_partitions), partition_id)
