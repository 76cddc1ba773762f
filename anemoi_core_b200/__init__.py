"""anemoi_core_b200 — B200-native (sm_100a) forward pass of anemoi-models' graph message-passing stack.

Public API mirrors ``anemoi.models.layers`` / ``anemoi.models.distributed`` for the hot path:
``layers.processor.{GNNProcessor, GraphTransformerProcessor}``, ``layers.mapper.*Mapper``, ``layers.block.*Block``,
``layers.conv.{GraphConv, GraphTransformerConv}``, ``distributed.shapes.{GraphShardInfo, BipartiteGraphShardInfo}``.
All numerics run in ``lib/libanemoi_b200.so`` (C ABI in ``include/anemoi_b200.h``); there is no CPU fallback.
"""

__version__ = "0.1.0"
