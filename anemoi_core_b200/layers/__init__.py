"""Host-side mirror of ``anemoi.models.layers`` for the graph message-passing hot path: same class names,
constructor kwargs, forward signatures and ``state_dict`` keys as the reference; every forward runs on
libanemoi_b200.so (sm_100a)."""
from .block import GraphConvMapperBlock
from .block import GraphConvProcessorBlock
from .block import GraphTransformerMapperBlock
from .block import GraphTransformerProcessorBlock
from .conv import GraphConv
from .conv import GraphTransformerConv
from .graph import NamedNodesAttributes
from .graph import TrainableTensor
from .graph_provider import NoOpGraphProvider
from .graph_provider import StaticGraphProvider
from .graph_provider import create_graph_provider
from .mapper import GNNBackwardMapper
from .mapper import GNNForwardMapper
from .mapper import GraphTransformerBackwardMapper
from .mapper import GraphTransformerForwardMapper
from .mlp import MLP
from .processor import GNNProcessor
from .processor import GraphTransformerProcessor

__all__ = [
    "MLP",
    "GraphConv",
    "GraphTransformerConv",
    "GraphConvProcessorBlock",
    "GraphConvMapperBlock",
    "GraphTransformerProcessorBlock",
    "GraphTransformerMapperBlock",
    "GNNProcessor",
    "GraphTransformerProcessor",
    "GNNForwardMapper",
    "GNNBackwardMapper",
    "GraphTransformerForwardMapper",
    "GraphTransformerBackwardMapper",
    "TrainableTensor",
    "NamedNodesAttributes",
    "StaticGraphProvider",
    "NoOpGraphProvider",
    "create_graph_provider",
]
