"""Make the reference's dotted class paths resolve to the B200 classes ("_target_ whitelist").

anemoi-training instantiates the hot-path modules through Hydra from literal ``_target_`` strings, and with ``config_validation: True``
pydantic only accepts the reference's own paths (``schemas/processor.py:29,38``, ``encoder.py:25,30``, ``decoder.py:25,30``):

    anemoi.models.layers.processor.{GNNProcessor, GraphTransformerProcessor}
    anemoi.models.layers.mapper.{GNNForwardMapper, GNNBackwardMapper, GraphTransformerForwardMapper, GraphTransformerBackwardMapper}
    anemoi.models.layers.block.{GraphConvProcessorBlock, GraphConvMapperBlock, GraphTransformerProcessorBlock, GraphTransformerMapperBlock}
    anemoi.models.layers.conv.{GraphConv, GraphTransformerConv}

``install()`` rebinds exactly those names: if anemoi-models is importable its modules are patched in place (everything else in them stays
the reference's), otherwise stub modules with those names are registered in ``sys.modules``.  A config written for the reference then
builds this package's modules unchanged, validated or not; ``uninstall()`` restores what was there.  Checkpoints of whole pickled
modules (training/utils/checkpoint.py:84-107) resolve the same way.
"""

from __future__ import annotations

import importlib
import sys
import types

_TARGETS = {
    "anemoi.models.layers.processor": ("GNNProcessor", "GraphTransformerProcessor"),
    "anemoi.models.layers.mapper": ("GNNForwardMapper", "GNNBackwardMapper", "GraphTransformerForwardMapper", "GraphTransformerBackwardMapper"),
    "anemoi.models.layers.block": ("GraphConvProcessorBlock", "GraphConvMapperBlock", "GraphTransformerProcessorBlock", "GraphTransformerMapperBlock"),
    "anemoi.models.layers.conv": ("GraphConv", "GraphTransformerConv"),
}
_SAVED: list = []  # (module, name, previous attribute | _MISSING) and ("module", dotted name) for stubs we created
_MISSING = object()


def targets() -> list[str]:
    """The dotted ``_target_`` strings ``install()`` takes over."""
    return [f"{mod}.{name}" for mod, names in _TARGETS.items() for name in names]


def _ensure_module(dotted: str) -> types.ModuleType:
    """Import ``dotted`` if the reference is installed, else create (and register) empty stub packages down to it."""
    try:
        return importlib.import_module(dotted)
    except Exception:  # noqa: BLE001 - reference absent, or present without its dependencies (torch_geometric, hydra): stub it
        pass
    parts = dotted.split(".")
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []  # a package, so that sub-modules can hang off it
            m.__anemoi_b200_stub__ = True
            sys.modules[name] = m
            _SAVED.append(("module", name))
            if i > 1:
                setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)
    return sys.modules[dotted]


def install() -> list[str]:
    """Rebind the reference's hot-path class paths to this package's classes; returns the dotted names now served from here."""
    from . import layers

    if _SAVED:
        return targets()
    for dotted, names in _TARGETS.items():
        mod = _ensure_module(dotted)
        for name in names:
            _SAVED.append((mod, name, getattr(mod, name, _MISSING)))
            setattr(mod, name, getattr(layers, name))
    return targets()


def uninstall() -> None:
    while _SAVED:
        item = _SAVED.pop()
        if item[0] == "module":
            sys.modules.pop(item[1], None)
        else:
            mod, name, prev = item
            if prev is _MISSING:
                delattr(mod, name)
            else:
                setattr(mod, name, prev)
