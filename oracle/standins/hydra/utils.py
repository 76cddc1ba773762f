import functools
import importlib

from hydra.errors import InstantiationException


def get_class(path: str):
    mod, _, name = path.rpartition(".")
    try:
        return getattr(importlib.import_module(mod), name)
    except Exception as e:  # noqa: BLE001
        raise InstantiationException(str(e)) from e


def instantiate(config, *args, _partial_=None, _recursive_=True, _convert_=None, **kwargs):
    cfg = dict(config)
    target = cfg.pop("_target_")
    partial = cfg.pop("_partial_", False) if _partial_ is None else (cfg.pop("_partial_", None), _partial_)[1]
    cfg.pop("_recursive_", None)
    cfg.pop("_convert_", None)
    cls = get_class(target) if isinstance(target, str) else target
    cfg.update(kwargs)
    if partial:
        return functools.partial(cls, *args, **cfg)
    try:
        return cls(*args, **cfg)
    except Exception as e:  # noqa: BLE001
        raise InstantiationException(f"Error instantiating {target}: {e}") from e
