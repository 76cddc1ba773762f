class _Store(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class HeteroData(dict):
    """Dict-of-stores stand-in: data["name"].x, data[("a","to","b")].edge_index ..."""

    def __getitem__(self, key):
        if key not in self:
            super().__setitem__(key, _Store())
        return super().__getitem__(key)

    def to(self, *a, **k):
        return self
