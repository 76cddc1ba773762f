"""oracle/check_signatures.py — the drop-in classes' call surface against the UNMODIFIED reference classes.

*** TEST INFRASTRUCTURE (run by tests/test_host_logic.py in a subprocess).  Never imported by anemoi_core_b200. ***

For both processors, the four mappers, the four blocks, the two conv operators, MLP and the two LayerNorm kernels: every constructor parameter of
the reference is accepted (by name or through ``**kwargs``), every default the reference declares is the effective default here (walking the MRO:
sub-classes forward ``**kwargs`` to a base constructor), and ``forward`` takes the reference's parameters in the reference's positional order.
    python oracle/check_signatures.py   ->   one JSON line {"classes": n, "problems": [...]}"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_step import reference_root  # noqa: E402


def _effective_defaults(cls) -> dict:
    out = {}
    for k in reversed(cls.__mro__):
        if "__init__" in k.__dict__:
            for n, p in inspect.signature(k.__init__).parameters.items():
                if p.default is not inspect.Parameter.empty:
                    out[n] = p.default
    return out


def _accepts(cls, name: str) -> bool:
    for k in cls.__mro__:
        if "__init__" in k.__dict__:
            ps = inspect.signature(k.__init__).parameters
            if name in ps:
                return True
            if not any(p.kind == p.VAR_KEYWORD for p in ps.values()):
                return False
    return False


def main() -> dict:
    root = reference_root()
    if root is None:
        return {"unavailable": "reference not found (neither /root/reference/models/src nor baseline/_ref)"}
    for p in (root, os.path.join(HERE, "standins")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from anemoi.models.layers import block as RB
    from anemoi.models.layers import conv as RC
    from anemoi.models.layers import mapper as RM
    from anemoi.models.layers import mlp as RL
    from anemoi.models.layers import normalization as RN
    from anemoi.models.layers import processor as RP

    from anemoi_core_b200.layers import block as B
    from anemoi_core_b200.layers import conv as C
    from anemoi_core_b200.layers import mapper as M
    from anemoi_core_b200.layers import mlp as L
    from anemoi_core_b200.layers import normalization as N
    from anemoi_core_b200.layers import processor as P

    groups = ((RP, P, ["GNNProcessor", "GraphTransformerProcessor"]),
              (RM, M, ["GNNForwardMapper", "GNNBackwardMapper", "GraphTransformerForwardMapper", "GraphTransformerBackwardMapper"]),
              (RB, B, ["GraphConvProcessorBlock", "GraphConvMapperBlock", "GraphTransformerProcessorBlock", "GraphTransformerMapperBlock"]),
              (RC, C, ["GraphConv", "GraphTransformerConv"]), (RL, L, ["MLP"]), (RN, N, ["ConditionalLayerNorm", "AutocastLayerNorm"]))  # fmt: skip
    problems, n = [], 0
    for rmod, omod, names in groups:
        for name in names:
            n += 1
            rc, oc = getattr(rmod, name), getattr(omod, name)
            rp = inspect.signature(rc.__init__).parameters
            for k, p in rp.items():
                if k == "self" or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
                    continue
                if not _accepts(oc, k):
                    problems.append(f"{name}.__init__: parameter {k!r} is not accepted")
            eff = _effective_defaults(oc)
            for k, p in rp.items():
                if p.default is not inspect.Parameter.empty and k in eff and eff[k] != p.default:
                    problems.append(f"{name}.__init__: default of {k!r} is {eff[k]!r}, reference {p.default!r}")
            rf, of = inspect.signature(rc.forward).parameters, inspect.signature(oc.forward).parameters
            r_order = [k for k, p in rf.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
            o_order = [k for k, p in of.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
            if o_order[: len(r_order)] != r_order:
                problems.append(f"{name}.forward: positional parameters {o_order} vs reference {r_order}")
            for k, p in rf.items():
                if p.default is not inspect.Parameter.empty and k in of and of[k].default != p.default:
                    problems.append(f"{name}.forward: default of {k!r} is {of[k].default!r}, reference {p.default!r}")
    return {"classes": n, "problems": problems, "reference": root}


if __name__ == "__main__":
    print(json.dumps(main()))
