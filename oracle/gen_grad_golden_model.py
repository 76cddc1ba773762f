"""oracle/gen_grad_golden_model.py — gradient fixture of the UNMODIFIED reference ``AnemoiModelEncProcDec`` (PyTorch autograd on CPU, fp32).

*** TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference). ***     python oracle/gen_grad_golden_model.py

The reference model of ``tests/golden/model_forward.pt`` (same graph, same ``state_dict``; built like ``oracle/gen_golden.py::model_cases``: the
reference ``forward`` / ``_assemble_input`` / ``_assemble_output`` on reference graph providers, node attributes, mappers, processor, SkipConnection
residual and ReluBounding) is run in TRAINING mode for both model kinds: loss = sum(y * w) with a seeded cotangent, gradients of every parameter —
including the trainable node and edge tensors — and of the input.  -> tests/golden/grads_model.pt
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import gen_golden as G  # noqa: E402  (sets up the reference import paths; nothing runs at import)
import torch  # noqa: E402


def main():
    from torch_geometric.data import HeteroData

    from anemoi.models.layers.bounding import ReluBounding
    from anemoi.models.layers.graph import NamedNodesAttributes
    from anemoi.models.layers.graph_provider import create_graph_provider

    fx = torch.load(os.path.join(ROOT, "tests", "golden", "model_forward.pt"), weights_only=False)
    d = fx["dims"]
    Model = G._load_reference_model_class()
    graph = HeteroData()
    graph["data"].x, graph["hidden"].x = fx["coords"]["data"], fx["coords"]["hidden"]
    subs = {}
    for name, (src, dst) in (("enc", ("data", "hidden")), ("proc", ("hidden", "hidden")), ("dec", ("hidden", "data"))):
        st = graph[(src, "to", dst)]
        st.edge_index, st.edge_length, st.edge_dirs = (fx["graph"][name][k] for k in ("edge_index", "edge_length", "edge_dirs"))
        subs[name] = st
    n_data, n_hid, C, heads = d["n_data"], d["n_hid"], d["C"], d["heads"]
    out = {"kind": "model_grads", "cases": {}}
    for kind in ("graphtransformer", "gnn"):
        m = Model.__new__(Model)
        torch.nn.Module.__init__(m)
        m._graph_name_hidden, m.n_step_output, m.latent_skip = "hidden", d["t_out"], True
        m._internal_input_idx, m._internal_output_idx = {"data": fx["in_prog"]}, {"data": fx["out_prog"]}
        m.node_attributes = NamedNodesAttributes({"data": 0, "hidden": 3}, graph)
        prov = {k: create_graph_provider(graph=v, edge_attributes=["edge_length", "edge_dirs"], src_size=s, dst_size=t, trainable_size=2)
                for (k, v), (s, t) in zip(subs.items(), ((n_data, n_hid), (n_hid, n_hid), (n_hid, n_data)))}  # fmt: skip
        m.encoder_graph_provider = torch.nn.ModuleDict({"data": prov["enc"]})
        m.processor_graph_provider = prov["proc"]
        m.decoder_graph_provider = torch.nn.ModuleDict({"data": prov["dec"]})
        in_dim, lat_dim, edge_dim = fx["cases"][kind]["in_dim"], fx["cases"][kind]["lat_dim"], fx["cases"][kind]["edge_dim"]
        if kind == "graphtransformer":
            kw = dict(num_heads=heads, mlp_hidden_ratio=4, edge_dim=edge_dim, layer_kernels=None, graph_attention_backend="pyg", num_chunks=1)
            enc = G.GraphTransformerForwardMapper(in_channels_src=in_dim, in_channels_dst=lat_dim, hidden_dim=C, **kw)
            proc = G.GraphTransformerProcessor(num_layers=2, num_channels=C, **kw)
            dec = G.GraphTransformerBackwardMapper(in_channels_src=C, in_channels_dst=in_dim, hidden_dim=C, out_channels_dst=d["t_out"] * d["n_out"], **kw)
        else:
            kw = dict(mlp_extra_layers=0, edge_dim=edge_dim, layer_kernels=None, num_chunks=1)
            enc = G.GNNForwardMapper(in_channels_src=in_dim, in_channels_dst=lat_dim, hidden_dim=C, **kw)
            proc = G.GNNProcessor(num_layers=2, num_channels=C, **kw)
            dec = G.GNNBackwardMapper(in_channels_src=C, in_channels_dst=in_dim, hidden_dim=C, out_channels_dst=d["t_out"] * d["n_out"], **kw)
        m.encoder, m.processor, m.decoder = torch.nn.ModuleDict({"data": enc}), proc, torch.nn.ModuleDict({"data": dec})
        m.residual = torch.nn.ModuleDict({"data": G._SkipConnection(step=-1)})
        name_to_index = {f"v{i}": i for i in range(d["n_out"])}
        names = [f"v{i}" for i in fx["bound_vars"]]
        m.boundings = torch.nn.ModuleDict({"data": torch.nn.ModuleList([ReluBounding(variables=names, name_to_index=name_to_index)])})
        m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
        m.eval()
        with torch.no_grad():
            assert torch.equal(m({"data": fx["x"].clone()})["data"], fx["cases"][kind]["y"])  # the very model of model_forward.pt
        m.train()
        x = fx["x"].clone().requires_grad_()
        y = m({"data": x})["data"]
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(77))
        (y * w).sum().backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        assert len(grads) == len(list(m.parameters())), "a parameter without gradient"
        out["cases"][kind] = {"w": w, "y": y.detach().clone(), "x_grad": x.grad.clone(), "grads": grads}
        print("model grads", kind, len(grads), float(x.grad.abs().mean()))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "grads_model.pt"))


if __name__ == "__main__":
    main()
