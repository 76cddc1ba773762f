"""Model glue either side of the step: ``AnemoiModelEncProcDec`` against golden outputs of the UNMODIFIED reference
``forward`` / ``_assemble_input`` / ``_assemble_output`` (``oracle/gen_golden.py::model_cases`` -> tests/golden/model_forward.pt), the
oracle restatement against the same goldens (CPU), and the two glue kernels against the oracle on ragged shapes (GPU)."""
import pytest
import torch

from anemoi_core_b200.model import AnemoiModelEncProcDec
from oracle import restatement as R


def build_model(fx, kind):
    d = fx["dims"]
    graph = {"data": {"x": fx["coords"]["data"]}, "hidden": {"x": fx["coords"]["hidden"]},
             ("data", "to", "hidden"): fx["graph"]["enc"], ("hidden", "to", "hidden"): fx["graph"]["proc"], ("hidden", "to", "data"): fx["graph"]["dec"]}  # fmt: skip
    if kind == "graphtransformer":
        common = dict(num_heads=d["heads"], mlp_hidden_ratio=4, num_chunks=1)
        enc, proc, dec = common, dict(num_layers=2, **common), common
    else:
        common = dict(mlp_extra_layers=0, num_chunks=1)
        enc, proc, dec = common, dict(num_layers=2, **common), common
    return AnemoiModelEncProcDec(kind, graph_data=graph, edge_attributes=["edge_length", "edge_dirs"], num_channels=d["C"], n_step_input=d["t_in"],
                                 n_step_output=d["t_out"], num_input_channels={"data": d["n_in"]}, num_output_channels={"data": d["n_out"]},
                                 internal_input_idx={"data": fx["in_prog"]}, internal_output_idx={"data": fx["out_prog"]}, encoder=enc, processor=proc,
                                 decoder=dec, trainable_parameters={"hidden": 3, "data2hidden": 2, "hidden2hidden": 2, "hidden2data": 2},
                                 boundings={"data": [("relu", fx["bound_vars"])]}).eval()  # fmt: skip


@pytest.mark.parametrize("kind", ["graphtransformer", "gnn"])
def test_oracle_model_forward_matches_reference(golden, kind):
    fx = golden("model_forward")
    y = R.anemoi_model_forward(kind, fx["cases"][kind]["sd"], fx)
    ref = fx["cases"][kind]["y"]
    assert y.shape == ref.shape == (fx["dims"]["batch"], fx["dims"]["t_out"], 1, fx["dims"]["n_data"], fx["dims"]["n_out"])
    assert (y - ref).abs().max() <= 2e-5 * ref.abs().max()
    assert (y[..., fx["bound_vars"]] >= 0).all()


@pytest.mark.parametrize("kind", ["graphtransformer", "gnn"])
def test_model_state_dict_is_the_reference_one(golden, kind):
    """Same keys and shapes as the reference model: its checkpoint loads with strict=True (no GPU needed to build the module)."""
    fx = golden("model_forward")
    m = build_model(fx, kind)
    ref_sd = fx["cases"][kind]["sd"]
    assert sorted(m.state_dict().keys()) == sorted(ref_sd.keys())
    # ... in the reference's ORDER: parameters() enumerates like the reference's, which is what an optimizer state_dict indexes by
    assert list(m.state_dict().keys()) == list(ref_sd.keys())
    missing, unexpected = m.load_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected
    assert m.input_dim["data"] == fx["cases"][kind]["in_dim"] and m.input_dim_latent == fx["cases"][kind]["lat_dim"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # the product path never computes on the CPU
        with torch.no_grad():
            m({"data": fx["x"]})


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["graphtransformer", "gnn"])
def test_model_forward_matches_reference_golden(golden, kind):
    fx = golden("model_forward")
    m = build_model(fx, kind)
    m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
    m = m.cuda()
    ref = fx["cases"][kind]["y"]
    x = {"data": fx["x"].cuda()}
    with torch.no_grad():
        y32 = m(x)["data"]
        y32_again = m(x)["data"]  # second call runs on the cached edge tensors / CSR plans
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(x)["data"]
    assert y32.shape == ref.shape and y32.dtype == torch.float32
    # fp32: 1e-4 of the output scale (the tolerance BASELINE.json's north_star states); bf16: 2e-2 rel-L2 against the fp32 reference
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert torch.equal(y32, y32_again)
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2
    assert (y32[..., fx["bound_vars"]] >= 0).all() and (y16[..., fx["bound_vars"]] >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 2, 1, 37, 5), (2, 3, 2, 64, 33), (1, 1, 1, 1, 1), (3, 2, 1, 0, 4)])
def test_assemble_input_kernel(shape, dt):
    from anemoi_core_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, T, E, G, V = shape
    x = torch.randn(shape, generator=g)
    attrs = torch.randn(B * E * G, 6, generator=g)
    ref = R.assemble_input(x, attrs)
    for k_pad in (None, (T * V + 6 + 7) // 8 * 8 + 8):
        out = ops.assemble_input(x.cuda(), attrs.cuda(), dt, k_pad=k_pad)
        K = T * V + 6
        assert out.shape == (B * E * G, K if k_pad is None else k_pad) and out.dtype == dt
        if G == 0:
            continue
        assert torch.equal(out[:, :K].float().cpu(), ref.to(dt).float())  # pure data movement + one rounding: exact
        assert (out[:, K:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_assemble_output_kernel(dt):
    from anemoi_core_b200 import ops

    g = torch.Generator().manual_seed(6)
    B, T_in, E, G, V_in, T_out, V_out = 2, 3, 2, 41, 9, 2, 7
    x = torch.randn(B, T_in, E, G, V_in, generator=g)
    dec = torch.randn(B * E * G, T_out * V_out, generator=g).to(dt)
    in_prog, out_prog, relu_vars, leaky_vars = [0, 2, 3, 8], [1, 2, 4, 6], [2, 5], [0]
    ref = R.assemble_output(dec.float(), x, B, E, T_out, in_prog, out_prog, relu_vars, step=1)
    ref[..., leaky_vars] = torch.nn.functional.leaky_relu(ref[..., leaky_vars])
    skip = torch.full((V_out,), -1, dtype=torch.int32)
    skip[out_prog] = torch.tensor(in_prog, dtype=torch.int32)
    bound = torch.zeros(V_out, dtype=torch.int32)
    bound[relu_vars], bound[leaky_vars] = 1, 2
    y = ops.assemble_output(dec.cuda(), x.cuda(), B, E, T_out, 1, skip.cuda(), bound.cuda())
    assert y.shape == ref.shape and torch.allclose(y.cpu(), ref, rtol=0, atol=1e-6)
    # no residual, no bounding: a pure rearrange
    y0 = ops.assemble_output(dec.cuda(), None, B, E, T_out)
    assert torch.equal(y0.cpu(), dec.float().reshape(B, E, G, T_out, V_out).permute(0, 3, 1, 2, 4))


def test_from_reference_config(golden):
    """The model built from reference-shaped config objects (DotDict-like config, IndexCollection-like data indices) is the one built by hand:
    same state_dict keys / shapes as the reference model, same index tables."""
    from types import SimpleNamespace as NS

    fx = golden("model_forward")
    d = fx["dims"]
    graph = {"data": {"x": fx["coords"]["data"]}, "hidden": {"x": fx["coords"]["hidden"]},
             ("data", "to", "hidden"): fx["graph"]["enc"], ("hidden", "to", "hidden"): fx["graph"]["proc"], ("hidden", "to", "data"): fx["graph"]["dec"]}  # fmt: skip

    class Idx(list):  # a tensor-index list with the .prognostic / .name_to_index attributes of the reference's InputTensorIndex
        pass

    inp, outp = Idx(range(d["n_in"])), Idx(range(d["n_out"]))
    inp.prognostic, outp.prognostic = fx["in_prog"], fx["out_prog"]
    outp.name_to_index = {f"v{i}": i for i in range(d["n_out"])}
    data_indices = {"data": NS(model=NS(input=inp, output=outp))}
    common = {"trainable_size": 2, "sub_graph_edge_attributes": ["edge_length", "edge_dirs"], "num_chunks": 1, "num_heads": d["heads"], "mlp_hidden_ratio": 4,
              "qk_norm": False, "cpu_offload": False, "gradient_checkpointing": True, "shard_strategy": "edges", "graph_attention_backend": "triton"}  # fmt: skip
    cfg = {"model": {
        "num_channels": d["C"], "model": {"_target_": "anemoi.models.models.AnemoiModelEncProcDec", "hidden_nodes_name": "hidden", "latent_skip": True},
        "processor": {"_target_": "anemoi.models.layers.processor.GraphTransformerProcessor", "num_layers": 2, **common},
        "encoder": {"_target_": "anemoi.models.layers.mapper.GraphTransformerForwardMapper", **common},
        "decoder": {"_target_": "anemoi.models.layers.mapper.GraphTransformerBackwardMapper", "initialise_data_extractor_zero": False, **common},
        "residual": {"_target_": "anemoi.models.layers.residual.SkipConnection", "step": -1},
        "trainable_parameters": {"data": 0, "hidden": 3, "data2hidden": 2, "hidden2data": 2, "hidden2hidden": 2},
        "attributes": {"edges": ["edge_length", "edge_dirs"], "nodes": []},
        "bounding": [{"_target_": "anemoi.models.layers.bounding.ReluBounding", "variables": ["v1", "v3"]}],
    }}  # fmt: skip
    m = AnemoiModelEncProcDec.from_reference_config(model_config=cfg, data_indices=data_indices, n_step_input=d["t_in"], n_step_output=d["t_out"],
                                                    graph_data=graph).eval()  # fmt: skip
    ref_sd = fx["cases"]["graphtransformer"]["sd"]
    assert sorted(m.state_dict().keys()) == sorted(ref_sd.keys())
    m.load_state_dict(ref_sd, strict=True)
    skip, bound = m._output_tables("data", torch.device("cpu"))
    assert skip.tolist() == [0, 1, 2, 4, 5] and bound.tolist() == [0, 1, 0, 1, 0] and m.kind == "graphtransformer"
    bad = {"model": {**cfg["model"], "residual": {"_target_": "anemoi.models.layers.residual.TruncatedConnection"}}}
    with pytest.raises(NotImplementedError, match="SkipConnection"):
        AnemoiModelEncProcDec.from_reference_config(model_config=bad, data_indices=data_indices, n_step_input=2, n_step_output=1, graph_data=graph)


def build_two_dataset_model(fx):
    graph = {n: {"x": c} for n, c in fx["coords"].items()}
    graph.update({(src, "to", dst): sub for (src, dst), sub in fx["graph"].items()})
    names = list(fx["sizes"])
    common = dict(num_heads=fx["heads"], mlp_hidden_ratio=4, num_chunks=1)
    return AnemoiModelEncProcDec("graphtransformer", graph_data=graph, dataset_names=names, edge_attributes=["edge_length", "edge_dirs"],
                                 num_channels=fx["C"], n_step_input=fx["t_in"], n_step_output=fx["t_out"], num_input_channels=fx["n_in"],
                                 num_output_channels=fx["n_out"], internal_input_idx=fx["prog_in"], internal_output_idx=fx["prog_out"],
                                 encoder=common, processor=dict(num_layers=2, **common), decoder=common,
                                 trainable_parameters={"hidden": 2, "data2hidden": 1, "hidden2data": 1, "hidden2hidden": 1}).eval()  # fmt: skip


def test_two_dataset_model_state_dict(golden):
    """Two datasets (encoder_processor_decoder.py:203-330).  The forward itself is checked against the reference golden on CPU with the stand-in
    arithmetic (tests/test_sharded_forward_gloo.py::test_host_logic_against_reference_goldens); the GPU-parity case for it is still to be
    added (no GPU time was left in round 1 to run it, and an unrun GPU test does not belong in the suite)."""
    fx = golden("model_forward_two_datasets")
    m = build_two_dataset_model(fx)
    assert list(m.state_dict().keys()) == list(fx["sd"].keys())  # encoder.era.*, encoder.obs.*, decoder_graph_provider.obs.* ..., reference order
    m.load_state_dict(fx["sd"], strict=True)
