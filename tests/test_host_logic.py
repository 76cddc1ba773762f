"""CPU tests of the host-side mirror: integer path vs the reference goldens (bit-exact), state_dict compatibility with
the reference modules, constructor validation, layer_kernels plugin resolution, and the no-CPU-fallback contract."""
import pytest
import torch

from anemoi_core_b200.distributed import khop_edges as K
from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_range
from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
from anemoi_core_b200.distributed.shapes import GraphShardInfo
from anemoi_core_b200.layers import GNNBackwardMapper
from anemoi_core_b200.layers import GNNForwardMapper
from anemoi_core_b200.layers import GNNProcessor
from anemoi_core_b200.layers import GraphTransformerBackwardMapper
from anemoi_core_b200.layers import GraphTransformerForwardMapper
from anemoi_core_b200.layers import GraphTransformerProcessor
from anemoi_core_b200.layers.utils import compute_mlp_hidden_dim
from anemoi_core_b200.layers.utils import load_layer_kernels


def test_balanced_partition():
    # reference semantics: first `rem` parts get one extra (balanced_partition.py:16-41)
    assert get_balanced_partition_sizes(10, 3) == [4, 3, 3]
    assert get_balanced_partition_sizes(3, 5) == [1, 1, 1, 0, 0]
    assert get_balanced_partition_sizes(0, 2) == [0, 0]
    assert sum(get_balanced_partition_sizes(40962, 8)) == 40962
    assert get_balanced_partition_range(10, 3, 1) == (4, 7)
    with pytest.raises(ValueError):
        get_balanced_partition_sizes(4, 0)


def test_integer_path_matches_reference_goldens(golden):
    for c in golden("integer_path")["cases"]:
        ei, nn = c["edge_index"], c["num_nodes"]
        s, perm = K.sort_edge_index_by_dst(ei)
        assert torch.equal(s, c["sorted"]) and torch.equal(perm, c["perm"])
        assert K.is_edge_index_dst_sorted(s)
        for parts, p in c["partitions"].items():
            gp = K.build_graph_partition(s, parts, nn)
            assert list(gp.dst_splits) == p["dst_splits"] and list(gp.edge_splits) == p["edge_splits"]
            for cid, m in enumerate(p["chunks"]):
                (d0, d1), (e0, e1), connected, local = gp.materialise(cid, s)
                assert torch.equal(torch.arange(d0, d1), m["dst_ids"]) and torch.equal(torch.arange(e0, e1), m["edge_ids"])
                assert torch.equal(connected, m["src_ids"]) and torch.equal(local, m["edge_index"])


def test_ensure_sorted_permutes_attributes_with_edges():
    ei = torch.tensor([[0, 1, 2, 3], [3, 1, 2, 1]])
    ea = torch.arange(4.0).view(4, 1)
    ea2, ei2 = K.ensure_edges_are_dst_sorted(ea, ei, edges_are_dst_sorted=False)
    assert ei2.tolist() == [[1, 3, 2, 0], [1, 1, 2, 3]] and ea2.view(-1).tolist() == [1.0, 3.0, 2.0, 0.0]
    ea3, ei3 = K.ensure_edges_are_dst_sorted(ea, ei, edges_are_dst_sorted=True)
    assert ea3 is ea and ei3 is ei


CASES = [
    ("gnn_processor_small", GNNProcessor, dict(num_chunks=1, mlp_extra_layers=0)),
    ("gt_processor_small", GraphTransformerProcessor, dict(num_chunks=1, mlp_hidden_ratio=4)),
    ("gt_processor_qknorm", GraphTransformerProcessor, dict(num_chunks=1, mlp_hidden_ratio=4)),
    ("gnn_forward_mapper", GNNForwardMapper, dict(num_chunks=1, mlp_extra_layers=0)),
    ("gnn_backward_mapper", GNNBackwardMapper, dict(num_chunks=1, mlp_extra_layers=0)),
    ("gt_forward_mapper_chunks4", GraphTransformerForwardMapper, dict(mlp_hidden_ratio=4)),
    ("gt_backward_mapper", GraphTransformerBackwardMapper, dict(mlp_hidden_ratio=4)),
]


@pytest.mark.parametrize("name,cls,extra", CASES)
def test_reference_state_dicts_load_strictly(golden, name, cls, extra):
    """Parameter names/shapes are part of the drop-in surface (inference checkpoints pickle modules)."""
    g = golden(name)
    m = cls(**g["cfg"], **extra)
    res = m.load_state_dict(g["sd"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # same keys in the same ORDER: parameters() enumerates like the reference's, so a reference optimizer state (indexed by position) fits
    assert list(m.state_dict().keys()) == list(g["sd"].keys())


def test_constructor_validation_like_the_reference():
    with pytest.raises(AssertionError, match="divisible"):
        GNNProcessor(num_channels=8, num_layers=3, num_chunks=2, mlp_extra_layers=0, edge_dim=3)
    with pytest.raises(ValueError, match="divisible by num_heads"):
        GraphTransformerProcessor(num_layers=1, num_channels=30, num_chunks=1, num_heads=4, mlp_hidden_ratio=4, edge_dim=3)
    with pytest.raises(AssertionError, match="out_channels_dst"):
        GraphTransformerForwardMapper(in_channels_src=4, in_channels_dst=4, hidden_dim=8, out_channels_dst=3, num_heads=2, mlp_hidden_ratio=2,
                                      edge_dim=3)  # fmt: skip
    # tolerated extra YAML keys (gnn.yaml:22-32)
    GNNProcessor(num_channels=8, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=3, trainable_size=8, sub_graph_edge_attributes=["a"],
                 gradient_checkpointing=False)  # fmt: skip
    assert compute_mlp_hidden_dim(10, 0.25) == 3
    with pytest.raises(ValueError):
        compute_mlp_hidden_dim(10, 0)


def test_layer_kernels_plugin_hook():
    k = load_layer_kernels({"LayerNorm": {"_target_": "anemoi.models.layers.normalization.AutocastLayerNorm"},
                            "Linear": {"_target_": "torch.nn.Linear", "_partial_": True}})  # fmt: skip
    from anemoi_core_b200.layers.normalization import AutocastLayerNorm

    assert isinstance(k.LayerNorm(normalized_shape=8), AutocastLayerNorm)
    assert isinstance(k.Linear(4, 4), torch.nn.Linear)
    assert isinstance(k.Activation(), torch.nn.GELU)
    with pytest.raises(ImportError):
        load_layer_kernels({"Linear": {"_target_": "no.such.module.Linear"}})
    m = GNNProcessor(num_channels=8, num_layers=1, num_chunks=1, mlp_extra_layers=0, edge_dim=3,
                     layer_kernels={"LayerNorm": {"_target_": "anemoi.models.layers.normalization.AutocastLayerNorm"}})  # fmt: skip
    assert isinstance(m.proc[0].node_mlp.layer_norm, AutocastLayerNorm)


def test_no_cpu_fallback_in_eval_and_training_mode():
    m = GNNProcessor(num_channels=8, num_layers=1, num_chunks=1, mlp_extra_layers=0, edge_dim=3).eval()
    ei = torch.tensor([[0, 1], [0, 1]])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 8), 1, GraphShardInfo(nodes=[2]), torch.randn(2, 3), ei)
    m.train()  # training mode takes the differentiable path (layers/_train.py): the same kernels, so the same loud failure on CPU tensors
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 8), 1, GraphShardInfo(nodes=[2]), torch.randn(2, 3), ei)
    assert BipartiteGraphShardInfo().edges_are_sharded() is False and GraphShardInfo(nodes=[1]).nodes_are_sharded()


def test_gelu_exp2_polynomial():
    """The tensor-core epilogue's GELU (csrc/common.cuh gelu_erf_fast: max(x,0) - |x| * exp2(P5(|x|))) restated in fp32 numpy with the
    coefficients parsed from the source: within 1e-6 absolute of the exact erf-GELU (torch.nn.GELU(), reference layers/utils.py:107-110)."""
    import re
    from pathlib import Path

    import numpy as np

    src = (Path(__file__).resolve().parents[1] / "anemoi_core_b200" / "csrc" / "common.cuh").read_text()
    m = re.search(r"#define ANEMOI_GELU_P5 (.*)", src)
    c = [np.float32(t.strip().rstrip("f")) for t in m.group(1).split(",")]
    assert len(c) == 6
    x = np.concatenate([np.linspace(-30, 30, 600001), [0.0, -0.0, 1e-30, -1e-30, 1e4, -1e4]]).astype(np.float32)
    t = np.minimum(np.abs(x), np.float32(10.0))
    p = np.full_like(t, c[5])
    for k in range(4, -1, -1):
        p = (p * t + c[k]).astype(np.float32)
    y = np.maximum(x, 0) - t.astype(np.float64) * np.exp2(p.astype(np.float64))
    ref = torch.nn.functional.gelu(torch.from_numpy(x).double()).numpy()
    assert np.abs(y - ref).max() <= 1e-6


def test_gelu_backward_two_exp2_form():
    """The bf16 GELU backward (csrc/backward.cu gelu_vec_kernel mode 2: Phi(-|x|) = exp2(P5(|x|)), phi(x) = exp2(-x^2 / (2 ln 2)) / sqrt(2 pi))
    restated in fp32 numpy with the coefficients parsed from the source: gelu'(x) within 1.5e-5 absolute of the derivative of the exact erf-GELU
    (the bound the header and DESIGN.md quote; a bf16 cotangent's ulp is 4e-3)."""
    import re
    from pathlib import Path

    import numpy as np

    src = (Path(__file__).resolve().parents[1] / "anemoi_core_b200" / "csrc" / "common.cuh").read_text()
    c = [np.float32(t.strip().rstrip("f")) for t in re.search(r"#define ANEMOI_GELU_P5 (.*)", src).group(1).split(",")]
    bwd = (Path(__file__).resolve().parents[1] / "anemoi_core_b200" / "csrc" / "backward.cu").read_text()
    k = np.float32(re.search(r'"f"\((-0\.7213475\d*)f \* v\[j\] \* v\[j\]\)', bwd).group(1))
    assert abs(float(k) + 1.0 / (2.0 * np.log(2.0))) < 1e-7
    x = np.concatenate([np.linspace(-30, 30, 600001), [0.0, -0.0, 1e-30, -1e-30, 1e4, -1e4]]).astype(np.float32)
    t = np.minimum(np.abs(x), np.float32(10.0))
    p = np.full_like(t, c[5])
    for i in range(4, -1, -1):
        p = (p * t + c[i]).astype(np.float32)
    ex = np.exp2(p).astype(np.float32)
    pd = np.exp2((k * x * x).astype(np.float32)).astype(np.float32)
    cdf = np.where(x >= 0, np.float32(1.0) - ex, ex)
    d = (x * (np.float32(0.39894228040143267794) * pd) + cdf).astype(np.float32)
    xd = torch.from_numpy(x).double().requires_grad_()
    torch.nn.functional.gelu(xd).sum().backward()
    assert np.abs(d.astype(np.float64) - xd.grad.numpy()).max() <= 1.5e-5


def test_row_stats_tags_follow_the_tensor_identity():
    """Statistics handed from a producing GEMM to the consuming LayerNorm-GEMM travel as a tag on the tensor object and are dropped as soon
    as the tensor is not the one that was produced (slice, in-place update through torch, another object on the same storage)."""
    from anemoi_core_b200.layers import _functional as Fn

    x = torch.randn(8, 64)
    stats = torch.zeros(8, 1, 2)
    assert Fn.tagged_row_stats(x) is None
    assert Fn.tag_row_stats(x, stats) is x and Fn.tagged_row_stats(x) is stats
    assert Fn.tagged_row_stats(x[:4]) is None and Fn.tagged_row_stats(x.view(8, 64)) is None  # other tensor objects carry no tag
    x.add_(1.0)  # torch bumps the version counter: the statistics no longer describe the contents
    assert Fn.tagged_row_stats(x) is None
    assert Fn.wants_row_stats(512, torch.bfloat16) and not Fn.wants_row_stats(512, torch.float32)
    assert not Fn.wants_row_stats(36, torch.bfloat16) and not Fn.wants_row_stats(4096, torch.bfloat16)


def test_model_output_tables_encode_residual_and_boundings(golden):
    """skip_src[v] = input variable whose last step is added to output variable v (prognostic pairs), bound[v] = 0 / 1 (relu) / 2 (leaky)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("_glue", os.path.join(os.path.dirname(__file__), "test_model_glue.py"))
    glue = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(glue)
    build_model = glue.build_model

    fx = golden("model_forward")
    m = build_model(fx, "gnn")
    m._bound_spec["data"] = [("relu", [1, 3]), ("leaky_relu", [4])]
    skip, bound = m._output_tables("data", torch.device("cpu"))
    assert skip.dtype == torch.int32 and skip.tolist() == [0, 1, 2, 4, 5]  # out_prog [0..4] <- in_prog [0, 1, 2, 4, 5]
    assert bound.tolist() == [0, 1, 0, 1, 2]
    assert m._output_tables("data", torch.device("cpu"))[0] is skip  # cached


def test_reference_default_yaml_kwargs_construct():
    """The reference's default model configs (training/config/model/graphtransformer.yaml:11-75, gnn.yaml:11-57) instantiate the mirror
    classes unchanged: every key of the YAML blocks is passed as a keyword (interpolations resolved), `_target_`-style layer_kernels included."""
    gt_kernels = {"LayerNorm": {"_target_": "torch.nn.LayerNorm"}, "Linear": {"_target_": "torch.nn.Linear"}, "Activation": {"_target_": "torch.nn.GELU"},
                  "QueryNorm": {"_target_": "anemoi_core_b200.layers.normalization.AutocastLayerNorm", "bias": False},
                  "KeyNorm": {"_target_": "anemoi_core_b200.layers.normalization.AutocastLayerNorm", "bias": False}}  # fmt: skip
    common = dict(trainable_size=8, sub_graph_edge_attributes=["edge_length", "edge_dirs"], mlp_hidden_ratio=4, mlp_implementation="mlp", num_heads=16,
                  qk_norm=False, cpu_offload=False, gradient_checkpointing=True, layer_kernels=gt_kernels, shard_strategy="edges",
                  graph_attention_backend="triton", edge_pre_mlp=False)  # fmt: skip
    C = 64  # (the YAML says 1024; the constructor logic does not depend on it)
    GraphTransformerProcessor(num_layers=16, num_chunks=4, num_channels=C, edge_dim=11, **common)
    GraphTransformerForwardMapper(num_chunks=4, in_channels_src=30, in_channels_dst=12, hidden_dim=C, edge_dim=11, **common)
    GraphTransformerBackwardMapper(num_chunks=4, initialise_data_extractor_zero=False, in_channels_src=C, in_channels_dst=30, hidden_dim=C,
                                   out_channels_dst=9, edge_dim=11, **common)  # fmt: skip
    gnn_kernels = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.AutocastLayerNorm"}, "Linear": {"_target_": "torch.nn.Linear"},
                   "Activation": {"_target_": "torch.nn.GELU"}}  # fmt: skip
    gcommon = dict(trainable_size=8, sub_graph_edge_attributes=["edge_length", "edge_dirs"], mlp_extra_layers=0, mlp_hidden_ratio=1.0,
                   mlp_implementation="mlp", cpu_offload=False, gradient_checkpointing=True, layer_kernels=gnn_kernels)  # fmt: skip
    GNNProcessor(num_layers=16, num_chunks=2, num_channels=C, edge_dim=11, **gcommon)
    GNNForwardMapper(num_chunks=1, in_channels_src=30, in_channels_dst=12, hidden_dim=C, edge_dim=11, **gcommon)
    GNNBackwardMapper(num_chunks=1, in_channels_src=C, in_channels_dst=C, hidden_dim=C, out_channels_dst=9, edge_dim=11, **gcommon)
    # the gated variants the YAML comments name
    GraphTransformerProcessor(num_layers=2, num_chunks=1, num_channels=C, edge_dim=11, **{**common, "mlp_implementation": "swiglu", "mlp_hidden_ratio": 2.67})


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints one JSON line with the contract's keys, no GPU needed."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]  # fmt: skip
    line = json.loads(out)
    assert line["impl"] == "reference" and line["unit"] == "ms/step" and line["higher_is_better"] is False and line["value"] > 0
    # "reference" = the unmodified reference modules (present here under /root/reference or baseline/_ref), "port" = the oracle restatement
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["steps"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 of a torchrun launch exit without work and without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    quiet = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1"],
                           capture_output=True, text=True, timeout=600, env=env)  # fmt: skip
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_integer_path_fuzz_against_the_unmodified_reference():
    """A few hundred random graphs (incl. parts without edges, one hub destination, duplicate edges, parts == destinations) through the integer /
    index path — sort, partition splits, part materialisation, src compaction, balanced sizes — bit-exact against the UNMODIFIED reference
    functions (oracle/fuzz_integer_path.py, in a subprocess: the reference's ``anemoi`` package must not share a process with the overlay tests)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "fuzz_integer_path.py"), "150"], capture_output=True, text=True, timeout=600,
                         check=True).stdout.strip().splitlines()[-1]  # fmt: skip
    res = json.loads(out)
    if "unavailable" in res:
        pytest.skip(res["unavailable"])
    assert res["failures"] == [] and res["checks"] > 5000, res


def test_module_fuzz_against_the_unmodified_reference():
    """Random constructor configurations of both processors and all four mappers (channels, heads, layers, chunks, hidden ratio, qk_norm, gated
    MLPs, edge_pre_mlp, attn_channels, unsorted edges, bipartite sizes): the inference path and the training path (output + gradients of every
    parameter, the node inputs, the edge attributes) of the drop-in modules over the CPU stand-ins against the UNMODIFIED reference modules with the
    same ``state_dict`` (oracle/fuzz_modules.py, in a subprocess), 1e-4 of each tensor's scale."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "fuzz_modules.py"), "60"], capture_output=True, text=True, timeout=900,
                         check=True).stdout.strip().splitlines()[-1]  # fmt: skip
    res = json.loads(out)
    if "unavailable" in res:
        pytest.skip(res["unavailable"])
    assert res["failures"] == [] and res["cases"] == 60 and set(res["worst"]) == {"inference", "training forward", "input gradients", "parameter gradients"}, res


def test_modules_deepcopy_and_pickle(golden):
    """``copy.deepcopy`` (EMA / SWA wrappers, Lightning callbacks) and pickling of whole modules (the reference's inference checkpoints pickle the
    model object) work on every drop-in class and on the model, and the copies carry the same ``state_dict``."""
    import copy
    import io

    from anemoi_core_b200.layers import GNNBackwardMapper
    from anemoi_core_b200.layers import GNNForwardMapper
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerBackwardMapper
    from anemoi_core_b200.layers import GraphTransformerForwardMapper
    from anemoi_core_b200.layers import GraphTransformerProcessor
    from test_model_glue import build_model

    mods = [GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=5, qk_norm=True),
            GNNProcessor(num_channels=16, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=5, mlp_implementation="swiglu"),
            GraphTransformerForwardMapper(in_channels_src=7, in_channels_dst=5, hidden_dim=32, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=3),
            GraphTransformerBackwardMapper(in_channels_src=32, in_channels_dst=7, hidden_dim=32, out_channels_dst=4, num_chunks=1, num_heads=4,
                                           mlp_hidden_ratio=2, edge_dim=3),
            GNNForwardMapper(in_channels_src=7, in_channels_dst=5, hidden_dim=16, num_chunks=1, mlp_extra_layers=0, edge_dim=3),
            GNNBackwardMapper(in_channels_src=16, in_channels_dst=16, hidden_dim=16, out_channels_dst=4, num_chunks=1, mlp_extra_layers=0, edge_dim=3),
            build_model(golden("model_forward"), "graphtransformer"), build_model(golden("model_forward"), "gnn")]  # fmt: skip
    for m in mods:
        for clone in (copy.deepcopy(m), torch.load(io.BytesIO(_dump(m)), weights_only=False)):
            assert type(clone) is type(m) and clone is not m
            a, b = m.state_dict(), clone.state_dict()
            assert list(a) == list(b) and all(torch.equal(a[k], b[k]) and a[k].data_ptr() != b[k].data_ptr() for k in a if a[k].numel())


def _dump(m) -> bytes:
    import io

    buf = io.BytesIO()
    torch.save(m, buf)
    return buf.getvalue()


def test_compute_dtype_policy(monkeypatch):
    """fp32 / bf16 by input dtype or bf16 autocast; fp16 autocast (the reference's default "16-mixed") maps to bf16 with one warning, to fp32 or to
    an error by ANEMOI_B200_FP16_AUTOCAST; fp16 TENSORS are refused with a message that says what to do."""
    import warnings

    from anemoi_core_b200.layers import _functional as Fn

    a, b = torch.zeros(2, 2), torch.zeros(2, 2, dtype=torch.bfloat16)
    assert Fn.compute_dtype(a) == torch.float32 and Fn.compute_dtype(b) == torch.bfloat16 and Fn.compute_dtype(a, b) == torch.float32
    with pytest.raises(NotImplementedError, match="bfloat16"):
        Fn.compute_dtype(a.half())
    state = {"dt": torch.bfloat16}
    monkeypatch.setattr(torch, "is_autocast_enabled", lambda *args: True)
    monkeypatch.setattr(torch, "get_autocast_dtype", lambda *args: state["dt"])
    assert Fn.compute_dtype(a) == torch.bfloat16
    state["dt"] = torch.float16
    monkeypatch.setattr(Fn, "_warned_fp16", False)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert Fn.compute_dtype(a) == torch.bfloat16 and Fn.compute_dtype(a) == torch.bfloat16
    assert len([w for w in rec if "fp16 autocast" in str(w.message)]) == 1  # once
    monkeypatch.setattr(Fn, "FP16_AUTOCAST", "fp32")
    assert Fn.compute_dtype(a) == torch.float32
    monkeypatch.setattr(Fn, "FP16_AUTOCAST", "error")
    with pytest.raises(NotImplementedError):
        Fn.compute_dtype(a)


def test_call_surface_matches_the_unmodified_reference():
    """Constructor parameters, their effective defaults and the positional order of ``forward`` of the 15 drop-in classes against the reference
    classes themselves (oracle/check_signatures.py, in a subprocess)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "check_signatures.py")], capture_output=True, text=True, timeout=300,
                         check=True).stdout.strip().splitlines()[-1]  # fmt: skip
    res = json.loads(out)
    if "unavailable" in res:
        pytest.skip(res["unavailable"])
    assert res["problems"] == [] and res["classes"] == 15, res
