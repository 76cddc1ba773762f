"""Kernel-level parity on a B200 (-m gpu): every C-ABI entry point against the oracle / a plain fp32 PyTorch statement
of the same op on the same seeded inputs.  Integer outputs bit-exact; fp32 kernels to ~1e-5; bf16 kernels are compared
with an fp32 reference evaluated on the SAME bf16-rounded operands (so the only differences are fp32 accumulation
order and the final bf16 rounding: tolerance 2^-7 relative to the row scale)."""
import math

import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from anemoi_core_b200 import ops as _ops

    return _ops


def _rand_graph(n_src, n_dst, n_edges, seed, zero_tail=0):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n_src, (n_edges,), generator=g), torch.randint(0, max(n_dst - zero_tail, 1), (n_edges,), generator=g)])
    return ei[:, torch.sort(ei[1], stable=True)[1]]


# ---------------------------------------------------------------------------------------------------------------
def test_csr_build_bit_exact(ops, golden):
    for c in golden("integer_path")["cases"]:
        n_src, n_dst = c["num_nodes"]
        csr = ops.build_csr(c["sorted"].cuda(), n_src, n_dst)
        assert csr.colptr.dtype == torch.int64
        assert torch.equal(csr.colptr.cpu(), c["colptr"])
        assert torch.equal(csr.colptr32.cpu().long(), c["colptr"])
        assert torch.equal(csr.src32.cpu().long(), c["sorted"][0])
        assert torch.equal(csr.dst32.cpu().long(), c["sorted"][1])
    # larger, with empty rows at both ends
    ei = _rand_graph(5000, 7000, 60000, 3, zero_tail=50)
    ei[1] += 0
    csr = ops.build_csr(ei.cuda(), 5000, 7000)
    assert torch.equal(csr.colptr.cpu(), R.index2ptr(ei[1], 7000))


def test_csr_build_rejects_unsorted_and_out_of_range(ops):
    ei = torch.tensor([[0, 1, 2], [2, 1, 0]])
    with pytest.raises(ValueError, match="not sorted"):
        ops.build_csr(ei.cuda(), 3, 3)
    with pytest.raises(ValueError, match="outside"):
        ops.build_csr(torch.tensor([[0, 5], [0, 1]]).cuda(), 3, 3)


@pytest.mark.parametrize("C,groups", [(32, 1), (64, 1), (512, 1), (1024, 1), (100, 1), (2048, 1), (16, 4), (32, 16), (6, 6)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layer_norm(ops, C, groups, dt):
    g = torch.Generator().manual_seed(C + groups)
    M = 777
    x = (torch.randn(M, groups * C, generator=g) * 2 + 0.5).to(dt)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    r = torch.randn(M, groups * C, generator=g).to(dt)
    ref = torch.nn.functional.layer_norm(x.float().view(M, groups, C), (C,), w, b, 1e-5).view(M, -1) + r.float()
    y = ops.layer_norm(x.cuda(), w.cuda(), b.cuda(), 1e-5, residual=r.cuda(), groups=groups)
    assert y.dtype == dt
    tol = 2e-5 if dt == torch.float32 else 2**-7
    torch.testing.assert_close(y.float().cpu(), ref, atol=tol * 4, rtol=tol)
    # no affine, no residual, fp32 output from any input, strided input (column slice)
    wide = torch.zeros(M, groups * C + 8, dtype=dt)
    wide[:, : groups * C] = x
    y2 = ops.layer_norm(wide.cuda()[:, : groups * C], None, None, 1e-5, out_dtype=torch.float32, groups=groups)
    ref2 = torch.nn.functional.layer_norm(x.float().view(M, groups, C), (C,), None, None, 1e-5).view(M, -1)
    torch.testing.assert_close(y2.cpu(), ref2, atol=1e-4, rtol=1e-4)


def _linear_ref(a, w, bias, gelu, residual, g1, i1, g2, i2):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if g1 is not None:
        y = y + g1[i1.long()][:, : y.shape[1]]
    if g2 is not None:
        y = y + g2[i2.long()][:, : y.shape[1]]
    if gelu:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


LINEAR_SHAPES = [
    # M, N, K           fp32 FFMA path and bf16 tcgen05 path (K >= 64, K % 8 == 0) / bf16 FFMA path (small or odd K)
    (300, 512, 512),
    (1000, 2048, 512),
    (129, 512, 2048),
    (257, 88, 512),
    (4099, 64, 64),
    (515, 200, 216),
    (100, 64, 11),
    (77, 33, 70),
    (1, 8, 64),
]


@pytest.mark.parametrize("M,N,K", LINEAR_SHAPES)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_res", "gather"])
def test_linear(ops, M, N, K, dt, epi):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(dt)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dt)
    bias = torch.randn(N, generator=g) if epi != "plain" else None
    residual = torch.randn(M, N, generator=g).to(dt) if epi == "bias_res" else None
    g1 = i1 = g2 = i2 = None
    if epi == "gather":
        ldg = (N + 3) // 4 * 4
        g1, g2 = torch.randn(50, ldg, generator=g), torch.randn(60, ldg, generator=g)
        i1 = torch.randint(0, 50, (M,), generator=g, dtype=torch.int32)
        i2 = torch.randint(0, 60, (M,), generator=g, dtype=torch.int32)
    ref = _linear_ref(a, w, bias, epi in ("bias_gelu", "gather"), residual, g1, i1, g2, i2)
    cu = lambda t: None if t is None else t.cuda()
    y = ops.linear(cu(a), cu(w), cu(bias), gelu=epi in ("bias_gelu", "gather"), residual=cu(residual),
                   gather1=None if g1 is None else (cu(g1)[:, :N] if ldg == N else cu(g1), cu(i1)),
                   gather2=None if g2 is None else (cu(g2)[:, :N] if ldg == N else cu(g2), cu(i2)))  # fmt: skip
    assert y.dtype == dt and tuple(y.shape) == (M, N)
    if dt == torch.float32:
        torch.testing.assert_close(y.cpu(), ref, atol=2e-5, rtol=2e-5)
    else:
        scale = ref.abs().max().item()
        err = (y.float().cpu() - ref).abs().max().item()
        assert err <= 2**-7 * scale + 1e-3, f"max err {err} at scale {scale}"
    # fp32 output from bf16 operands (node projections for the gather-add epilogue)
    if dt == torch.bfloat16 and epi == "plain":
        y32 = ops.linear(cu(a), cu(w), out_dtype=torch.float32)
        torch.testing.assert_close(y32.cpu(), ref, atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("M,N,K,epi", [(40962, 2240, 512, "bias"), (40962, 512, 704, "bias_res"), (20001, 2048, 512, "bias_gelu"), (37900, 512, 2048, "bias_res"),
                                       (40962, 1000, 520, "bias_gelu")])
def test_linear_large_two_cta(ops, M, N, K, epi):
    """Shapes large enough for the cta_group::2 kernel (256x256 tiles per CTA pair), ragged M / N / K tails included.
    (The 2-CTA kernel is the DEFAULT for problems that fill the 74 CTA pairs; ANEMOI_B200_GEMM_CG=1 forces the single-CTA kernel.)
    Checker: fp32 torch matmul on the same bf16-rounded operands (on the GPU: the CPU would take minutes)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16) if epi == "bias_res" else None
    y = ops.linear(a, w, bias, gelu=epi == "bias_gelu", residual=res)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = a.float() @ w.float().t() + bias
    if epi == "bias_gelu":
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res.float()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2**-7 * ref.abs().max().item() + 1e-3, f"max err {err}"
    # a second call with fp32 output (node projections of the GNN path)
    y32 = ops.linear(a, w, bias, gelu=epi == "bias_gelu", residual=None, out_dtype=torch.float32)
    ref32 = a.float() @ w.float().t() + bias
    if epi == "bias_gelu":
        ref32 = torch.nn.functional.gelu(ref32)
    assert (y32 - ref32).abs().max().item() <= 2e-3 * max(1.0, ref32.abs().max().item())


@pytest.mark.parametrize("M,N,K,gelu", [(5000, 2240, 512, False), (40962, 2048, 512, True), (777, 88, 512, False), (3000, 512, 64, True)])
def test_linear_with_folded_layer_norm(ops, M, N, K, gelu):
    """LN(x) W^T + b through the folded form (row_stats + gamma-scaled weight + column sums) against LayerNorm -> Linear in fp32."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = (torch.randn(M, K, generator=g, device="cuda") * 1.7 + 0.6).to(torch.bfloat16)
    w32 = torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)
    b32 = torch.randn(N, generator=g, device="cuda")
    gamma, beta = 1 + 0.2 * torch.randn(K, generator=g, device="cuda"), 0.3 * torch.randn(K, generator=g, device="cuda")
    wf = (w32 * gamma).to(torch.bfloat16)
    bias = b32 + w32 @ beta
    stats = ops.row_stats(x, 1e-5)
    xf = x.float()
    torch.testing.assert_close(stats[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(stats[:, 1], (xf.var(1, unbiased=False) + 1e-5).rsqrt(), atol=1e-4, rtol=1e-4)
    y = ops.linear(x, wf, bias, gelu=gelu, ln_stats=stats, ln_colsum=wf.float().sum(1).contiguous())
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ w32.t() + b32
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    err = (y.float() - ref).abs().max().item()
    assert err <= 2**-6 * ref.abs().max().item() + 1e-3, f"max err {err}"


def _block_stats(y: torch.Tensor) -> torch.Tensor:
    """[M, ceil(N/64), 2] (block mean, block M2 = sum of squared deviations from it) of the stored values, fp64 reference."""
    M, N = y.shape
    P = (N + 63) // 64
    out = torch.zeros(M, P, 2, dtype=torch.float64, device=y.device)
    for b in range(P):
        blk = y[:, b * 64 : min(N, (b + 1) * 64)].double()
        mean = blk.mean(1)
        out[:, b, 0], out[:, b, 1] = mean, ((blk - mean[:, None]) ** 2).sum(1)
    return out


@pytest.mark.parametrize("M,N,K,epi", [(40962, 512, 704, "res"), (40962, 512, 2048, "res"), (1000, 512, 512, "plain"), (333, 200, 128, "res"),
                                       (5000, 1024, 256, "res"), (2000, 512, 512, "gelu"), (700, 512, 96, "f32")])  # fmt: skip
def test_linear_row_stats_epilogue(ops, M, N, K, epi):
    """stats_out of the producing GEMM == block sums of its STORED output (fused in the tcgen05 epilogue for plain / residual bf16; the
    gelu and fp32 cases take the separate pass), and a consumer GEMM fed those partials == the same GEMM fed ops.row_stats."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dt = torch.float32 if epi == "f32" else torch.bfloat16
    a = torch.randn(M, K, generator=g, device="cuda").to(dt)
    w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).to(dt)
    bias = torch.randn(N, generator=g, device="cuda")
    res = (torch.randn(M, N, generator=g, device="cuda") * 2 + 0.5).to(dt) if epi == "res" else None
    stats = ops.partial_stats_buffer(M, N, a.device)
    stats.fill_(float("nan"))
    y = ops.linear(a, w, bias, gelu=epi == "gelu", residual=res, stats_out=stats)
    ref = _block_stats(y)
    assert torch.isfinite(stats).all()
    std = (ref[..., 1] / 64).sqrt().clamp_min(1e-3)
    assert ((stats[..., 0].double() - ref[..., 0]).abs() <= 1e-5 * (std + ref[..., 0].abs())).all()
    assert ((stats[..., 1].double() - ref[..., 1]).abs() <= 1e-4 * ref[..., 1].clamp_min(1.0)).all()
    if dt != torch.bfloat16 or N % 8 or N < 64:
        return
    # consumer: LayerNorm over y's N columns folded into the next GEMM, statistics from the partials vs from the row_stats pass
    N2 = 256
    w2 = (torch.randn(N2, N, generator=g, device="cuda") / math.sqrt(N)).to(dt)
    b2 = torch.randn(N2, generator=g, device="cuda")
    colsum = w2.float().sum(1).contiguous()
    y_partial = ops.linear(y, w2, b2, ln_stats=stats, ln_dim=N, ln_eps=1e-5, ln_colsum=colsum)
    y_rowstats = ops.linear(y, w2, b2, ln_stats=ops.row_stats(y, 1e-5), ln_colsum=colsum)
    ref2 = torch.nn.functional.layer_norm(y.float(), (N,), None, None, 1e-5) @ w2.float().t() + b2
    tol = 2**-6 * ref2.abs().max().item() + 1e-3
    assert (y_partial.float() - ref2).abs().max().item() <= tol
    assert (y_partial.float() - y_rowstats.float()).abs().max().item() <= 2**-7 * ref2.abs().max().item()


def test_linear_row_stats_constant_rows(ops):
    """A constant output row has zero variance: the E[x^2] - mean^2 form must clamp at 0 and the folded LayerNorm of it must give the bias."""
    M, N, K = 512, 512, 128
    a = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    res = torch.full((M, N), 1.0078125, dtype=torch.bfloat16, device="cuda")  # exactly representable; rows 0..255
    res[256:] = torch.randn(256, N, device="cuda").to(torch.bfloat16)
    stats = ops.partial_stats_buffer(M, N, a.device)
    y = ops.linear(a, w, None, residual=res, stats_out=stats)
    assert torch.equal(y, res)
    w2 = torch.randn(64, N, device="cuda").to(torch.bfloat16)
    b2 = torch.randn(64, device="cuda")
    out = ops.linear(y, w2, b2, ln_stats=stats, ln_dim=N, ln_eps=1e-5, ln_colsum=w2.float().sum(1).contiguous())
    assert torch.isfinite(out).all()
    # LN(constant) = 0 -> out = bias, up to the fp32 accumulation error of sum_k x w'_k against mean * colsum times rstd = 316
    assert (out[:256].float() - b2).abs().max().item() <= 0.05


@pytest.mark.parametrize("M,C", [(40962, 512), (1001, 1024), (37, 64), (500, 200), (3, 2048)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("mean", [0.0, 300.0])
def test_row_stats(ops, M, C, dt, mean):
    """(mean, rstd) per row against the two-pass fp64 statement, also for rows whose mean dwarfs their spread (the streaming kernel
    accumulates shifted sums); bf16 rows of a multiple of 64 elements take row_stats_stream_kernel, the rest the per-warp-row kernel."""
    g = torch.Generator().manual_seed(M + C)
    x = (mean + torch.randn(M, C, generator=g) * (2.0 if mean else 1.0)).to(dt).cuda()
    st = ops.row_stats(x, 1e-5).double().cpu()
    xd = x.double().cpu()
    mu = xd.mean(1)
    rstd = 1.0 / torch.sqrt(xd.var(1, unbiased=False) + 1e-5)
    assert (st[:, 0] - mu).abs().max().item() <= 1e-5 * max(1.0, abs(mean))
    assert ((st[:, 1] - rstd).abs() / rstd).max().item() <= 2e-4


def test_linear_row_stats_large_mean_small_variance(ops):
    """rows with |mean| >> std.  The producer accumulates its block sums shifted by the block's first element and hands over
    (block mean, block M2), the consumer merges them with Chan's formula: the folded LayerNorm must agree with the two-pass row_stats
    path even where E[x^2] - mean^2 in fp32 would have lost every digit of the variance (mean 300, std 0.05: mean^2 / var = 3.6e7)."""
    torch.manual_seed(3)
    M, N, K = 2048, 512, 256
    a = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    res = (300.0 + 2.0 * torch.randn(M, N, device="cuda")).to(torch.bfloat16)  # bf16 spacing at 300 is 2: std ~ 2 on a mean of 300
    stats = ops.partial_stats_buffer(M, N, a.device)
    y = ops.linear(a, w, None, residual=res, stats_out=stats)
    assert torch.equal(y, res)
    ref = _block_stats(y)
    assert ((stats[..., 0].double() - ref[..., 0]).abs() <= 1e-4).all()
    assert ((stats[..., 1].double() - ref[..., 1]).abs() <= 1e-4 * ref[..., 1].clamp_min(1.0)).all()
    w2 = (torch.randn(64, N, device="cuda") / math.sqrt(N)).to(torch.bfloat16)
    b2 = torch.randn(64, device="cuda")
    colsum = w2.float().sum(1).contiguous()
    out_partial = ops.linear(y, w2, b2, ln_stats=stats, ln_dim=N, ln_eps=1e-5, ln_colsum=colsum)
    out_twopass = ops.linear(y, w2, b2, ln_stats=ops.row_stats(y, 1e-5), ln_colsum=colsum)
    # same GEMM, same epilogue: only (mean, rstd) differ, and they must agree to fp32 rounding
    assert (out_partial.float() - out_twopass.float()).abs().max().item() <= 2**-7 * out_twopass.float().abs().max().item()


def test_linear_strided_views(ops):
    """Column slices of wider buffers as A, residual and out (how the blocks pass q|k|v|self and x|aggregate)."""
    g = torch.Generator().manual_seed(5)
    for dt in (torch.float32, torch.bfloat16):
        big_a = torch.randn(333, 1024, generator=g).to(dt).cuda()
        w = (torch.randn(256, 512, generator=g) / 22).to(dt).cuda()
        big_o = torch.zeros(333, 768, dtype=dt, device="cuda")
        res = torch.randn(333, 512, generator=g).to(dt).cuda()
        ops.linear(big_a[:, 512:], w, residual=res[:, 256:], out=big_o[:, 256:512])
        ref = big_a[:, 512:].float() @ w.float().t() + res[:, 256:].float()
        tol = 2e-5 if dt == torch.float32 else 2**-7 * ref.abs().max().item()
        assert (big_o[:, 256:512].float() - ref).abs().max().item() <= tol + 1e-5
        assert torch.all(big_o[:, :256] == 0) and torch.all(big_o[:, 512:] == 0)


# ---------------------------------------------------------------------------------------------------------------
def test_gt_attention_golden_conv_cases(ops, golden):
    """The reference's own Triton-parity shapes (test_triton_gt.py:49-57, non-power-of-two heads/channels) + zero in-degree rows,
    at the reference's bar atol=1e-4 (test_triton_gt.py:135-136)."""
    for c in golden("gt_conv")["cases"]:
        n_dst, H, Ch = c["q"].shape
        n_src = c["k"].shape[0]
        csr = ops.build_csr(c["edge_index"].cuda(), n_src, n_dst)
        out = ops.gt_attention(c["q"].reshape(n_dst, -1).cuda(), c["k"].reshape(n_src, -1).cuda(), c["v"].reshape(n_src, -1).cuda(), csr, H,
                               e_proj=c["e"].reshape(-1, H * Ch).cuda())  # fmt: skip
        torch.testing.assert_close(out.cpu().view(n_dst, H, Ch), c["out"], atol=1e-4, rtol=0)
        assert torch.all(out[-2:] == 0)


@pytest.mark.parametrize("H,Ch", [(16, 32), (16, 64), (4, 16), (8, 8), (2, 32), (3, 20)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gt_attention_modes(ops, H, Ch, dt):
    """Materialised-e and fused-lin_edge forms against the oracle (PyG-softmax statement) on a bipartite random graph."""
    g = torch.Generator().manual_seed(H * 100 + Ch)
    n_src, n_dst, E, d_e = 300, 257, 2500, 11
    ei = _rand_graph(n_src, n_dst, E, 11, zero_tail=3)
    C = H * Ch
    q = torch.randn(n_dst, C, generator=g).to(dt)
    k = torch.randn(n_src, C, generator=g).to(dt)
    v = torch.randn(n_src, C, generator=g).to(dt)
    a = torch.randn(E, d_e, generator=g)
    w_e, b_e = torch.randn(C, d_e, generator=g) / 3, torch.randn(C, generator=g)
    add = torch.randn(n_dst, C, generator=g).to(dt)
    e_proj = (a @ w_e.t() + b_e).to(dt)
    sh = lambda t, n: t.float().view(n, H, Ch)
    ref = R.gt_attention(sh(q, n_dst), sh(k, n_src), sh(v, n_src), sh(e_proj, E), ei, n_dst).view(n_dst, C)
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    out1 = ops.gt_attention(q.cuda(), k.cuda(), v.cuda(), csr, H, e_proj=e_proj.cuda())
    a_pad = torch.zeros(E, 12)
    a_pad[:, :d_e] = a
    out2 = ops.gt_attention(q.cuda(), k.cuda(), v.cuda(), csr, H, edge_attr=a_pad.cuda(), w_edge=w_e.cuda(), b_edge=b_e.cuda(), add=add.cuda())
    tol = 1e-4 if dt == torch.float32 else 2**-6 * ref.abs().max().item()
    assert (out1.float().cpu() - ref).abs().max().item() <= tol
    # fused form: e_proj is never rounded to bf16, so compare against the fp32-e reference for bf16 too
    ref2 = R.gt_attention(sh(q, n_dst), sh(k, n_src), sh(v, n_src), (a @ w_e.t() + b_e).view(E, H, Ch), ei, n_dst).view(n_dst, C) + add.float()
    assert (out2.float().cpu() - ref2).abs().max().item() <= (2e-4 if dt == torch.float32 else 2**-6 * ref2.abs().max().item())
    assert torch.equal(out2[-3:].float().cpu(), add[-3:].float())  # zero in-degree rows: exactly 0 + add


@pytest.mark.parametrize("H,Ch,d_e", [(16, 32, 11), (16, 64, 11), (4, 16, 3), (8, 64, 16), (2, 32, 5)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gt_attention_folded_lin_edge(ops, H, Ch, d_e, dt):
    """Folded form (include/anemoi_b200.h form 3, the path the blocks use): qw = W_e,h^T q_h in, abar = sum alpha a out;
    out + W_e abar must equal attention with the materialised projection.  Strided q|k|v|self|qw buffer like the block's."""
    if not ops.attention_fold_supported(H * Ch, H, dt, d_e):
        pytest.skip("shape served by the generic kernel")
    g = torch.Generator().manual_seed(H * 1000 + Ch + d_e)
    n_src, n_dst, E = 300, 301, 2600
    ei = _rand_graph(n_src, n_dst, E, 13, zero_tail=5)
    C = H * Ch
    dp = (d_e + 3) // 4 * 4
    a = torch.randn(E, d_e, generator=g)
    w_e, b_e = torch.randn(C, d_e, generator=g) / 3, torch.randn(C, generator=g)
    buf = torch.randn(n_dst, 2 * C + H * dp, generator=g)  # q | self | qw
    k = torch.randn(n_src, C, generator=g).to(dt)
    v = torch.randn(n_src, C, generator=g).to(dt)
    q32 = buf[:, :C].to(dt).float()
    qw = torch.zeros(n_dst, H, dp)
    qw[:, :, :d_e] = torch.einsum("nhc,hca->nha", q32.view(n_dst, H, Ch), w_e.view(H, Ch, d_e))
    buf[:, 2 * C :] = qw.view(n_dst, -1)
    buf = buf.to(dt).cuda()
    a16 = torch.zeros(E, 16)
    a16[:, :d_e] = a
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    out = torch.empty(n_dst, C + H * dp, dtype=dt, device="cuda")
    ops.gt_attention(buf[:, :C], k.cuda(), v.cuda(), csr, H, edge_attr=a16.cuda(), b_edge=b_e.cuda(), qw=buf[:, 2 * C :], abar=out[:, C:], dp=dp,
                     add=buf[:, C : 2 * C], out=out[:, :C])  # fmt: skip
    o = out.float().cpu()
    abar = o[:, C:].view(n_dst, H, dp)[:, :, :d_e]
    full = o[:, :C] + torch.einsum("nha,hca->nhc", abar, w_e.view(H, Ch, d_e)).reshape(n_dst, C)
    sh = lambda t, n: t.float().view(n, H, Ch)
    ref = R.gt_attention(sh(q32, n_dst), sh(k, n_src), sh(v, n_src), (a @ w_e.t() + b_e).view(E, H, Ch), ei, n_dst).view(n_dst, C)
    ref = ref + buf[:, C : 2 * C].float().cpu()
    tol = 3e-4 if dt == torch.float32 else 2**-5 * ref.abs().max().item()
    assert (full - ref).abs().max().item() <= tol
    assert torch.equal(o[-5:, :C], buf[-5:, C : 2 * C].float().cpu())  # zero in-degree: 0 + add, no bias
    assert torch.all(o[-5:, C:] == 0)


def _local_graph(n_src, n_dst, deg, span, seed, zero_tail=0):
    """dst d takes ``deg`` sources from a window of ``span`` rows around d * n_src / n_dst: neighbouring destinations share sources."""
    g = torch.Generator().manual_seed(seed)
    d = torch.arange(n_dst - zero_tail).repeat_interleave(deg)
    centre = (d.float() * n_src / n_dst).long()
    src = (centre + torch.randint(-span, span + 1, (d.numel(),), generator=g)).clamp_(0, n_src - 1)
    return torch.stack([src, d])


@pytest.mark.parametrize("H,Ch,d_e", [(16, 32, 11), (16, 64, 11), (8, 32, 3), (4, 64, 16), (8, 64, 5)])
@pytest.mark.parametrize("graph", ["random", "local", "bipartite_local"])
def test_gt_attention_tiled(ops, H, Ch, d_e, graph):
    """Destination-tile tensor-core kernel (anemoi_b200_gt_attention_tiled_fwd + the host planner) against the oracle attention with the
    materialised projection on the same bf16-rounded operands, and against the warp-per-node kernel.  Random graphs (no locality: small
    tiles, duplicate (src, dst) pairs, rows without edges), local graphs (full 16-row tiles, heavy source re-use) and bipartite sizes."""
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(H * 1000 + Ch + d_e)
    if graph == "random":
        n_src, n_dst, E = 300, 301, 2600
        ei = _rand_graph(n_src, n_dst, E, 13, zero_tail=5)
    elif graph == "local":
        n_src = n_dst = 1000
        ei = _local_graph(n_src, n_dst, 9, 6, 3, zero_tail=5)
    else:
        n_src, n_dst = 700, 1003
        ei = _local_graph(n_src, n_dst, 5, 4, 4, zero_tail=5)
    E = ei.shape[1]
    C = H * Ch
    dp = (d_e + 3) // 4 * 4
    assert ops.attention_tiles_supported(C, H, dt, dp)
    a = torch.randn(E, d_e, generator=g)
    w_e, b_e = torch.randn(C, d_e, generator=g) / 3, torch.randn(C, generator=g)
    buf = torch.randn(n_dst, 2 * C + H * dp, generator=g)  # q | self | qw
    k = torch.randn(n_src, C, generator=g).to(dt)
    v = torch.randn(n_src, C, generator=g).to(dt)
    q32 = buf[:, :C].to(dt).float()
    qw = torch.zeros(n_dst, H, dp)
    qw[:, :, :d_e] = torch.einsum("nhc,hca->nha", q32.view(n_dst, H, Ch), w_e.view(H, Ch, d_e))
    buf[:, 2 * C :] = qw.view(n_dst, -1)
    buf = buf.to(dt).cuda()
    a16 = torch.zeros(E, 16)
    a16[:, :d_e] = a
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    plan = ops.attention_tiles(csr)
    assert plan is not None and plan.n_tiles >= (n_dst + 15) // 16
    if graph != "random":
        assert plan.reuse > 1.5  # the tiles really share sources
    out = torch.full((n_dst, C + H * dp), float("nan"), dtype=dt, device="cuda")
    args = dict(edge_attr=a16.cuda(), b_edge=b_e.cuda(), qw=buf[:, 2 * C :], dp=dp, add=buf[:, C : 2 * C])
    ops.gt_attention(buf[:, :C], k.cuda(), v.cuda(), csr, H, abar=out[:, C:], out=out[:, :C], tiles=plan, **args)
    out_pipe = torch.empty_like(out)
    ops.gt_attention(buf[:, :C], k.cuda(), v.cuda(), csr, H, abar=out_pipe[:, C:], out=out_pipe[:, :C], **args)
    o = out.float().cpu()
    assert torch.isfinite(o).all()
    abar = o[:, C:].view(n_dst, H, dp)[:, :, :d_e]
    full = o[:, :C] + torch.einsum("nha,hca->nhc", abar, w_e.view(H, Ch, d_e)).reshape(n_dst, C)
    sh = lambda t, n: t.float().view(n, H, Ch)
    ref = R.gt_attention(sh(q32, n_dst), sh(k, n_src), sh(v, n_src), (a @ w_e.t() + b_e).view(E, H, Ch), ei, n_dst).view(n_dst, C)
    ref = ref + buf[:, C : 2 * C].float().cpu()
    tol = 2**-5 * ref.abs().max().item()
    assert (full - ref).abs().max().item() <= tol
    assert torch.equal(o[-5:, :C], buf[-5:, C : 2 * C].float().cpu())  # zero in-degree: 0 + add, no bias
    assert torch.all(o[-5:, C:] == 0)
    # the two kernels agree to bf16 rounding of the outputs (the tile kernel rounds the softmax weights to bf16 before P.V)
    op = out_pipe.float().cpu()
    assert (o[:, :C] - op[:, :C]).abs().max().item() <= 2**-6 * ref.abs().max().item()
    assert (o[:, C:] - op[:, C:]).abs().max().item() <= 2**-6 * max(op[:, C:].abs().max().item(), 1.0)


def test_attention_tile_plan_rejects_wide_destinations(ops):
    """A destination with more than 64 distinct sources has no tile plan (the caller keeps the warp-per-node kernel)."""
    ei = torch.stack([torch.arange(100), torch.zeros(100, dtype=torch.long)])
    csr = ops.build_csr(ei.cuda(), 100, 3)
    assert ops.attention_tiles(csr) is None and csr.tiles is False


def test_gt_attention_strided_qkv(ops):
    """q|k|v|self as column slices of one [N, 4C] buffer (the block layout)."""
    g = torch.Generator().manual_seed(9)
    N, E, H, Ch = 400, 3000, 16, 32
    C = H * Ch
    ei = _rand_graph(N, N, E, 21)
    buf = torch.randn(N, 4 * C, generator=g).to(torch.bfloat16).cuda()
    a_pad = torch.randn(E, 12, generator=g)
    a_pad[:, 11] = 0
    w_e, b_e = torch.randn(C, 11, generator=g) / 3, torch.randn(C, generator=g)
    csr = ops.build_csr(ei.cuda(), N, N)
    out = ops.gt_attention(buf[:, :C], buf[:, C : 2 * C], buf[:, 2 * C : 3 * C], csr, H, edge_attr=a_pad.cuda(), w_edge=w_e.cuda(), b_edge=b_e.cuda(),
                           add=buf[:, 3 * C :])  # fmt: skip
    b = buf.float().cpu()
    sh = lambda t: t.reshape(-1, H, Ch)
    ref = R.gt_attention(sh(b[:, :C]), sh(b[:, C : 2 * C]), sh(b[:, 2 * C : 3 * C]), sh(a_pad[:, :11] @ w_e.t() + b_e), ei, N).view(N, C) + b[:, 3 * C :]
    assert (out.float().cpu() - ref).abs().max().item() <= 2**-6 * ref.abs().max().item()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [32, 64, 512, 1024, 100])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_graphconv_ln_aggregate(ops, C, dt):
    g = torch.Generator().manual_seed(C)
    n_src, n_dst, E = 200, 150, 1300
    ei = _rand_graph(n_src, n_dst, E, 5, zero_tail=4)
    h = (torch.randn(E, C, generator=g) * 1.5).to(dt)
    e = torch.randn(E, C, generator=g).to(dt)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    e_new, out = ops.graphconv_ln_aggregate(h.cuda(), w.cuda(), b.cuda(), e.cuda(), csr)
    ref_e = (torch.nn.functional.layer_norm(h.float(), (C,), w, b, 1e-5) + e.float()).to(dt).float()
    ref_out = R.scatter_sum(ref_e, ei[1], n_dst)
    tol = 2e-5 if dt == torch.float32 else 2**-7
    torch.testing.assert_close(e_new.float().cpu(), ref_e, atol=4 * tol, rtol=2 * tol)
    torch.testing.assert_close(out.float().cpu(), ref_out, atol=16 * tol * (1 if dt == torch.float32 else 4), rtol=2 * tol)
    assert torch.all(out[-4:] == 0)


def test_cast_pad(ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(100, 13, generator=g)
    idx = torch.randint(0, 100, (37,), generator=g, dtype=torch.int32)
    y = ops.cast_pad(x.cuda(), torch.bfloat16, 16, idx=idx.cuda())
    ref = torch.zeros(37, 16)
    ref[:, :13] = x[idx.long()]
    assert torch.equal(y.float().cpu(), ref.to(torch.bfloat16).float())


@pytest.mark.parametrize("M,K,Kpad", [(40320, 212, 216), (1000, 12, 64), (333, 64, 64), (7, 4, 8), (50, 13, 16)])
@pytest.mark.parametrize("dti,dto", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                     (torch.bfloat16, torch.bfloat16)])  # fmt: skip
def test_cast_pad_vectorised_and_scalar_paths(ops, M, K, Kpad, dti, dto):
    """The 4-columns-per-thread path (K, Kpad, strides multiples of 4) and the scalar path give the same exact copy / rounding."""
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(M + K)).to(dti)
    y = ops.cast_pad(x.cuda(), dto, Kpad)
    ref = torch.zeros(M, Kpad, dtype=dto)
    ref[:, :K] = x.to(dto)
    assert y.shape == (M, Kpad) and torch.equal(y.cpu(), ref)


@pytest.mark.parametrize("M,C", [(40962, 512), (100, 8), (33, 20)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_add(ops, M, C, dt):
    g = torch.Generator().manual_seed(M)
    a, b = torch.randn(M, C, generator=g).to(dt), torch.randn(M, C, generator=g).to(dt)
    y = ops.add(a.cuda(), b.cuda())
    assert torch.equal(y.cpu(), (a.float() + b.float()).to(dt))
    big = torch.randn(M, 2 * C + 8, generator=g).to(dt).cuda()  # strided operands (column slices)
    y2 = ops.add(big[:, 8 : 8 + C], b.cuda())
    assert torch.equal(y2.cpu(), (big[:, 8 : 8 + C].float().cpu() + b.float()).to(dt))


def test_cpu_tensor_is_an_error(ops):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layer_norm(torch.randn(4, 8), None, None)


def test_reverse_traversal_hint_changes_nothing(ops):
    """``ops.set_traversal`` (ANEMOI_EPI_REVERSE: GEMM row blocks / attention destination ranges / row-statistics rows walked bottom-up, an
    L2 scheduling hint of the GraphTransformer block) must give bit-identical results."""
    g = torch.Generator().manual_seed(11)
    M, K, N = 5000, 512, 768
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K**0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    n, e, H, C = 3000, 24000, 16, 512
    ei = _rand_graph(n, n, e, 3, zero_tail=7)
    csr = ops.build_csr(ei.cuda(), n, n)
    qkv = torch.randn(n, 3 * C, generator=g).to(torch.bfloat16).cuda()
    ep = torch.randn(e, C, generator=g).to(torch.bfloat16).cuda()
    res = {}
    for rev in (False, True):
        ops.set_traversal(gemm=rev, stats=rev)
        try:
            res[rev] = (ops.linear(a, w, b, gelu=True), ops.linear(a, w, b, residual=r), ops.row_stats(a),
                        ops.gt_attention(qkv[:, :C], qkv[:, C : 2 * C], qkv[:, 2 * C :], csr, H, e_proj=ep))
        finally:
            ops.set_traversal()
    for x, y in zip(res[False], res[True]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("M,N,K", [(4096, 512, 512), (40962, 2048, 512), (1000, 88, 512), (5000, 512, 2048)])
def test_fp32_linear_on_tensor_cores_is_fp32_grade(ops, M, N, K):
    """fp32 GEMMs run as ONE bf16 tcgen05 GEMM over the exact three-way bf16 split of both operands (6 partial products, fp32
    accumulation): the result must be as good as an fp32 FFMA evaluation — compared here with a float64 reference to 5e-6 of the output scale (K up to 2048),
    with bias / GELU / residual epilogues — and the FFMA kernel (ANEMOI_B200_FP32_TC off) must agree."""
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))  # rows of very different scale
    w = torch.randn(N, K, generator=g) / K**0.5
    b, r = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    ref = (a.double() @ w.double().t() + b.double())
    assert ops._fp32_on_tensor_cores(a.cuda(), w.cuda()) is not None
    y = ops.linear(a.cuda(), w.cuda(), b.cuda()).cpu().double()
    assert ((y - ref).abs().max() / ref.abs().max()).item() <= 5e-6
    row_scale = ref.abs().amax(1, keepdim=True)
    assert ((y - ref).abs() / row_scale).max().item() <= 1e-5  # also for the small rows
    y2 = ops.linear(a.cuda(), w.cuda(), b.cuda(), gelu=True, residual=r.cuda()).cpu().double()
    ref2 = torch.nn.functional.gelu(ref) + r.double()
    assert ((y2 - ref2).abs().max() / ref2.abs().max()).item() <= 5e-6
    old = ops.FP32_TC
    try:
        ops.FP32_TC = False
        y3 = ops.linear(a.cuda(), w.cuda(), b.cuda()).cpu().double()
    finally:
        ops.FP32_TC = old
    assert ((y3 - ref).abs().max() / ref.abs().max()).item() <= 5e-6


def test_linear_tail_wave_split_matches_single_launch(ops):
    """Narrow GEMMs (N <= 512) on cfg2-sized row counts run their last, partially filled wave as a second launch on a forked stream
    (``ops.TAIL_SPLIT``): same result as the single launch, also under CUDA-graph capture."""
    g = torch.Generator().manual_seed(3)
    M, N, K = 40962, 512, 704
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K**0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    old = ops.TAIL_SPLIT
    try:
        ops.TAIL_SPLIT = False
        y_one = ops.linear(a, w, b, residual=r)
        ops.TAIL_SPLIT = True
        assert ops._tail_split_rows(a, w, {}) == 37888
        y_split = ops.linear(a, w, b, residual=r)
        assert torch.equal(y_split, y_one)
        out = torch.empty_like(y_one)
        ops.linear(a, w, b, residual=r, out=out)  # warm-up outside the capture (descriptor cache, side stream)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops.linear(a, w, b, residual=r, out=out)
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, y_one)
    finally:
        ops.TAIL_SPLIT = old


@pytest.mark.parametrize("M,N,K", [(5000, 512, 512), (1000, 200, 216), (300, 1024, 1024)])
def test_linear_gather_tables_bf16(ops, M, N, K):
    """Gather-add epilogue with bf16 tables (ANEMOI_EPI_G1_BF16 / _G2_BF16; GraphConv's src-indexed term on the bf16 path): same result as with
    the fp32 copies of the same (bf16-valued) tables, for one or both tables in bf16."""
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K**0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    ld = (N + 7) // 8 * 8
    t1 = torch.randn(70, ld, generator=g).to(torch.bfloat16).cuda()
    t2 = torch.randn(90, ld, generator=g).to(torch.bfloat16).cuda()
    i1 = torch.randint(0, 70, (M,), generator=g, dtype=torch.int32).cuda()
    i2 = torch.randint(0, 90, (M,), generator=g, dtype=torch.int32).cuda()
    sl = (lambda t: t[:, :N] if ld != N else t)
    ref = ops.linear(a, w, b, gelu=True, gather1=(sl(t1.float()), i1), gather2=(sl(t2.float()), i2))
    for g1, g2 in ((sl(t1.float()), sl(t2)), (sl(t1), sl(t2))):
        y = ops.linear(a, w, b, gelu=True, gather1=(g1, i1), gather2=(g2, i2))
        assert torch.equal(y, ref)
