"""GPU: every kernel behind the C ABI against a plain PyTorch fp32 statement of the same op.
16-bit-operand kernels are compared on the same (already rounded) inputs, so the tolerance only covers the
output rounding; fp32-output kernels are held to fp32 round-off."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max()).item()


def _gemm_inputs(M, N, K, dtype, ints=False, seed=0):
    g = torch.Generator().manual_seed(seed + M + 3 * N + 7 * K)
    if ints:
        a = torch.randint(-2, 3, (M, K), generator=g).to(dtype)
        w = torch.randint(-2, 3, (N, K), generator=g).to(dtype)
    else:
        a = (torch.randn(M, K, generator=g) * 0.5).to(dtype)
        w = (torch.randn(N, K, generator=g) * 0.05).to(dtype)
    bias = torch.randn(N, generator=g) * 0.1
    gamma = torch.rand(N, generator=g) + 0.5
    return a.to(DEV), w.to(DEV), bias.to(DEV), gamma.to(DEV), g


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 256, 128), (1, 128, 64), (129, 768, 768)])
def test_gemm_exact_on_integer_operands(M, N, K):
    """Small-integer operands make every product and sum exact: any mismatch is a layout/descriptor bug."""
    from keep_b200 import ops

    a, w, bias, _, _ = _gemm_inputs(M, N, K, torch.float16, ints=True)
    bias = bias.round()
    out = ops.gemm(a, w, ops.EPI_BIAS_F32, bias=bias)
    assert torch.equal(out, a.float() @ w.float().T + bias)


@pytest.mark.parametrize("M,N,K", [(197 * 3, 768, 1024), (128 * 160, 1024, 1024), (197 * 64, 3072, 1024), (50, 2304, 768)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_bias_16bit(M, N, K, dtype):
    from keep_b200 import ops

    a, w, bias, _, _ = _gemm_inputs(M, N, K, dtype)
    out = ops.gemm(a, w, ops.EPI_BIAS_HALF, bias=bias)
    ref = a.float() @ w.float().T + bias
    assert out.dtype == dtype
    assert _rel(out, ref) < (1e-3 if dtype == torch.float16 else 8e-3)


def test_gemm_gelu_is_exact_erf():
    from keep_b200 import ops

    a, w, bias, _, _ = _gemm_inputs(197 * 16, 4096, 1024, torch.float16)
    out = ops.gemm(a, w, ops.EPI_BIAS_GELU_HALF, bias=bias)
    pre = a.float() @ w.float().T + bias
    assert _rel(out, F.gelu(pre)) < 1e-3
    # the tanh approximation differs from erf-GELU by up to 4.7e-4, systematically: visible in the mean error
    err_erf = (out.float() - F.gelu(pre)).abs().mean().item()
    err_tanh = (out.float() - F.gelu(pre, approximate="tanh")).abs().mean().item()
    assert err_erf < 0.8 * err_tanh, (err_erf, err_tanh)
    assert abs((out.float() - F.gelu(pre)).mean().item()) < 5e-6


def test_gemm_residual_layerscale_in_place():
    from keep_b200 import ops

    M, N, K = 197 * 16, 1024, 4096
    a, w, bias, gamma, g = _gemm_inputs(M, N, K, torch.float16)
    resid = torch.randn(M, N, generator=g).to(DEV)
    ref = resid + gamma * (a.float() @ w.float().T + bias)
    x = resid.clone()
    ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=x, out=x)
    assert _rel(x, ref) < 1e-5
    y = ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=None, resid=resid)  # BERT form: no LayerScale
    assert _rel(y, resid + a.float() @ w.float().T + bias) < 1e-5


def test_gemm_patch_embed_scatter():
    from keep_b200 import ops

    P, B, N = 196, 5, 1024
    a, w, bias, _, g = _gemm_inputs(P * B, N, 768, torch.float16)
    pos = torch.randn(P + 1, N, generator=g).to(DEV)
    out = ops.gemm(a, w, ops.EPI_PATCH_F32, bias=bias, pos=pos, patches=P)
    ref = torch.zeros(B, P + 1, N, device=DEV)
    ref[:, 1:] = (a.float() @ w.float().T + bias).view(B, P, N) + pos[1:]
    assert _rel(out, ref.view(-1, N)) < 1e-5
    assert torch.equal(out.view(B, P + 1, N)[:, 0], torch.zeros(B, N, device=DEV))  # CLS rows untouched


def test_gemm_strided_rows():
    """A operand with a row pitch (the [CLS]-row gather of the BERT pooler)."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = (torch.randn(9, 20, 768, generator=g) * 0.5).half().to(DEV)
    w = (torch.randn(768, 768, generator=g) * 0.05).half().to(DEV)
    a = x[:, 0, :]  # stride 20*768
    out = ops.gemm(a, w, ops.EPI_BIAS_F32)
    assert _rel(out, a.float() @ w.float().T) < 1e-5


def test_gemm_rejects_bad_shapes():
    from keep_b200 import KeepB200Error, ops

    a = torch.zeros(8, 100, dtype=torch.float16, device=DEV)
    w = torch.zeros(128, 100, dtype=torch.float16, device=DEV)
    with pytest.raises(KeepB200Error, match="multiple of 64"):
        ops.gemm(a, w, ops.EPI_BIAS_F32)


def test_layernorm_variants():
    from keep_b200 import ops

    x = torch.randn(1000, 1024, device=DEV) * 3 + 1
    w, b = torch.rand(1024, device=DEV) + 0.5, torch.randn(1024, device=DEV)
    y16, y32 = ops.layernorm(x, w, b, 1e-6, want_f32=True)
    ref = F.layer_norm(x, (1024,), w, b, 1e-6)
    assert _rel(y32, ref) < 1e-5 and _rel(y16, ref) < 1e-3
    x2 = torch.randn(77, 768, device=DEV)
    w2, b2 = torch.rand(768, device=DEV) + 0.5, torch.randn(768, device=DEV)
    _, y = ops.layernorm(x2, w2, b2, 1e-12, want_f32=True)
    assert _rel(y, F.layer_norm(x2, (768,), w2, b2, 1e-12)) < 1e-5
    x3 = torch.randn(5, 197, 1024, device=DEV)
    _, y = ops.layernorm(x3, w, b, 1e-6, want_f32=True, rows=5, row_stride=197 * 1024)
    assert _rel(y, F.layer_norm(x3[:, 0], (1024,), w, b, 1e-6)) < 1e-5
    yb, _ = ops.layernorm(x, w, b, 1e-6, out_dtype=torch.bfloat16)
    assert yb.dtype == torch.bfloat16 and _rel(yb, ref) < 8e-3


def test_l2norm_and_tanh():
    from keep_b200 import ops

    x = torch.randn(33, 768, device=DEV)
    assert _rel(ops.act_l2norm(x, 0), F.normalize(x, dim=-1)) < 1e-6
    assert _rel(ops.act_l2norm(x, 1), F.normalize(torch.tanh(x), dim=-1)) < 1e-6
    z = torch.zeros(2, 768, device=DEV)
    assert torch.equal(ops.act_l2norm(z, 0), z)  # eps clamp: 0 / max(0, 1e-12) = 0, no NaN


@pytest.mark.parametrize("B,S,H,masked", [(3, 197, 16, False), (5, 256, 12, True), (7, 32, 12, True), (2, 77, 4, True),
                                            (1, 1, 2, False), (2, 300, 2, True), (150, 1, 2, True), (23, 5, 3, True),
                                            (9, 12, 12, True), (40, 33, 2, False), (3, 56, 4, True), (3, 57, 4, True),
                                            (5, 64, 2, False), (2, 112, 3, True), (2, 113, 3, True), (3, 224, 2, False),
                                            (2, 225, 2, True), (1, 512, 2, True), (2, 485, 2, False), (300, 197, 1, False),
                                            (4, 240, 3, False), (3, 241, 2, True), (200, 256, 12, False), (2, 257, 2, True)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention(B, S, H, masked, dtype):
    """Every shape class of the tcgen05 kernels: packed sequences (S <= 56, whole sequences of a head share one tile,
    ragged last group), the single-tile kernel (65..224 keys with a shared O accumulator, 225..256 keys with O inside the
    unit's own TMEM region; many units per CTA), two to five KV blocks of up to 112 keys beyond, ragged last query tile,
    key masks of random lengths."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(S * 3 + H)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).to(dtype).to(DEV)
    mask = None
    bias = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(DEV)
        bias = torch.zeros(B, 1, 1, S, device=DEV).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    out = ops.attention(qkv, B, S, H, key_mask=mask)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=0.125).transpose(1, 2).reshape(B * S, H * 64)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < (2e-3 if dtype == torch.float16 else 1.5e-2)
    assert torch.equal(out, ops.attention(qkv, B, S, H, key_mask=mask))  # deterministic


@pytest.mark.parametrize("S,hot", [(197, (150, 196)), (256, (120, 250)), (300, (299, 230)), (40, (39, 20)), (512, (100, 460))])
def test_attention_late_maximum_in_a_later_block(S, hot):
    """The lazy softmax reference: rows whose maximum sits in a LATER KV block, ~2^30 above everything before it, after the
    earlier blocks were already multiplied into O (the O-rescale path), and in a late chunk of the same block."""
    from keep_b200 import ops

    B, H = 2, 8
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 0.5).view(B, S, 3, H, 64)
    qkv[:, :, 0, 3, :] = 3.0
    qkv[:, hot[0], 1, 3, :] = 4.0        # one key aligned with every query of head 3
    qkv[:, :, 0, 5, :] = -2.0
    qkv[:, hot[1], 1, 5, :] = -5.0       # ... and another one for head 5
    qkv[:, : S // 2, 0, 6, :] = 2.5       # head 6: only the first half of the queries sees the spike
    qkv[:, hot[0], 1, 6, :] = 3.0
    qkv = qkv.reshape(B * S, 3 * H * 64).half().to(DEV)
    out = ops.attention(qkv, B, S, H)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v, scale=0.125).transpose(1, 2).reshape(B * S, H * 64)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("tensor_cores", [True, False])
def test_similarity_and_group_softmax(tensor_cores):
    """TF32 tcgen05 kernel (10-bit operand mantissa: cosine error ~3e-5, gate 1e-3) and the fp32 FMA kernel."""
    from keep_b200 import ops

    tol_l, tol_p = (2e-4, 1e-3) if tensor_cores else (1e-5, 1e-5)
    feats = torch.randn(1000, 768, device=DEV) * 2
    cls = F.normalize(torch.randn(768, 32, device=DEV), dim=0)
    logits, probs = ops.similarity(feats, cls, group=2, temp=10.0, tensor_cores=tensor_cores)
    ref = F.normalize(feats, dim=-1) @ cls
    assert (logits - ref).abs().max().item() < tol_l
    assert (probs - torch.softmax(ref.view(1000, 16, 2) * 10, -1).view(1000, 32)).abs().max().item() < tol_p
    for P, group in ((2, 0), (3, 0), (4, 4), (256, 4), (70, 0), (640, 2), (48, 16), (96, 32)):  # ragged prompt counts / groupings
        c = F.normalize(torch.randn(768, P, device=DEV), dim=0)
        lg, pr = ops.similarity(feats[:333], c, group=group, tensor_cores=tensor_cores)
        r = F.normalize(feats[:333], dim=-1) @ c
        g = group or P
        assert (lg - r).abs().max().item() < tol_l, (P, group)
        assert (pr - torch.softmax(r.view(333, P // g, g) * 10, -1).view(333, P)).abs().max().item() < tol_p, (P, group)
    lg, pr = ops.similarity(feats[:0], cls, tensor_cores=tensor_cores)
    assert lg.shape == (0, 32)
    big = torch.randn(20000, 768, device=DEV)
    lg, _ = ops.similarity(big, cls, group=2, tensor_cores=tensor_cores)  # many tiles per CTA: persistent loop, TMEM double buffer
    assert (lg - F.normalize(big, dim=-1) @ cls).abs().max().item() < tol_l
    zero = torch.zeros(5, 768, device=DEV)
    lg, pr = ops.similarity(zero, cls[:, :2], tensor_cores=tensor_cores)   # eps clamp: 0 / max(0, 1e-12)
    assert torch.equal(lg, torch.zeros(5, 2, device=DEV)) and (pr - 0.5).abs().max().item() < 1e-6


def _screen_ref(feats, cls, K, C):
    lg = (F.normalize(feats.double(), dim=-1) @ cls.double()).view(feats.shape[0], K, C)
    top = lg.topk(2, dim=2).values
    return ((top[..., 0] - top[..., 1]) - (top[..., 0] + top[..., 1] - 1).abs()).mean(0).float()


def test_prompt_scores_chunked():
    """The default path: similarity GEMM -> logits chunk -> top-2 margin reduction, several row chunks."""
    from keep_b200 import ops

    N, K, C = 3000, 70, 4
    feats = torch.randn(N, 768, device=DEV)
    cls = F.normalize(torch.randn(768, K * C, device=DEV), dim=0)
    s = ops.prompt_scores(feats, cls, K, C, workspace_mb=1)  # forces several row chunks
    assert (s - _screen_ref(feats, cls, K, C)).abs().max().item() < 1e-4
    s3 = ops.prompt_scores(feats, cls[:, :69].contiguous(), 23, 3)  # a class count that does not divide 16
    assert (s3 - _screen_ref(feats, cls[:, :69], 23, 3)).abs().max().item() < 1e-4


@pytest.mark.parametrize("N,K,C", [(3000, 70, 4), (10_000, 1386, 2), (777, 5, 2), (4097, 33, 8), (130, 3, 16)])
def test_prompt_scores_fused_epilogue(N, K, C):
    """fused=True: the top-2 margin reduced inside the similarity epilogue (utils.py:107-130 in one kernel, no [N, K*C]
    logits in memory): ragged row counts, ragged column tiles, every supported class count; deterministic."""
    from keep_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(N + K)
    feats = torch.randn(N, 768, device=DEV, generator=g) * 2
    cls = F.normalize(torch.randn(768, K * C, device=DEV, generator=g), dim=0)
    s = ops.prompt_scores(feats, cls, K, C, fused=True)
    assert (s - _screen_ref(feats, cls, K, C)).abs().max().item() < 1e-4
    assert torch.equal(s, ops.prompt_scores(feats, cls, K, C, fused=True))
    d = ops.prompt_scores(feats, cls, K, C, workspace_mb=4)  # default path, chunked: fixed-order partials, no atomics
    assert (d - s).abs().max().item() < 1e-5
    assert torch.equal(d, ops.prompt_scores(feats, cls, K, C, workspace_mb=4))


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("C", [2, 4])
def test_refine_bit_exact_vs_dict_walk(overlap, C):
    """Integer/index work and the float32 neighbour mean must equal the reference's dict walk bit for bit."""
    from keep_b200 import ops
    from oracle import wsi_oracle as wo

    g = torch.Generator().manual_seed(5 + C)
    N = 20000
    xy = torch.randint(0, 90, (N, 2), generator=g) * 224  # duplicates (first wins) and holes
    probs = torch.softmax(torch.randn(N, C, generator=g), 1)
    keep, refined = ops.refine(xy.to(DEV), probs.to(DEV), 224, overlap)
    exp = wo.refine_mean(probs.numpy(), xy.numpy(), 224, overlap)
    first = wo._first_occurrence(xy.numpy())
    exp_keep = np.zeros(N, dtype=np.uint8)
    exp_keep[list(first.values())] = 1
    assert np.array_equal(keep.cpu().numpy(), exp_keep)
    got = refined.cpu().numpy()
    for (x, y), i in first.items():
        assert np.array_equal(got[i], exp[(x, y)]), (x, y)


# ---- LayerNorm fused across two GEMMs (EPI_RESID_F32_STATS -> EPI_LN_*) ------------------------------------------
@pytest.mark.parametrize("M,N,K", [(197 * 3, 1024, 1024), (197 * 40 + 5, 1024, 4096), (77, 256, 128)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_resid_stats_matches_residual_and_row_sums(M, N, K, dtype):
    """The stats-producing residual epilogue: same x as EPI_RESID_F32, x16 = 16-bit(x), and per-row partial sums
    over 64-column slices that add up to the row sum / sum of squares of the NEW x."""
    from keep_b200 import ops

    a, w, bias, gamma, g = _gemm_inputs(M, N, K, dtype)
    resid = torch.randn(M, N, generator=g).to(DEV)
    plain = ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=resid.clone())
    x = resid.clone()
    x16, stats = ops.gemm_resid_stats(a, w, x, bias=bias, gamma=gamma)
    assert torch.equal(x, plain)                       # bit-identical fp32 residual update
    assert torch.equal(x16, x.to(dtype))               # round-to-nearest copy
    xs = x.view(M, N // 64, 64)
    assert torch.allclose(stats[..., 0], xs.sum(-1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(stats[..., 1], (xs * xs).sum(-1), rtol=1e-5, atol=1e-4)
    # deterministic: fixed reduction order, no atomics
    x2 = resid.clone()
    _, stats2 = ops.gemm_resid_stats(a, w, x2, bias=bias, gamma=gamma)
    assert torch.equal(stats, stats2)


@pytest.mark.parametrize("M,N,K,gelu", [(197 * 3, 3072, 1024, False), (197 * 40 + 5, 4096, 1024, True),
                                         (197 * 70, 3072, 1024, False), (61, 256, 128, True)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_ln_equals_layernorm_then_linear(M, N, K, gelu, dtype):
    """fold_ln + EPI_LN_*: act(Linear(LayerNorm(x))) computed from 16-bit(x), the row statistics of x and the
    folded weight, against the fp32 statement; also against the two-kernel product path (layernorm -> gemm), whose
    error it must not exceed by more than rounding noise."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g) * 1.7 + 0.05 * torch.randn(M, 1, generator=g)).to(DEV)
    x[:, 3] += 6.0                                         # one heavy channel, as ViT residual streams have
    w32 = (torch.randn(N, K, generator=g) * 0.03).to(DEV)
    bias = (torch.randn(N, generator=g) * 0.1).to(DEV)
    lnw = (1 + 0.1 * torch.randn(K, generator=g)).to(DEV)
    lnb = (0.05 * torch.randn(K, generator=g)).to(DEV)
    eps = 1e-6
    ref = F.layer_norm(x.double(), (K,), lnw.double(), lnb.double(), eps) @ w32.double().T + bias.double()
    if gelu:
        ref = F.gelu(ref)
    # statistics as the producer leaves them: per 64-column slice
    xs = x.view(M, K // 64, 64)
    stats = torch.stack([xs.sum(-1), (xs * xs).sum(-1)], dim=-1).contiguous()
    wf, s, c = ops.fold_ln(w32, lnw, lnb, bias, dtype=dtype)
    assert torch.equal(wf, (w32 * lnw).to(dtype))
    assert torch.allclose(s, wf.float().sum(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(c, (bias.double() + w32.double() @ lnb.double()).float(), rtol=1e-5, atol=1e-5)
    out = ops.gemm_ln(x.to(dtype), wf, s, c, stats, eps, gelu=gelu)
    xn = ops.layernorm(x, lnw, lnb, eps, out_dtype=dtype)[0]
    two = ops.gemm(xn, w32.to(dtype), ops.EPI_BIAS_GELU_HALF if gelu else ops.EPI_BIAS_HALF, bias=bias)
    e_fused = ((out.double() - ref).norm() / ref.norm()).item()
    e_two = ((two.double() - ref).norm() / ref.norm()).item()
    tol = 1.5e-3 if dtype == torch.float16 else 1.2e-2
    assert e_fused < tol, (e_fused, e_two)
    assert e_fused < 1.5 * e_two + 1e-4, (e_fused, e_two)


@pytest.mark.parametrize("g0,gh,gw", [(14, 16, 16), (14, 10, 14), (14, 7, 22), (14, 32, 15), (4, 14, 14), (14, 14, 14)])
def test_pos_embed_resample_matches_torch_bicubic_antialias(g0, gh, gw):
    """timm resample_abs_pos_embed = F.interpolate(mode='bicubic', antialias=True) on the grid rows, prefix row kept."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(g0 * 100 + gh * 10 + gw)
    D = 256
    pos = torch.randn(1, 1 + g0 * g0, D, generator=g)
    grid = pos[:, 1:].reshape(1, g0, g0, D).permute(0, 3, 1, 2)
    ref = torch.cat([pos[:, :1], F.interpolate(grid, size=(gh, gw), mode="bicubic", antialias=True)
                     .permute(0, 2, 3, 1).reshape(1, gh * gw, D)], dim=1)[0]
    out = ops.pos_resample(pos.to(DEV), g0, gh, gw).cpu()
    assert out.shape == ref.shape
    assert torch.equal(out[0], pos[0, 0])
    assert (out - ref).abs().max().item() < 2e-6, (out - ref).abs().max().item()


def test_preprocess_resize_center_crop_bit_exact_vs_pil(golden_dir):
    """Resize(224, BICUBIC) + CenterCrop(224) on uint8 tiles (keep_inference.py:88-90): bit-identical to the real
    torchvision + Pillow output (tests/golden/transform.npz), for down-/up-sampling, both orientations, crop-only and
    identity, and for batches."""
    from keep_b200.transform import preprocess
    from oracle.make_golden import transform_inputs

    g = np.load(f"{golden_dir}/transform.npz")
    for i, tile in enumerate(transform_inputs()):
        t = torch.from_numpy(tile).to(DEV)
        out = preprocess(torch.stack([t, t.flip(0), t]))
        assert out.shape == (3, 224, 224, 3) and out.dtype == torch.uint8
        assert np.array_equal(out[0].cpu().numpy(), g[f"u8_{i}"]), (i, tile.shape)
        assert torch.equal(out[0], out[2]) and not (tile.shape[0] != 224 and torch.equal(out[0], out[1]))
    ex = torch.from_numpy(g["example_raw"]).to(DEV)
    assert np.array_equal(preprocess(ex[None])[0].cpu().numpy(), g["example_u8"])
    assert preprocess(torch.zeros(0, 300, 300, 3, dtype=torch.uint8, device=DEV)).shape == (0, 224, 224, 3)


def test_gemm_is_deterministic_on_every_epilogue():
    """One code path per shape (no run-time variant switches): two launches of the pair kernel agree bit for bit."""
    from keep_b200 import ops

    M, N, K = 197 * 130 + 7, 1024, 1024
    a, w, bias, gamma, g = _gemm_inputs(M, N, K, torch.float16)
    resid = torch.randn(M, N, generator=g).to(DEV)
    w4 = (torch.randn(4096, K, generator=g) * 0.05).half().to(DEV)
    b4 = (torch.randn(4096, generator=g) * 0.1).to(DEV)

    def run():
        out = {}
        out["bias"] = ops.gemm(a, w, ops.EPI_BIAS_HALF, bias=bias)
        out["gelu"] = ops.gemm(a, w4, ops.EPI_BIAS_GELU_HALF, bias=b4)
        out["resid"] = ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=resid.clone())
        x = resid.clone()
        out["x16"], out["stats"] = ops.gemm_resid_stats(a, w, x, bias=bias, gamma=gamma)
        out["x"] = x
        return out

    ref, got = run(), run()
    for k in ref:
        assert torch.equal(ref[k], got[k]), k
    assert _rel(ref["bias"], a.float() @ w.float().T + bias) < 1e-3


def _rel_l2(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm()).item()


@pytest.mark.parametrize("M,N,K", [(3, 768, 768), (197, 3072, 768), (1000, 768, 3072), (128 * 160, 1024, 1024), (7, 128, 256)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_split_operands_reach_fp32_accuracy(M, N, K, dtype):
    """Split-operand GEMM (hi|lo activations and weights, 3 passes; hi-only activations, 2 passes) against an fp64 product
    of the UNROUNDED fp32 operands: the plain 16-bit GEMM sits at the operand rounding (~3e-4 fp16, ~2.5e-3 bf16), the
    split ones must be two to three orders of magnitude closer. Covers all three main-loop variants (M sweeps them)."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    a32 = (torch.randn(M, K, generator=g) * 0.7).to(DEV)
    w32 = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = (torch.randn(N, generator=g) * 0.1).to(DEV)
    ref = a32.double() @ w32.double().T + bias.double()
    a_hl, w_hl = ops.cast_hilo(a32, dtype), ops.cast_hilo(w32, dtype)
    assert torch.equal(a_hl[:, :K], a32.to(dtype))                      # hi half = the plain rounding
    lo_ref = (a32 - a32.to(dtype).float()).to(dtype)
    assert torch.equal(a_hl[:, K:], lo_ref)                             # lo half = rounded remainder
    plain = ops.gemm(a_hl[:, :K], w_hl[:, :K], ops.EPI_BIAS_F32, bias=bias)   # row pitch 2K, hi halves only
    aw = ops.gemm_split(a_hl, w_hl, K, ops.EPI_BIAS_F32, ops.SPLIT_AW, bias=bias)
    wonly = ops.gemm_split(a_hl[:, :K], w_hl, K, ops.EPI_BIAS_F32, ops.SPLIT_W, bias=bias)
    e_plain, e_aw = _rel_l2(plain, ref), _rel_l2(aw, ref)
    ref_w = a32.to(dtype).double() @ w32.double().T + bias.double()      # W exact, A rounded
    e_w = _rel_l2(wonly, ref_w)
    # what is left: the lo.lo term (2^-22 fp16 / 2^-16 bf16 relative) and, for fp16, lo parts of small weights that fall
    # into the subnormal range (|lo| <= 2^-12 |w| < 6e-5: absolute step 6e-8) - both two orders below the 2^-11 rounding
    tol = 3e-5
    print(f"{M}x{N}x{K} {dtype}: plain {e_plain:.2e}  split-AW {e_aw:.2e}  split-W (vs A rounded) {e_w:.2e}")
    assert e_aw < tol and e_w < tol and e_plain > 10 * e_aw


def test_gemm_gelu_hilo_epilogue_and_residual_split():
    """The text-tower MLP in split mode: x[hi|lo] -> GELU(x W1^T + b1) as [hi|lo] -> resid + h W2^T + b2, vs fp64."""
    from keep_b200 import ops

    M, d, I = 300, 768, 3072
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, d, generator=g).to(DEV)
    w1 = (torch.randn(I, d, generator=g) / d ** 0.5).to(DEV)
    w2 = (torch.randn(d, I, generator=g) / I ** 0.5).to(DEV)
    b1, b2 = (torch.randn(I, generator=g) * 0.1).to(DEV), (torch.randn(d, generator=g) * 0.1).to(DEV)
    resid = torch.randn(M, d, generator=g).to(DEV)
    h = ops.gemm_split(ops.cast_hilo(x), ops.cast_hilo(w1), d, ops.EPI_BIAS_GELU_HILO, ops.SPLIT_AW, bias=b1)
    assert h.shape == (M, 2 * I)
    h_ref = F.gelu(x.double() @ w1.double().T + b1.double())
    assert _rel_l2(h[:, :I].float() + h[:, I:].float(), h_ref) < 1e-5   # hi + lo carries the value to ~20 bits
    assert _rel_l2(h[:, :I], h_ref) > 1e-4                              # ... which the hi half alone does not
    out = ops.gemm_split(h, ops.cast_hilo(w2), I, ops.EPI_RESID_F32, ops.SPLIT_AW, bias=b2, resid=resid.clone())
    assert _rel_l2(out, resid.double() + h_ref @ w2.double().T + b2.double()) < 3e-5  # K = 3072: fp16 lo parts of the small weights are subnormal


def test_layernorm_and_attention_hilo_outputs():
    from keep_b200 import ops

    x = torch.randn(300, 768, device=DEV) * 2 + 0.5
    w, b = torch.rand(768, device=DEV) + 0.5, torch.randn(768, device=DEV)
    y, y32 = ops.layernorm(x, w, b, 1e-12, want_f32=True, hilo=True)
    assert y.shape == (300, 1536)
    assert torch.equal(y[:, :768], y32.half()) and torch.equal(y[:, 768:], (y32 - y32.half().float()).half())
    B, S, H = 5, 24, 12
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).half().to(DEV)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(DEV)
    plain = ops.attention(qkv, B, S, H, key_mask=mask)
    hl = ops.attention(qkv, B, S, H, key_mask=mask, hilo=True)
    q, k, v = qkv.double().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    bias = torch.zeros(B, 1, 1, S, device=DEV, dtype=torch.double).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=0.125).transpose(1, 2).reshape(B * S, H * 64)
    assert torch.equal(hl[:, :H * 64], plain)
    e_hi, e_hl = _rel_l2(plain, ref), _rel_l2(hl[:, :H * 64].float() + hl[:, H * 64:].float(), ref)
    print(f"attention context: hi only {e_hi:.2e}, hi+lo {e_hl:.2e}")
    assert e_hl < e_hi  # what remains is the 16-bit rounding of P inside the kernel


@pytest.mark.parametrize("n", [1, 3, 150, 700, 1500])
def test_fused_fp32_head_and_pooler(n):
    """visual_head (LayerNorm -> Linear -> GELU(erf) -> Linear -> L2) and the BERT pooler (Linear -> tanh -> L2) as single
    fp32 kernels vs torch fp32 (keep_inference.py:42-46, 54-62): fp32 round-off only, every rows-per-CTA variant."""
    from keep_b200 import ops

    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, 1024, generator=g) * 3 + 1).to(DEV)
    lnw, lnb = (torch.rand(1024, generator=g) + 0.5).to(DEV), torch.randn(1024, generator=g).to(DEV)
    w0, b0 = (torch.randn(768, 1024, generator=g) / 32).to(DEV), (torch.randn(768, generator=g) * 0.1).to(DEV)
    w1, b1 = (torch.randn(768, 768, generator=g) / 27).to(DEV), (torch.randn(768, generator=g) * 0.1).to(DEV)
    got = ops.visual_head(x, lnw, lnb, 1e-6, w0, b0, w1, b1)
    h = F.layer_norm(x.double(), (1024,), lnw.double(), lnb.double(), 1e-6)
    ref = F.normalize(F.linear(F.gelu(F.linear(h, w0.double(), b0.double())), w1.double(), b1.double()), dim=-1)
    assert ((got.double() - ref).norm(dim=1) / ref.norm(dim=1)).max().item() < 5e-6
    assert torch.allclose(got.norm(dim=1), torch.ones(n, device=DEV), atol=1e-6)
    xp = torch.randn(n, 768, generator=g).to(DEV)
    gp = ops.pooler(xp, w1, b1)
    rp = F.normalize(torch.tanh(F.linear(xp.double(), w1.double(), b1.double())), dim=-1)
    assert ((gp.double() - rp).norm(dim=1) / rp.norm(dim=1)).max().item() < 5e-6


def test_empty_and_single_element_inputs():
    """Empty slides / single tiles / single prompts through the task kernels (the reference handles them in Python)."""
    from keep_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(77)
    cls = F.normalize(torch.randn(768, 2, device=DEV, generator=g), dim=0)
    keep, refined = ops.refine(torch.zeros(0, 2, dtype=torch.long, device=DEV), torch.zeros(0, 2, device=DEV), 224, True)
    assert keep.shape == (0,) and refined.shape == (0, 2)
    one = torch.softmax(torch.randn(1, 2, device=DEV, generator=g), 1)
    keep, refined = ops.refine(torch.tensor([[448, 224]], device=DEV), one, 224, True)
    assert keep.tolist() == [1] and torch.equal(refined, one)  # a lone tile is its own neighbourhood
    # negative coordinates are ordinary dict keys in the reference ((-1,-1) included); beyond 32 bits per component: refused
    import numpy as np
    from oracle import wsi_oracle as wo
    xy = torch.tensor([[-1, -1], [223, 223], [-1, 223], [-225, -225], [-1, -1], [2 ** 31 - 2, 5], [-2 ** 31, 5]])
    pr = torch.rand(7, 2)
    keep, refined = ops.refine(xy.to(DEV), pr.to(DEV), 224, True)
    exp = wo.refine_mean(pr.numpy(), xy.numpy(), 224, True)
    assert keep.tolist() == [1, 1, 1, 1, 0, 1, 1]
    got = refined.cpu().numpy()
    for i in (0, 1, 2, 3, 5, 6):
        assert np.array_equal(got[i], exp[tuple(xy[i].tolist())]), xy[i]
    with pytest.raises(ValueError):
        ops.refine(torch.tensor([[2 ** 31 - 1, 0]], device=DEV), one, 224, True)
    with pytest.raises(ValueError):
        ops.refine(torch.tensor([[0, -2 ** 31 - 1]], device=DEV), one, 224, True)
    x = torch.randn(1, 768, device=DEV, generator=g)
    lg, pr = ops.similarity(x, cls[:, :1].contiguous())        # one tile x one prompt
    assert lg.shape == (1, 1) and abs(pr.item() - 1.0) < 1e-6
    assert abs(lg.item() - (F.normalize(x, dim=-1) @ cls[:, :1]).item()) < 2e-4
    s = ops.prompt_scores(x, cls, 1, 2)                          # one tile, one two-class classifier
    l2 = (F.normalize(x, dim=-1) @ cls).flatten().sort(descending=True).values
    assert abs(s.item() - ((l2[0] - l2[1]) - (l2[0] + l2[1] - 1).abs()).item()) < 3e-4  # TF32 logits: ~3e-5 each


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_full_and_ragged_row_tiles_are_the_same_arithmetic(dtype):
    """The GEMMs run one kernel per (epilogue, 16-bit format, full / ragged row tiles): rows shared by a problem whose M is
    a multiple of the tile height (unguarded epilogue) and by the same problem with a ragged last tile (guarded epilogue)
    must come out bit for bit the same, on both main-loop variants (CTA pairs: 256-row tiles; single CTA: 128-row tiles)."""
    from keep_b200 import ops

    for (M_full, extra, N, K) in ((256 * 80, 37, 1024, 256), (128 * 3, 5, 256, 128)):
        g = torch.Generator(device="cpu").manual_seed(M_full + N)
        a = (torch.randn(M_full + extra, K, generator=g) * 0.5).to(dtype).to(DEV)
        w = (torch.randn(N, K, generator=g) * 0.05).to(dtype).to(DEV)
        bias = torch.randn(N, generator=g).to(DEV)
        gamma = torch.rand(N, generator=g).to(DEV)
        x0 = torch.randn(M_full + extra, N, generator=g).to(DEV)
        for epi in (ops.EPI_BIAS_HALF, ops.EPI_BIAS_GELU_HALF, ops.EPI_BIAS_F32):
            full = ops.gemm(a[:M_full], w, epi, bias=bias)
            rag = ops.gemm(a, w, epi, bias=bias)
            assert torch.equal(full, rag[:M_full]), (epi, M_full, dtype)
            ref = a.float() @ w.float().T + bias
            if epi == ops.EPI_BIAS_GELU_HALF:
                ref = F.gelu(ref)
            assert _rel(rag.float(), ref) < (2e-2 if dtype == torch.bfloat16 else 3e-3)
        xf, xr = x0[:M_full].clone(), x0.clone()
        ops.gemm(a[:M_full], w, ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=xf, out=xf)
        ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=xr, out=xr)
        assert torch.equal(xf, xr[:M_full])
        if N % 64 == 0:
            xf, xr = x0[:M_full].clone(), x0.clone()
            f16, fst = ops.gemm_resid_stats(a[:M_full], w, xf, bias=bias, gamma=gamma)
            r16, rst = ops.gemm_resid_stats(a, w, xr, bias=bias, gamma=gamma)
            assert torch.equal(xf, xr[:M_full]) and torch.equal(f16, r16[:M_full]) and torch.equal(fst, rst[:M_full])


@pytest.mark.parametrize("N,P,group", [(1, 32, 2), (7, 16, 0), (129, 32, 2), (5000, 32, 2), (10_000, 32, 2), (9_999, 64, 4),
                                       (19_003, 256, 4), (19_003, 128, 2), (40_000, 256, 16), (20_000, 144, 0), (30_001, 512, 2)])
def test_similarity_tile_shapes(N, P, group):
    """Every scheduling variant of the TF32 similarity kernel: rows spread over all SMs in tiles of fewer than 128 rows
    (small N), 128-row single-CTA tiles, CTA pairs on 256-row tiles with half a classifier block each (wide P, many rows),
    ragged last tiles, several column tiles (P > 256) - logits and grouped probabilities against fp32, with and without the
    logits output."""
    from keep_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(N * 7 + P)
    feats = (torch.randn(N, 768, generator=g) * 1.5).to(DEV)
    cls = F.normalize(torch.randn(768, P, generator=g), dim=0).to(DEV)
    ref = F.normalize(feats, dim=-1) @ cls
    gg = group or P
    ref_p = torch.softmax(ref.view(N, P // gg, gg) * 10, -1).view(N, P)
    lg, pr = ops.similarity(feats, cls, group=group, temp=10.0)
    assert (lg - ref).abs().max().item() < 2e-4 and (pr - ref_p).abs().max().item() < 1e-3
    assert (pr.view(N, P // gg, gg).sum(-1) - 1).abs().max().item() < 1e-5
    if 16 % gg == 0:
        none, pr2 = ops.similarity(feats, cls, group=group, temp=10.0, want_logits=False)
        assert none is None and torch.equal(pr2, pr)
