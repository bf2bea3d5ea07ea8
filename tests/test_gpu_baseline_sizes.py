"""GPU: the hot path at BASELINE.json's full sizes. The fp32 CPU oracle cannot encode 10k-200k tiles in test time, so
these check size-independent properties of the domain: row independence / permutation equivariance,
unit norms, probabilities summing to one per classifier, linearity of the similarity in the classifier, exactness of
trimming, first-occurrence de-duplication, plus the oracle itself on random sub-samples and on everything that is
integer/index work (the refine walk at 200k tiles is compared with the reference dict walk in full)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import common

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def full_pair():
    oracle, sd = common.full_oracle(seed=0)
    return oracle, common.full_product(sd)


def _unit(x):
    return (x.norm(dim=1) - 1).abs().max().item()


def test_config2_detection_10k_tiles_x_32_prompts(full_pair):
    """zeroshot_detection_WSI: 10,000 tiles in batches of 1024 (the bench workload), 16 two-class classifiers."""
    from keep_b200 import ops

    oracle, prod = full_pair
    N, P = 10_000, 32
    g = torch.Generator(device=DEV).manual_seed(1234)
    tiles = torch.empty(N, 3, 224, 224, device=DEV)
    for b0 in range(0, N, 1024):
        tiles[b0:b0 + 1024].normal_(generator=g)
    # duplicates planted far apart: identical tiles must give identical embeddings wherever they sit in a batch/chunk
    tiles[9_999] = tiles[0]
    tiles[5_123] = tiles[1_024]
    feats = torch.cat([prod.encode_image(tiles[b0:b0 + 1024]) for b0 in range(0, N, 1024)])
    assert feats.shape == (N, 768) and torch.isfinite(feats).all()
    assert _unit(feats) < 1e-5
    assert torch.equal(feats[9_999], feats[0]) and torch.equal(feats[5_123], feats[1_024])
    # the oracle on a random sample of the 10k (fp32 CPU, a few seconds)
    idx = torch.tensor([0, 777, 1023, 1024, 4095, 9_998])
    with torch.no_grad():
        ref = oracle.encode_image(tiles[idx].cpu())
    rl, cos = common.row_metrics(feats[idx], ref)
    print(f"config 2, fast path, 6 of the 10k tiles vs the fp32 oracle: rel-L2 {rl:.2e}")
    assert rl <= 1.25e-3 and cos >= 0.99999, (rl, cos)  # FAST_REL_IMAGE of tests/test_gpu_model.py
    # a different batch split gives the same rows (row independence of every kernel on the path)
    again = prod.encode_image(tiles[700:1500])
    assert (again - feats[700:1500]).abs().max().item() < 1e-5
    # similarity + per-classifier softmax at full size
    text = common.ko.synthetic_text_inputs(P, seq_len=256, seed=3000)
    cls = prod.encode_text(common.to_device(text, DEV)).t().contiguous()
    logits, probs = ops.similarity(feats, cls, group=2, temp=10.0)
    ref_l = feats.double() @ cls.double()
    assert (logits.double() - ref_l).abs().max().item() <= 1e-3
    assert (probs.view(N, 16, 2).sum(-1) - 1).abs().max().item() < 1e-5
    assert (probs.argmax(1) == torch.softmax(ref_l.view(N, 16, 2) * 10, -1).view(N, P).argmax(1)).float().mean().item() >= 0.999


def test_config3_subtyping_50k_x_256_similarity_properties():
    """zeroshot_subtyping_WSI: 50,000 tile embeddings x 256 prompt columns (64 four-class classifiers)."""
    from keep_b200 import ops

    N, P, G = 50_000, 256, 4
    g = torch.Generator(device=DEV).manual_seed(1235)
    feats = torch.randn(N, 768, device=DEV, generator=g) * 3.0  # un-normalised on purpose: the kernel normalises
    cls = F.normalize(torch.randn(768, P, device=DEV, generator=g), dim=0)
    logits, probs = ops.similarity(feats, cls, group=G, temp=10.0)
    # (1) against fp64 on a sample
    idx = torch.randint(0, N, (512,), device=DEV, generator=g)
    ref = F.normalize(feats[idx].double(), dim=-1) @ cls.double()
    assert (logits[idx].double() - ref).abs().max().item() <= 1e-3
    assert (probs[idx].double() - torch.softmax(ref.view(-1, P // G, G) * 10, -1).view(-1, P)).abs().max().item() <= 5e-3
    # (2) cosine range, probabilities of every classifier sum to one
    assert logits.abs().max().item() <= 1 + 1e-4
    assert (probs.view(N, P // G, G).sum(-1) - 1).abs().max().item() < 1e-5
    # (3) scale invariance of the rows (F.normalize) and row-permutation equivariance, bit for bit
    l2, _ = ops.similarity(feats * 0.25, cls, group=G)
    assert (l2 - logits).abs().max().item() < 2e-4
    perm = torch.randperm(N, device=DEV, generator=g)
    l3, p3 = ops.similarity(feats[perm], cls, group=G)
    assert torch.equal(l3, logits[perm]) and torch.equal(p3, probs[perm])
    # (4) linearity in the classifier: sim(x, a + b) = sim(x, a) + sim(x, b) (TF32 operand rounding apart)
    c2 = F.normalize(torch.randn(768, P, device=DEV, generator=g), dim=0)
    la, _ = ops.similarity(feats[:8192], cls, group=G)
    lb, _ = ops.similarity(feats[:8192], c2, group=G)
    lab, _ = ops.similarity(feats[:8192], cls + c2, group=G)
    assert (lab - (la + lb)).abs().max().item() < 1e-3
    # (5) subtype call: argmax of the per-class tile fraction agrees with the fp64 statement on the sample
    cls4 = cls[:, :4].contiguous()
    _, p4 = ops.similarity(feats, cls4, group=0)
    frac = torch.bincount(p4.argmax(1), minlength=4).float() / N
    ref4 = torch.softmax((F.normalize(feats.double(), dim=-1) @ cls4.double()) * 10, 1)
    assert torch.equal(frac.argmax(), (torch.bincount(ref4.argmax(1), minlength=4).float() / N).argmax())


def test_config4_segmentation_200k_overlapping_tiles():
    """zeroshot_segmentation_WSI: 200,000 overlapping tiles (stride 112, patch 224) x 2 prompts, streamed in batches
    of 1024; similarity, then refine_seg over the whole slide compared with the reference's dict walk IN FULL."""
    from keep_b200 import ops
    from oracle import wsi_oracle as wo

    N, P, ps = 200_000, 2, 224
    g = torch.Generator(device=DEV).manual_seed(2000)
    feats = torch.randn(N, 768, device=DEV, generator=g)
    cls = F.normalize(torch.randn(768, P, device=DEV, generator=g), dim=0)
    probs = torch.empty(N, P, device=DEV)
    for b0 in range(0, N, 1024):  # streamed
        _, pr = ops.similarity(feats[b0:b0 + 1024], cls, group=0, temp=10.0)
        probs[b0:b0 + 1024] = pr
    whole = ops.similarity(feats, cls, group=0, temp=10.0)[1]
    assert torch.equal(whole, probs)  # streaming changes nothing, bit for bit
    assert (probs.sum(1) - 1).abs().max().item() < 1e-6
    ref = torch.softmax((F.normalize(feats.double(), dim=-1) @ cls.double()) * 10, 1)
    assert (probs.double() - ref).abs().max().item() <= 5e-3
    # slide grid 500 x 400 at stride 112 (= 200,000 positions), shuffled, with 1,000 duplicated coordinates
    ys, xs = torch.meshgrid(torch.arange(400), torch.arange(500), indexing="ij")
    coords = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1) * 112
    cg = torch.Generator().manual_seed(7)
    coords = coords[torch.randperm(N, generator=cg)]
    coords[torch.randint(0, N, (1000,), generator=cg)] = coords[torch.randint(0, N, (1000,), generator=cg)]
    keep, refined = ops.refine(coords.to(DEV), probs, ps, True)
    pn, cn = probs.cpu().numpy(), coords.numpy()
    first = wo._first_occurrence(cn)
    exp_keep = np.zeros(N, dtype=np.uint8)
    exp_keep[list(first.values())] = 1
    assert np.array_equal(keep.cpu().numpy(), exp_keep)
    exp = wo.refine_mean(pn, cn, ps, True)
    got = refined.cpu().numpy()
    order = np.fromiter(first.values(), dtype=np.int64)
    exp_rows = np.stack([exp[k] for k in first.keys()])
    assert np.array_equal(got[order], exp_rows)  # float32 neighbour means, bit for bit
    # idempotence of the de-duplication: refining the kept tiles again keeps all of them
    kept = torch.from_numpy(order)
    keep2, _ = ops.refine(coords[kept].to(DEV), probs[kept.to(DEV)], ps, True)
    assert int(keep2.sum()) == len(order)


def test_config5_prompt_bank_91632_prompts(full_pair):
    """encode_text over 11,454 disease names x 8 templates at seq_len 256: the shipped (trimmed) path on the whole
    bank, the padded computation and the fp32 oracle on sub-samples."""
    oracle, prod = full_pair
    P = 11_454 * 8
    text = common.ko.synthetic_text_inputs(P, seq_len=256, seed=3000)
    # plant duplicates: the same prompt at both ends of the bank
    for k in text:
        text[k][P - 1] = text[k][0]
    dev_text = common.to_device(text, DEV)
    # a bank of this size runs the one-pass GEMMs (text_precision "auto" -> "fast" above 8192 prompts per call); it is
    # pinned explicitly so that the sub-sample calls below compute the same thing
    old_precision = prod.config.text_precision
    prod.config.text_precision = "fast"
    try:
        _prompt_bank_checks(oracle, prod, text, dev_text, P)
    finally:
        prod.config.text_precision = old_precision


def _prompt_bank_checks(oracle, prod, text, dev_text, P):
    out = torch.cat([prod.encode_text({k: v[p0:p0 + 16384] for k, v in dev_text.items()}) for p0 in range(0, P, 16384)])
    assert out.shape == (P, 768) and torch.isfinite(out).all()
    assert _unit(out) < 1e-5
    assert torch.equal(out[P - 1], out[0])
    idx = torch.tensor([0, 1, 4097, 50_000, P - 2])
    with torch.no_grad():
        ref = oracle.encode_text({k: v[idx] for k, v in text.items()})
    rl, cos = common.row_metrics(out[idx], ref)
    assert rl <= 2e-3 and cos >= 0.99999, (rl, cos)  # FAST_REL_TEXT of tests/test_gpu_model.py
    # Trimming to the longest attended position changes nothing mathematically (masked keys weigh exp(-inf) = 0 and only
    # the [CLS] row is consumed). The padded S = 256 computation walks its keys in three 112-key blocks instead of packed
    # 32-key tiles, so its 16-bit attention probabilities are rounded against a different reference and fp32 sums run in
    # another order: the two one-pass (fp16 operand) results agree to that rounding (measured 6.5e-4; each is ~1.4e-3 from
    # the fp32 oracle).
    sub = {k: v[20_000:20_512] for k, v in dev_text.items()}
    old = prod.trim_text
    try:
        prod.trim_text = False
        padded = prod.encode_text(sub)
    finally:
        prod.trim_text = old
    rl_pad, _ = common.row_metrics(padded, out[20_000:20_512])
    print(f"config 5: padded S=256 vs trimmed, 512 prompts: max rel-L2 {rl_pad:.2e}")
    assert rl_pad <= 1e-3


@pytest.mark.parametrize("operand_dtype", ["float16", "bfloat16"])
def test_config3_config4_tile_tower_on_a_ragged_shard(full_pair, operand_dtype):
    """BASELINE configs 3 / 4 through the TILE TOWER (not random features): a 1,130-tile shard = two 512-tile workspace
    chunks + a ragged 106-tile one, fed as batches of 1024 + 106 exactly as bench.py's extras do (fp32 tiles for config 3,
    uint8 NHWC for config 4), 256 prompt columns / 2 prompt columns; a sample of the rows against the fp32 oracle."""
    from keep_b200 import distributed as kd
    from keep_b200 import ops

    oracle, prod16 = full_pair
    if operand_dtype == "float16":
        prod, gate = prod16, 1.25e-3
    else:
        prod, gate = common.full_product(common.full_oracle(seed=0)[1], operand_dtype="bfloat16"), 2e-2
    N = 1130
    g = torch.Generator(device=DEV).manual_seed(1235)
    tiles = torch.randn(N, 3, 224, 224, device=DEV, generator=g)
    feats = kd.encode_tiles_sharded(prod, N, lambda lo, hi: tiles[lo:hi], batch=1024, gather_dtype=torch.float16)
    assert feats.shape == (N, 768) and torch.isfinite(feats).all()
    idx = torch.tensor([0, 511, 512, 1023, 1024, 1129])  # chunk and batch boundaries, the ragged tail
    with torch.no_grad():
        ref = oracle.encode_image(tiles[idx].cpu())
    rl, cos = common.row_metrics(feats[idx], ref)
    print(f"ragged shard, {operand_dtype}: rel-L2 {rl:.2e} (embeddings gathered as fp16)")
    assert rl <= gate and cos >= (0.99999 if operand_dtype == "float16" else 0.9995)
    cls = F.normalize(torch.randn(768, 256, device=DEV, generator=g), dim=0)
    _, probs = ops.similarity(feats, cls, group=4, temp=10.0, want_logits=False)
    ref_p = torch.softmax((F.normalize(ref, dim=-1) @ cls.cpu()).view(len(idx), 64, 4) * 10, -1).view(len(idx), 256)
    assert (probs[idx].cpu() - ref_p).abs().max().item() <= (5e-3 if operand_dtype == "float16" else 5e-2)
    assert (probs.view(N, 64, 4).sum(-1) - 1).abs().max().item() < 1e-5
    if operand_dtype == "float16":  # config 4: the same shard as raw uint8 tiles with the normalisation fused into the patch gather
        u8 = torch.randint(0, 256, (N, 224, 224, 3), device=DEV, generator=g, dtype=torch.uint8)
        f8 = torch.cat([prod.encode_image(u8[b0:b0 + 1024]) for b0 in range(0, N, 1024)])
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        with torch.no_grad():
            ref8 = oracle.encode_image((u8[idx].cpu().permute(0, 3, 1, 2).float() / 255.0 - mean) / std)
        rl8, cos8 = common.row_metrics(f8[idx], ref8)
        print(f"ragged shard, uint8 tiles: rel-L2 {rl8:.2e}")
        assert rl8 <= gate and cos8 >= 0.99999
