"""GPU: encode_image / encode_text / forward through the C ABI against the CPU oracle (fp32) and the golden
vectors produced by the reference class (tests/golden, oracle/make_golden.py).

Tolerance (SURVEY.md §8d parity gate): fp16 operands with fp32 accumulation/residual/LayerNorm/softmax:
per-embedding rel-L2 <= 2e-3 and cosine >= 0.99999; similarity matrix max-abs <= 1e-3. bf16 operands are
reported against a looser 2e-2 (bf16 has 3 fewer mantissa bits; torch's own bf16 autocast lands at ~1e-2)."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP16_REL, FP16_COS, SIM_ABS = 2e-3, 0.99999, 1e-3


@pytest.fixture(scope="module")
def tiny_pair():
    oracle, sd, text_cfg = common.tiny_oracle(seed=1)
    return oracle, common.tiny_product(sd, text_cfg), sd, text_cfg


@pytest.fixture(scope="module")
def full_pair():
    oracle, sd = common.full_oracle(seed=0)
    return oracle, common.full_product(sd), sd


def test_library_weight_table_matches_python_spec(tiny_pair):
    from keep_b200 import _lib

    _, prod, _, _ = tiny_pair
    prod._sync()
    L = _lib.lib()
    n = L.keepb200_num_weights(prod._handle)
    names = {L.keepb200_weight_name(prod._handle, i).decode() for i in range(n)}
    assert names == set(prod.state_dict().keys())


def test_tiny_model_vs_oracle_and_golden(tiny_pair, golden_dir):
    oracle, prod, _, _ = tiny_pair
    g = common.load_golden(golden_dir, "keep_tiny.npz")
    tiles, text = common.tiny_inputs()
    out = prod(tiles.to(DEV), common.to_device(text, DEV))
    with torch.no_grad():
        ref = oracle(tiles, text)
    for key in ("vision_features", "text_features"):
        assert out[key].dtype == torch.float32 and out[key].device.type == "cuda"
        rl, cos = common.row_metrics(out[key], ref[key])
        assert rl <= FP16_REL and cos >= FP16_COS, (key, rl, cos)
        rl_g, _ = common.row_metrics(out[key], torch.from_numpy(g[key]))
        assert rl_g <= FP16_REL, (key, rl_g)
        assert torch.allclose(out[key].norm(dim=1), torch.ones(out[key].shape[0], device=DEV), atol=1e-5)
    sim = (out["vision_features"] @ out["text_features"].T).cpu().numpy()
    assert np.abs(sim - g["similarity"]).max() <= SIM_ABS


def test_full_model_config1_vs_golden(full_pair, golden_dir):
    """BASELINE config 1 (quick_start): example.tif + randn tile x 3 prompts, ViT-L/16 + BERT-base."""
    oracle, prod, _ = full_pair
    g = common.load_golden(golden_dir, "keep_full.npz")
    tiles, text = common.full_inputs(torch.from_numpy(g["example_tile_f16"]))
    img = prod.encode_image(tiles.to(DEV))
    txt = prod.encode_text(common.to_device(text, DEV))
    rl_i, cos_i = common.row_metrics(img, torch.from_numpy(g["vision_features"]))
    rl_t, cos_t = common.row_metrics(txt, torch.from_numpy(g["text_features"]))
    print(f"config1 fp16: image rel-L2 {rl_i:.2e} cos {cos_i:.7f}; text rel-L2 {rl_t:.2e} cos {cos_t:.7f}")
    assert rl_i <= FP16_REL and cos_i >= FP16_COS
    assert rl_t <= FP16_REL and cos_t >= FP16_COS
    sim = (img @ txt.T).cpu().numpy()
    assert np.abs(sim - g["similarity"]).max() <= SIM_ABS


def test_full_model_batch_vs_oracle(full_pair):
    """A ragged batch (not a multiple of any tile size) through the full towers vs the fp32 CPU oracle."""
    oracle, prod, _ = full_pair
    g = torch.Generator().manual_seed(99)
    tiles = torch.randn(7, 3, 224, 224, generator=g)
    from oracle import keep_oracle as ko

    text = ko.synthetic_text_inputs(9, seq_len=256, seed=5)
    with torch.no_grad():
        ref_i = oracle.encode_image(tiles)
        ref_t = oracle.encode_text(text)
    img = prod.encode_image(tiles.to(DEV))
    txt = prod.encode_text(common.to_device(text, DEV))
    rl_i, cos_i = common.row_metrics(img, ref_i)
    rl_t, cos_t = common.row_metrics(txt, ref_t)
    assert rl_i <= FP16_REL and cos_i >= FP16_COS, (rl_i, cos_i)
    assert rl_t <= FP16_REL and cos_t >= FP16_COS, (rl_t, cos_t)


def test_fused_layernorm_paths_match_standalone_layernorm(full_pair):
    """The image path folds norm1 (KEEPB200_LN_FUSE=1, default) or norm1 and norm2 (=2) into the following GEMMs
    (EPI_LN_*); =0 runs the stand-alone LayerNorm kernel in every block. All three must sit inside the parity gate
    and agree with each other."""
    import os

    oracle, prod, _ = full_pair
    g = torch.Generator().manual_seed(7)
    tiles = torch.randn(5, 3, 224, 224, generator=g)
    with torch.no_grad():
        ref = oracle.encode_image(tiles)
    outs = {}
    try:
        for mode in ("0", "1", "2"):
            os.environ["KEEPB200_LN_FUSE"] = mode
            outs[mode] = prod.encode_image(tiles.to(DEV)).clone()
    finally:
        del os.environ["KEEPB200_LN_FUSE"]
    for mode, out in outs.items():
        rl, cos = common.row_metrics(out, ref)
        rl0, _ = common.row_metrics(out, outs["0"])
        print(f"KEEPB200_LN_FUSE={mode}: rel-L2 vs fp32 oracle {rl:.2e} (cos {cos:.7f}); vs stand-alone LN {rl0:.2e}")
        assert rl <= FP16_REL and cos >= FP16_COS, (mode, rl, cos)
        assert rl0 <= FP16_REL
    assert not torch.equal(outs["0"], outs["1"]) and not torch.equal(outs["1"], outs["2"])  # the toggle switches paths
    assert torch.equal(outs["1"], prod.encode_image(tiles.to(DEV)))  # default = mode 1, and deterministic


def test_batch_invariance_and_chunking(tiny_pair):
    """Same tile alone, inside a batch, and across workspace chunks gives the same embedding."""
    _, prod, _, _ = tiny_pair
    g = torch.Generator().manual_seed(3)
    tiles = torch.randn(37, 3, 224, 224, generator=g).to(DEV)
    whole = prod.encode_image(tiles)
    old = prod.image_chunk
    try:
        prod.image_chunk = 8  # 37 tiles -> 5 chunks, the last one ragged
        prod._ws = None
        chunked = prod.encode_image(tiles)
    finally:
        prod.image_chunk = old
        prod._ws = None
    single = prod.encode_image(tiles[11:12])
    assert (whole - chunked).abs().max().item() < 1e-5
    assert (whole[11:12] - single).abs().max().item() < 1e-5


def test_text_trimming_matches_padded_computation(tiny_pair):
    """s_eff trimming (positions masked in every row are skipped) equals the padded computation."""
    from keep_b200 import _lib
    from oracle import keep_oracle as ko

    oracle, prod, _, _ = tiny_pair
    text = ko.synthetic_text_inputs(6, seq_len=64, vocab=1000, seed=12, min_len=3, max_len=17)
    trimmed = prod.encode_text(common.to_device(text, DEV))
    nomask_safe = dict(text)
    nomask_safe["attention_mask"] = text["attention_mask"].clone()
    nomask_safe["attention_mask"][0, -1] = 1  # forces s_eff = S; one extra attended PAD key in row 0 only
    padded = prod.encode_text(common.to_device(nomask_safe, DEV))
    assert (trimmed[1:] - padded[1:]).abs().max().item() < 2e-4
    with torch.no_grad():
        ref = oracle.encode_text(text)
    rl, cos = common.row_metrics(trimmed, ref)
    assert rl <= FP16_REL and cos >= FP16_COS


def test_text_optional_inputs(tiny_pair):
    """token_type_ids / attention_mask may be omitted (zeros / ones), as with BertModel."""
    from oracle import keep_oracle as ko

    oracle, prod, _, _ = tiny_pair
    text = ko.synthetic_text_inputs(3, seq_len=24, vocab=1000, seed=2, min_len=24, max_len=24)
    only_ids = {"input_ids": text["input_ids"]}
    with torch.no_grad():
        ref = oracle.encode_text(only_ids)
    got = prod.encode_text(common.to_device(only_ids, DEV))
    rl, _ = common.row_metrics(got, ref)
    assert rl <= FP16_REL


def test_uint8_nhwc_tiles_match_float_path(tiny_pair):
    """Fused ToTensor+Normalize on uint8 NHWC tiles == the reference transform followed by the float path."""
    _, prod, _, _ = tiny_pair
    g = torch.Generator().manual_seed(8)
    u8 = torch.randint(0, 256, (6, 224, 224, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    f32 = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std  # ToTensor + Normalize (keep_inference.py:91-92)
    a = prod.encode_image(u8.to(DEV))
    b = prod.encode_image(f32.to(DEV))
    assert (a - b).abs().max().item() < 2e-4


@pytest.mark.parametrize("H,W", [(256, 256), (160, 160), (224, 160), (112, 352)])
def test_dynamic_img_size_matches_oracle(full_pair, H, W):
    """dynamic_img_size=True (keep_inference.py:39): other multiples of 16 run with pos_embed resampled (bicubic,
    antialias) to the new grid. 257 tokens -> mma.sync attention, 101 / 141 / 155 tokens -> tcgen05 attention."""
    oracle, prod, _ = full_pair
    g = torch.Generator().manual_seed(H * 1000 + W)
    tiles = torch.randn(3, 3, H, W, generator=g)
    with torch.no_grad():
        ref = oracle.encode_image(tiles)
    out = prod.encode_image(tiles.to(DEV))
    rl, cos = common.row_metrics(out, ref)
    print(f"{H}x{W}: rel-L2 {rl:.2e} cos {cos:.7f}")
    assert rl <= FP16_REL and cos >= FP16_COS, (H, W, rl, cos)


def test_bf16_operands_reported(golden_dir):
    oracle, sd, text_cfg = common.tiny_oracle(seed=1)
    prod = common.tiny_product(sd, text_cfg, operand_dtype="bfloat16")
    tiles, text = common.tiny_inputs()
    out = prod(tiles.to(DEV), common.to_device(text, DEV))
    with torch.no_grad():
        ref = oracle(tiles, text)
    for key in ("vision_features", "text_features"):
        rl, cos = common.row_metrics(out[key], ref[key])
        print(f"bf16 {key}: rel-L2 {rl:.2e} cos {cos:.6f}")
        assert rl <= 2e-2 and cos >= 0.9995


def test_errors_are_loud(tiny_pair):
    from keep_b200 import KeepB200Error

    _, prod, sd, text_cfg = tiny_pair
    with pytest.raises(NotImplementedError):
        prod.encode_image(torch.zeros(1, 3, 368, 368, device=DEV))  # 530 tokens > 512: refused, not approximated
    with pytest.raises(ValueError):
        prod.encode_image(torch.zeros(1, 3, 230, 224, device=DEV))  # not a multiple of the patch size
    with pytest.raises(KeepB200Error):
        prod.encode_image(torch.zeros(1, 3, 224, 224))  # CPU input: no silent fallback
    assert prod.encode_image(torch.zeros(0, 3, 224, 224, device=DEV)).shape == (0, 128)
    ids = torch.full((2, 8), 5, dtype=torch.long, device=DEV)
    with pytest.raises(ValueError, match="no attended position"):
        prod.encode_text({"input_ids": ids, "attention_mask": torch.tensor([[1, 1, 0, 0, 0, 0, 0, 0], [0] * 8], device=DEV)})
    bad = dict(sd)
    bad.pop("visual.norm.weight")
    from keep_b200 import KEEPConfig, KEEPModel
    from oracle import keep_oracle as ko

    m = KEEPModel(KEEPConfig(text_config=text_cfg, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128))
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict(bad, strict=True)


def test_raw_uint8_tiles_through_transform_and_tower(full_pair, golden_dir):
    """The reference pipeline end to end on raw pixels: transform(PIL image) -> encode_image (keep_inference.py:88-101),
    here preprocess (device, bit-exact) -> encode_image(uint8) with ToTensor+Normalize fused into the patch gather."""
    from keep_b200.transform import preprocess

    oracle, prod, _ = full_pair
    g = common.load_golden(golden_dir, "transform.npz")
    raw = torch.from_numpy(g["example_raw"])                       # quick_start/example.tif, 224 x 298 RGB
    with torch.no_grad():
        ref = oracle.encode_image(torch.from_numpy(g["example_f32"])[None])  # the reference transform's own output
    out = prod.encode_image(preprocess(raw[None].to(DEV)))
    rl, cos = common.row_metrics(out, ref)
    print(f"example.tif raw pixels -> embedding: rel-L2 {rl:.2e} cos {cos:.7f}")
    assert rl <= FP16_REL and cos >= FP16_COS


def test_extreme_shapes_match_oracle(full_pair):
    """Edges of the accepted input space: one single-token prompt, prompts at BERT's 512-position limit (no padding at
    all), a tile at the 512-token limit of the attention kernels (352x352 -> 485 tokens), empty batches."""
    from oracle import keep_oracle as ko

    oracle, prod, _ = full_pair
    one = {"input_ids": torch.tensor([[2]]), "token_type_ids": torch.tensor([[0]]), "attention_mask": torch.tensor([[1]])}
    long = ko.synthetic_text_inputs(2, seq_len=512, seed=11, min_len=512, max_len=512)
    tile = torch.randn(1, 3, 352, 352, generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        ref_one, ref_long, ref_tile = oracle.encode_text(one), oracle.encode_text(long), oracle.encode_image(tile)
    for got, ref, what in ((prod.encode_text(common.to_device(one, DEV)), ref_one, "1-token prompt"),
                           (prod.encode_text(common.to_device(long, DEV)), ref_long, "512-token prompts"),
                           (prod.encode_image(tile.to(DEV)), ref_tile, "352x352 tile")):
        rl, cos = common.row_metrics(got, ref)
        print(f"{what}: rel-L2 {rl:.2e} cos {cos:.7f}")
        assert rl <= FP16_REL and cos >= FP16_COS, (what, rl, cos)
    empty = {k: v[:0] for k, v in common.to_device(long, DEV).items()}
    assert prod.encode_text(empty).shape == (0, 768)
    assert prod.encode_image(torch.zeros(0, 3, 224, 224, device=DEV)).shape == (0, 768)
