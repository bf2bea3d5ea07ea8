"""CPU: the oracle (oracle/) against the golden vectors produced by running the reference itself
(oracle/make_golden.py) — the pin that makes the oracle trustworthy before it judges the CUDA path."""
import os

import numpy as np
import pytest
import torch

from oracle import keep_oracle as ko
from oracle import wsi_oracle as wo
from oracle.fake_tokenizer import FakeTokenizer
from tests import common


def test_tiny_model_matches_reference_class(golden_dir):
    g = common.load_golden(golden_dir, "keep_tiny.npz")
    m, sd, _ = common.tiny_oracle(seed=1)
    assert np.allclose(common.weight_checksum(sd), g["checksum"], rtol=1e-12), "seeded weights differ from the golden run"
    tiles, text = common.tiny_inputs()
    with torch.no_grad():
        out = m(tiles, text)
        trunk = m.visual(tiles)
    np.testing.assert_allclose(trunk.numpy(), g["trunk_cls"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["vision_features"].numpy(), g["vision_features"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["text_features"].numpy(), g["text_features"], rtol=0, atol=1e-6)
    assert torch.allclose(out["vision_features"].norm(dim=1), torch.ones(5), atol=1e-6)


@pytest.fixture(scope="module")
def full_model():
    return common.full_oracle(seed=0)


def test_full_model_geometry(full_model):
    m, sd = full_model
    assert len(sd) == 546
    assert sum(p.numel() for p in m.visual.parameters()) == 303_350_784       # = timm ViT-L/16 (UNI)
    assert m.visual.pos_embed.shape == (1, 197, 1024)
    assert tuple(m.visual_head[0].weight.shape) == (768, 1024)


def test_full_model_matches_reference_class(full_model, golden_dir):
    """BASELINE config 1: example.tif + randn tile, 3 prompts (lengths 12/9/11), fp32 on CPU."""
    g = common.load_golden(golden_dir, "keep_full.npz")
    m, sd = full_model
    assert np.allclose(common.weight_checksum(sd), g["checksum"], rtol=1e-12)
    tiles, text = common.full_inputs(torch.from_numpy(g["example_tile_f16"]))
    with torch.no_grad():
        out = m(tiles, text)
    np.testing.assert_allclose(out["vision_features"].numpy(), g["vision_features"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["text_features"].numpy(), g["text_features"], rtol=0, atol=2e-6)
    sim = (out["vision_features"] @ out["text_features"].T).numpy()
    np.testing.assert_allclose(sim, g["similarity"], rtol=0, atol=2e-6)


def test_vit_restatement_matches_torchvision(golden_dir):
    """Independent implementation check: torchvision vit_l_16 outputs (stored) vs the timm restatement."""
    g = common.load_golden(golden_dir, "vit_torchvision.npz")
    vit = ko.VisionTransformer(init_values=1.0).eval()
    vsd = ko.synthetic_state_dict(vit, seed=3)
    for k in vsd:
        if "gamma" in k:
            vsd[k] = torch.ones_like(vsd[k])
    vit.load_state_dict(vsd)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        y = vit(x)
    assert (y.numpy() - g["torchvision_out"]).__abs__().max() < 5e-5


def test_padding_invariance_of_cls_output():
    """[CLS] output does not depend on fully masked trailing positions — the property encode_text's
    s_eff trimming relies on."""
    m, _, _ = common.tiny_oracle(seed=4)
    text = ko.synthetic_text_inputs(4, seq_len=64, vocab=1000, seed=9, min_len=3, max_len=20)
    with torch.no_grad():
        full = m.encode_text(text)
        short = m.encode_text({k: v[:, :24] for k, v in text.items()})
    assert (full - short).abs().max() < 2e-6


# ---- WSI task logic ---------------------------------------------------------------------------------------
def test_wsi_refine_and_heads_match_reference(golden_dir):
    g = common.load_golden(golden_dir, "wsi.npz")
    feats, coords, cls2, cls4, bank = common.wsi_inputs()
    _, probs2 = wo.tile_probs(cls2, feats)
    np.testing.assert_allclose(probs2[:64].numpy(), g["probs2_head"], rtol=0, atol=1e-7)
    preds, pr = wo.refine_seg_detection(probs2.numpy(), coords, patch_size=112, overlap=True)
    assert list(preds.keys()) == g["det_keys"].tolist()
    assert list(preds.values()) == g["det_preds"].tolist()
    assert np.array_equal(np.array(list(pr.values())), g["det_probs"])                  # bit-exact
    assert wo.zero_shot_detection(cls2, feats, coords, patch_size=112, overlap=False) == g["det_frac_no_overlap"]
    assert wo.zero_shot_detection(cls2, feats, coords, patch_size=112, overlap=True) == g["det_frac_overlap"]
    _, probs4 = wo.tile_probs(cls4, feats)
    sp = wo.refine_seg_subtyping(probs4.numpy(), coords, patch_size=112, overlap=True)
    assert list(sp.values()) == g["sub_preds"].tolist()
    assert int(wo.zero_shot_subtyping(cls4, feats, coords, patch_size=112, overlap=True)) == int(g["sub_label"])
    sg = wo.refine_seg_segment(probs2.numpy(), coords, patch_size=112, overlap=True)
    assert np.array_equal(np.array(list(sg.values())), g["seg_probs"])


def test_wsi_prompt_screening_matches_reference(golden_dir):
    g = common.load_golden(golden_dir, "wsi.npz")
    feats, _, _, _, bank = common.wsi_inputs()
    merged, scores = wo.zero_shot_prompt_select(bank, feats, topn=5)
    np.testing.assert_allclose(np.array(scores), g["select_scores"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(merged.numpy(), g["select_merged"], rtol=0, atol=1e-7)


def test_wsi_classifier_construction_matches_reference(golden_dir):
    g = common.load_golden(golden_dir, "wsi.npz")
    m, _, _ = common.tiny_oracle(seed=2, max_pos=256)
    tok = FakeTokenizer(1000)
    prompts = {"classnames": {"Tumor": "tumor tissue", "Normal": "normal tissue", "CCRCC": "clear cell renal cell carcinoma"},
               "templates": "CLASSNAME."}
    c1 = wo.get_zeroshot_classifier(m, tok, {"Normal": 0, "Tumor": 1}, prompts)
    c2 = wo.get_zeroshot_classifier(m, tok, {"CCRCC": 0, "Tumor": 1}, prompts, add_normal=True)
    multi = {"classnames": prompts["classnames"], "templates": ["a photo of CLASSNAME.", "CLASSNAME, H&E."]}
    c3 = wo.get_zeroshot_classifier(m, tok, {"Normal": 0, "Tumor": 1}, multi)
    np.testing.assert_allclose(c1.numpy(), g["classifier_basic"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(c2.numpy(), g["classifier_add_normal"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(c3.numpy(), g["classifier_multi_template"], rtol=0, atol=1e-6)
    assert c2.shape == (128, 3)


# ---- input transform (keep_inference.py:88-93): the oracle against the real torchvision + Pillow pipeline ----------
def test_transform_oracle_matches_torchvision_pil_goldens(golden_dir):
    from oracle import transform_oracle as to
    from oracle.make_golden import transform_inputs

    g = np.load(f"{golden_dir}/transform.npz")
    for i, tile in enumerate(transform_inputs()):
        assert np.array_equal(to.resize_center_crop(tile), g[f"u8_{i}"]), (i, tile.shape)  # bit-exact, every size
    assert np.array_equal(to.resize_center_crop(g["example_raw"]), g["example_u8"])
    assert np.abs(to.to_tensor_normalize(g["example_u8"]) - g["example_f32"]).max() <= 1e-6


def test_transform_oracle_matches_live_pil_when_available():
    PIL = pytest.importorskip("PIL.Image")
    tvt = pytest.importorskip("torchvision.transforms")
    from oracle import transform_oracle as to

    tf = tvt.Compose([tvt.Resize(224, interpolation=tvt.InterpolationMode.BICUBIC), tvt.CenterCrop((224, 224))])
    rng = np.random.default_rng(9)
    for (H, W) in [(257, 311), (640, 480), (225, 224), (150, 333)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(to.resize_center_crop(img), np.asarray(tf(PIL.fromarray(img)))), (H, W)


def test_seg_eval_matches_the_reference_metrics(golden_dir):
    """keep_b200.seg_eval (AUROC + Youden threshold, coarse Dice at the level nearest 16x) against the reference's own
    eval_seg_auc / eval_seg_coarse (WSI_evaluation/segment_utils.py:91-152) run on the same synthetic mask pyramid
    (tests/golden/seg_eval.npz, made by `python -m oracle.make_golden seg_eval` with oracle/fake_openslide.py standing in
    for openslide, which is not available offline)."""
    from keep_b200 import seg_eval
    from oracle import fake_openslide

    g = common.load_golden(golden_dir, "seg_eval.npz")
    mask, probs = fake_openslide.synthetic_case()
    assert len(probs) == int(g["n_tiles"])
    with fake_openslide.installed(mask):
        auc, thr = seg_eval.eval_seg_auc(probs, "mask.tif", patch_size=224)
        assert auc == pytest.approx(float(g["auc"]), abs=1e-12) and thr == pytest.approx(float(g["thr"]), abs=1e-12)
        for t, d in zip(g["dice_thd"], g["dice"]):
            assert seg_eval.eval_seg_coarse(probs, "mask.tif", patch_size=224, thd=float(t)) == pytest.approx(float(d), abs=1e-12)
        # nothing predicted: Dice 0 (the reference's "1 when both are empty" branch needs an empty mask as well)
        assert seg_eval.eval_seg_coarse({k: 0.0 for k in probs}, "mask.tif", thd=0.5) == float(g["dice_no_prediction"])
        assert seg_eval.eval_seg_coarse({}, "mask.tif") == 0.0
    with fake_openslide.installed(np.zeros_like(mask)):
        assert seg_eval.eval_seg_coarse({k: 0.0 for k in probs}, "mask.tif", thd=0.5) == 1
