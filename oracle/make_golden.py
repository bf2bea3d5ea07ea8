"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container — TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # needs /root/reference (read-only); writes tests/golden/

The reference has no golden vectors of its own, so the pins are outputs of its own code:

  keep_tiny.npz / keep_full.npz
      the class body of /root/reference/quick_start/keep_inference.py (lines 1-76, exec'd verbatim; `timm` is
      served by oracle.keep_oracle's restatement because timm is not installed) is instantiated, loaded with
      seeded synthetic weights, and run on seeded inputs. Stored: outputs, a weight checksum, and (full) the
      example tile `quick_start/example.tif` after the reference transform (:88-93), rounded to fp16.
  vit_torchvision.npz
      the ViT restatement against the independent `torchvision.models.vit_l_16` (weights remapped, gamma = 1).
  transform.npz
      the reference's input transform (keep_inference.py:88-93: torchvision Resize(224, BICUBIC) + CenterCrop(224) on the
      PIL image, then ToTensor + Normalize) run with the real torchvision + Pillow on seeded uint8 tiles of several
      sizes and on quick_start/example.tif (whose raw pixels are stored too: /root/reference does not travel).
  wsi.npz
      outputs of /root/reference/WSI_evaluation/{utils,detection_utils,subtyping_utils,segment_utils}.py
      (imported as they lie, with empty stub modules for the absent h5py/openslide) on seeded inputs.

Inputs are regenerated from seeds by the tests (only outputs are stored), except the example tile.
"""
from __future__ import annotations

import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import keep_oracle as ko  # noqa: E402
from oracle.fake_tokenizer import FakeTokenizer  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------------------------------------
# the reference's own model class
# ---------------------------------------------------------------------------------------------------------
def reference_model_namespace():
    """exec lines 1-76 of quick_start/keep_inference.py (everything before the checkpoint loading)."""
    ko.install_timm_shim()
    src = open(os.path.join(REF, "quick_start", "keep_inference.py")).read().splitlines()
    end = next(i for i, l in enumerate(src) if l.startswith("model_path"))
    if "keep_inference_reference" in sys.modules:
        return sys.modules["keep_inference_reference"].__dict__
    mod = types.ModuleType("keep_inference_reference")  # transformers looks the class's module up in sys.modules
    mod.__file__ = os.path.join(REF, "quick_start", "keep_inference.py")
    sys.modules[mod.__name__] = mod
    exec(compile("\n".join(src[:end]), "keep_inference.py[1:%d]" % end, "exec"), mod.__dict__)
    return mod.__dict__


LAYER_TOKENS_IMAGE = [0, 1, 98, 196]  # CLS, first / middle / last patch
LAYER_TOKENS_TEXT = [0, 1, 4, 8]       # [CLS] and three attended positions (config 1 prompts have 9-12 tokens)


def weight_checksum(sd) -> np.ndarray:
    """Order-independent digest of a state-dict: float64 [sum, sum of squares, sum of |x|*index-hash]."""
    acc = np.zeros(3, dtype=np.float64)
    for k in sorted(sd):
        v = sd[k].double().flatten()
        if v.numel() == 0:
            continue
        acc[0] += v.sum().item()
        acc[1] += (v * v).sum().item()
        acc[2] += (v[:: max(1, v.numel() // 97)].abs().sum().item()) * ((hash_name(k) % 1000) + 1)
    return acc


def hash_name(s: str) -> int:
    import zlib

    return zlib.crc32(s.encode())


def full_inputs(example_tile_f16: torch.Tensor):
    g = torch.Generator().manual_seed(0)
    tiles = torch.stack([example_tile_f16.float(), torch.randn(3, 224, 224, generator=g)])
    text = ko.synthetic_text_inputs(3, seq_len=256, seed=0)
    # lengths 12 / 9 / 11 as BASELINE.md config 1
    g2 = torch.Generator().manual_seed(100)
    ids = torch.zeros(3, 256, dtype=torch.long)
    mask = torch.zeros(3, 256, dtype=torch.long)
    for i, L in enumerate((12, 9, 11)):
        ids[i, 0] = 2
        ids[i, 1:L - 1] = torch.randint(5, 30522, (L - 2,), generator=g2)
        ids[i, L - 1] = 3
        mask[i, :L] = 1
    text = {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask}
    return tiles, text


def tiny_inputs():
    g = torch.Generator().manual_seed(7)
    tiles = torch.randn(5, 3, 224, 224, generator=g)
    text = ko.synthetic_text_inputs(6, seq_len=64, vocab=1000, seed=7, min_len=3, max_len=40)
    return tiles, text


def make_model_goldens():
    ns = reference_model_namespace()
    RefConfig, RefModel = ns["KEEPConfig"], ns["KEEPModel"]

    # ---- tiny geometry: the reference class with the shim told to build a small tower ----
    import timm  # the shim

    orig = timm.create_model
    timm.create_model = lambda *a, **kw: orig(*a, **{**kw, **{k: v for k, v in ko.TINY_VISION_CONFIG.items()
                                                              if k not in ("img_size", "patch_size")}})
    ref_tiny = RefModel(RefConfig(text_config=ko.TINY_TEXT_CONFIG, projection_dim=128)).eval()
    timm.create_model = orig
    sd = ko.synthetic_state_dict(ref_tiny, seed=1)
    ref_tiny.load_state_dict(sd, strict=True)
    tiles, text = tiny_inputs()
    with torch.no_grad():
        out = ref_tiny(tiles, text)
        trunk = ref_tiny.visual(tiles)
    np.savez_compressed(
        os.path.join(OUT, "keep_tiny.npz"),
        vision_features=out["vision_features"].numpy(), text_features=out["text_features"].numpy(),
        trunk_cls=trunk.numpy(), checksum=weight_checksum(sd),
        similarity=(out["vision_features"] @ out["text_features"].T).numpy())
    print("tiny: keys", len(sd), "vision", tuple(out["vision_features"].shape), "text", tuple(out["text_features"].shape))

    # ---- full geometry (ViT-L/16 + BERT-base), BASELINE config 1 ----
    from PIL import Image
    from torchvision import transforms

    transform = transforms.Compose([  # keep_inference.py:88-93
        transforms.Resize(size=224, interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.CenterCrop(size=(224, 224)),
        transforms.ToTensor(),
        transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
    ])
    tile = transform(Image.open(os.path.join(REF, "quick_start", "example.tif")).convert("RGB"))
    tile16 = tile.to(torch.float16)
    ref_full = RefModel(RefConfig(text_config=ko.DEFAULT_TEXT_CONFIG, projection_dim=768)).eval()
    n_vit = sum(p.numel() for p in ref_full.visual.parameters())
    assert n_vit == 303_350_784, n_vit
    sd = ko.synthetic_state_dict(ref_full, seed=0)
    assert len(sd) == 546, len(sd)
    ref_full.load_state_dict(sd, strict=True)
    tiles, text = full_inputs(tile16)
    # per-layer activations of the reference class for config 1: the residual stream after every
    # ViT block / BERT layer, sampled at LAYER_TOKENS (the whole stream would be 38 MB)
    vis_layers, txt_layers = [], []
    hooks = [blk.register_forward_hook(lambda m, i, o: vis_layers.append(o[:, LAYER_TOKENS_IMAGE].clone()))
             for blk in ref_full.visual.blocks]
    hooks += [lay.register_forward_hook(lambda m, i, o: txt_layers.append((o[0] if isinstance(o, tuple) else o)[:, LAYER_TOKENS_TEXT].clone()))
              for lay in ref_full.text.encoder.layer]
    with torch.no_grad():
        out = ref_full(tiles, text)
    for h in hooks:
        h.remove()
    assert len(vis_layers) == 24 and len(txt_layers) == 12
    np.savez_compressed(os.path.join(OUT, "keep_full_layers.npz"),
                        vision_layers=torch.stack(vis_layers).numpy(), text_layers=torch.stack(txt_layers).numpy(),
                        image_tokens=np.asarray(LAYER_TOKENS_IMAGE), text_tokens=np.asarray(LAYER_TOKENS_TEXT))
    with torch.no_grad():
        trunk = ref_full.visual(tiles)
        pooled = ref_full.text(**text).pooler_output
    np.savez_compressed(
        os.path.join(OUT, "keep_full.npz"),
        example_tile_f16=tile16.numpy(), vision_features=out["vision_features"].numpy(),
        text_features=out["text_features"].numpy(), trunk_cls=trunk.numpy(), pooled=pooled.numpy(),
        similarity=(out["vision_features"] @ out["text_features"].T).numpy(), checksum=weight_checksum(sd))
    print("full: similarity\n", (out["vision_features"] @ out["text_features"].T).numpy())

    # ---- independent cross-check of the ViT restatement: torchvision vit_l_16 ----
    import torchvision

    tv = torchvision.models.vit_l_16(weights=None).eval()
    vit = ko.VisionTransformer(init_values=1.0).eval()
    vsd = ko.synthetic_state_dict(vit, seed=3)
    for k in vsd:
        if "gamma" in k:
            vsd[k] = torch.ones_like(vsd[k])
    vit.load_state_dict(vsd)
    m = {"conv_proj.weight": vsd["patch_embed.proj.weight"], "conv_proj.bias": vsd["patch_embed.proj.bias"],
         "class_token": vsd["cls_token"], "encoder.pos_embedding": vsd["pos_embed"],
         "encoder.ln.weight": vsd["norm.weight"], "encoder.ln.bias": vsd["norm.bias"]}
    for i in range(24):
        s, d = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        m[d + "ln_1.weight"], m[d + "ln_1.bias"] = vsd[s + "norm1.weight"], vsd[s + "norm1.bias"]
        m[d + "self_attention.in_proj_weight"], m[d + "self_attention.in_proj_bias"] = vsd[s + "attn.qkv.weight"], vsd[s + "attn.qkv.bias"]
        m[d + "self_attention.out_proj.weight"], m[d + "self_attention.out_proj.bias"] = vsd[s + "attn.proj.weight"], vsd[s + "attn.proj.bias"]
        m[d + "ln_2.weight"], m[d + "ln_2.bias"] = vsd[s + "norm2.weight"], vsd[s + "norm2.bias"]
        m[d + "mlp.0.weight"], m[d + "mlp.0.bias"] = vsd[s + "mlp.fc1.weight"], vsd[s + "mlp.fc1.bias"]
        m[d + "mlp.3.weight"], m[d + "mlp.3.bias"] = vsd[s + "mlp.fc2.weight"], vsd[s + "mlp.fc2.bias"]
    tv.heads = torch.nn.Identity()
    missing = tv.load_state_dict(m, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("heads")], missing
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        a, b = vit(x), tv(x)
    diff = (a - b).abs().max().item()
    print("torchvision cross-check max-abs diff:", diff)
    assert diff < 5e-5, diff
    np.savez_compressed(os.path.join(OUT, "vit_torchvision.npz"), torchvision_out=b.numpy(), max_abs_diff=np.float64(diff))


# ---------------------------------------------------------------------------------------------------------
# the reference's own WSI task functions
# ---------------------------------------------------------------------------------------------------------
def import_reference_wsi():
    for stub in ("h5py", "openslide"):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except ImportError:
                sys.modules[stub] = types.ModuleType(stub)
    sys.path.insert(0, os.path.join(REF, "WSI_evaluation"))
    import detection_utils  # noqa
    import segment_utils  # noqa
    import subtyping_utils  # noqa
    import utils  # noqa

    return utils, detection_utils, subtyping_utils, segment_utils


def wsi_inputs():
    g = torch.Generator().manual_seed(21)
    N = 1500
    feats = torch.randn(N, 768, generator=g) * 1.7
    # a 40x40 grid at stride 112 with duplicates and holes: exercises first-wins and missing neighbours
    gx = torch.randint(0, 40, (N,), generator=g) * 112
    gy = torch.randint(0, 40, (N,), generator=g) * 112
    coords = torch.stack([gx, gy], 1).numpy()
    cls2 = torch.nn.functional.normalize(torch.randn(768, 2, generator=g), dim=0)
    cls4 = torch.nn.functional.normalize(torch.randn(768, 4, generator=g), dim=0)
    bank = [torch.nn.functional.normalize(torch.randn(768, 2, generator=g), dim=0) for _ in range(37)]
    return feats, coords, cls2, cls4, bank


def make_wsi_goldens():
    utils, det, sub, seg = import_reference_wsi()
    feats, coords, cls2, cls4, bank = wsi_inputs()
    sink = io.StringIO()
    out = {}
    with redirect_stdout(sink):
        # detection: overlap False (script default) and True
        out["det_frac_no_overlap"] = np.float64(det.zero_shot_detection(cls2, feats, coords, patch_size=112, overlap=False))
        out["det_frac_overlap"] = np.float64(det.zero_shot_detection(cls2, feats, coords, patch_size=112, overlap=True))
        probs2 = torch.softmax((torch.nn.functional.normalize(feats, dim=-1) @ cls2) * 10, 1)
        preds, pr = det.refine_seg(probs2, coords, patch_size=112, overlap=True)
        out["det_keys"] = np.array(list(preds.keys()))
        out["det_preds"] = np.array(list(preds.values()), dtype=np.int64)
        out["det_probs"] = np.array(list(pr.values()), dtype=np.float64)
        # subtyping
        out["sub_label"] = np.int64(sub.zero_shot_subtyping(cls4, feats, coords, patch_size=112, overlap=True).item())
        probs4 = torch.softmax((torch.nn.functional.normalize(feats, dim=-1) @ cls4) * 10, 1)
        sp = sub.refine_seg(probs4, coords, patch_size=112, overlap=True)
        out["sub_preds"] = np.array(list(sp.values()), dtype=np.int64)
        # segmentation refine (eval needs openslide + a mask: out of scope)
        sg = seg.refine_seg(probs2, coords, patch_size=112, overlap=True)
        out["seg_probs"] = np.array(list(sg.values()), dtype=np.float64)
        # prompt screening
        merged = utils.zero_shot_prompt_select(bank, feats, topn=5, device="cpu")
        out["select_merged"] = merged.numpy()
        nf = torch.nn.functional.normalize(feats, dim=-1)
        out["select_scores"] = np.array([utils.rank_cls_score(nf @ c) for c in bank], dtype=np.float64)
    out["probs2_head"] = probs2[:64].numpy()

    # classifier construction through the reference code with the fake tokenizer + tiny reference model
    ns = reference_model_namespace()
    import timm

    orig = timm.create_model
    timm.create_model = lambda *a, **kw: orig(*a, **{**kw, **{k: v for k, v in ko.TINY_VISION_CONFIG.items()
                                                              if k not in ("img_size", "patch_size")}})
    tiny_text = dict(ko.TINY_TEXT_CONFIG, max_position_embeddings=256)
    model = ns["KEEPModel"](ns["KEEPConfig"](text_config=tiny_text, projection_dim=128)).eval()
    timm.create_model = orig
    model.load_state_dict(ko.synthetic_state_dict(model, seed=2))
    KEEP = {"model": model, "tokenizer": FakeTokenizer(1000)}
    prompts = {"classnames": {"Tumor": "tumor tissue", "Normal": "normal tissue", "CCRCC": "clear cell renal cell carcinoma"},
               "templates": "CLASSNAME."}
    c1 = utils.get_zeroshot_classifier(KEEP, {"Normal": 0, "Tumor": 1}, prompts, "cpu")
    c2 = utils.get_zeroshot_classifier(KEEP, {"CCRCC": 0, "Tumor": 1}, prompts, "cpu", add_normal=True)
    multi = {"classnames": prompts["classnames"], "templates": ["a photo of CLASSNAME.", "CLASSNAME, H&E."]}
    c3 = utils.get_zeroshot_classifier(KEEP, {"Normal": 0, "Tumor": 1}, multi, "cpu")
    out["classifier_basic"] = c1.numpy()
    out["classifier_add_normal"] = c2.numpy()
    out["classifier_multi_template"] = c3.numpy()
    np.savez_compressed(os.path.join(OUT, "wsi.npz"), **out)
    print("wsi: detection", out["det_frac_no_overlap"], out["det_frac_overlap"], "subtype", out["sub_label"],
          "kept tiles", len(out["det_preds"]))


TRANSFORM_SIZES = [(256, 256), (300, 260), (224, 298), (512, 384), (231, 227), (180, 200), (224, 224), (1024, 1024)]


def transform_inputs():
    """Seeded uint8 RGB tiles [H,W,3], smooth + noise so that the resampler sees both regimes."""
    rng = np.random.default_rng(4242)
    tiles = []
    for (H, W) in TRANSFORM_SIZES:
        yy, xx = np.mgrid[0:H, 0:W]
        base = 127 + 90 * np.sin(xx / 17.0)[..., None] * np.cos(yy / 11.0)[..., None] * np.asarray([1.0, 0.7, -0.8])
        tiles.append(np.clip(base + rng.normal(0, 40, (H, W, 3)), 0, 255).astype(np.uint8))
    return tiles


def make_transform_goldens():
    from PIL import Image
    from torchvision import transforms

    resize_crop = transforms.Compose([  # keep_inference.py:88-90
        transforms.Resize(size=224, interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.CenterCrop(size=(224, 224)),
    ])
    full = transforms.Compose([  # keep_inference.py:88-93
        resize_crop, transforms.ToTensor(),
        transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
    ])
    out = {}
    for i, t in enumerate(transform_inputs()):
        out[f"u8_{i}"] = np.asarray(resize_crop(Image.fromarray(t)))
    ex = Image.open(os.path.join(REF, "quick_start", "example.tif")).convert("RGB")
    out["example_raw"] = np.asarray(ex)
    out["example_u8"] = np.asarray(resize_crop(ex))
    out["example_f32"] = full(ex).numpy()
    import PIL
    import torchvision
    out["versions"] = np.asarray([PIL.__version__, torchvision.__version__])
    np.savez_compressed(os.path.join(OUT, "transform.npz"), **out)
    print("transform:", {k: v.shape for k, v in out.items()})


def make_seg_eval_goldens():
    """The reference's slide-level segmentation metrics (segment_utils.py:91-152) on a synthetic mask pyramid served by
    oracle/fake_openslide.py (openslide itself and the mask TIFFs are not available offline)."""
    from oracle import fake_openslide

    mask, probs = fake_openslide.synthetic_case()
    with fake_openslide.installed(mask):
        for m in ("segment_utils",):
            sys.modules.pop(m, None)  # re-import against the fake module
        if "h5py" not in sys.modules:
            try:
                __import__("h5py")
            except ImportError:
                sys.modules["h5py"] = types.ModuleType("h5py")
        sys.path.insert(0, os.path.join(REF, "WSI_evaluation"))
        import segment_utils as seg

        auc, thr = seg.eval_seg_auc(probs, "mask.tif", patch_size=224)
        dice = {t: seg.eval_seg_coarse(probs, "mask.tif", patch_size=224, thd=t) for t in (0.3, 0.5, float(thr))}
        empty = seg.eval_seg_coarse({k: 0.0 for k in probs}, "mask.tif", patch_size=224, thd=0.5)
    np.savez_compressed(os.path.join(OUT, "seg_eval.npz"), auc=np.float64(auc), thr=np.float64(thr),
                        dice_thd=np.array(list(dice.keys()), dtype=np.float64), dice=np.array(list(dice.values()), dtype=np.float64),
                        dice_no_prediction=np.float64(empty), n_tiles=np.int64(len(probs)))
    print("seg_eval: auc", auc, "thr", thr, "dice", dice, "no prediction", empty)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if len(sys.argv) > 1 and sys.argv[1] == "transform":
        make_transform_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "seg_eval":
        make_seg_eval_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "model":
        make_model_goldens()
        sys.exit(0)
    make_wsi_goldens()
    make_model_goldens()
    make_transform_goldens()
    make_seg_eval_goldens()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
