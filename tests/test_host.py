"""CPU: host-side mirror of the reference API (config, state-dict contract, registration, error behaviour)."""
import pytest
import torch

from keep_b200 import KEEPConfig, KEEPModel, KeepB200Error
from keep_b200.weights import state_dict_spec
from oracle import keep_oracle as ko
from tests import common


def test_spec_matches_reference_state_dict_full():
    spec = state_dict_spec(KEEPConfig(text_config=ko.DEFAULT_TEXT_CONFIG))
    assert len(spec) == 546
    from transformers import BertConfig, BertModel

    with torch.device("meta"):
        vit = ko.VisionTransformer()
        bert = BertModel(BertConfig(**ko.DEFAULT_TEXT_CONFIG))
    ref = {"visual." + k: tuple(v.shape) for k, v in vit.state_dict().items()}
    ref.update({"text." + k: tuple(v.shape) for k, v in bert.state_dict().items()})
    mine = {k: v for k, v in spec.items() if k.startswith(("visual.", "text."))}
    assert mine == ref


def test_automodel_registration_and_strict_loading():
    from transformers import AutoConfig, AutoModel

    cfg = KEEPConfig(text_config=ko.TINY_TEXT_CONFIG, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128)
    assert cfg.model_type == "keep"
    m = AutoModel.from_config(cfg)
    assert isinstance(m, KEEPModel)
    oracle, sd, _ = common.tiny_oracle(seed=1)
    assert set(m.state_dict().keys()) == set(oracle.state_dict().keys())
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.state_dict()["visual.blocks.1.attn.qkv.weight"], sd["visual.blocks.1.attn.qkv.weight"])
    assert float(m.logit_scale) == pytest.approx(3.2188758, abs=1e-6)  # log(1/0.04), keep_inference.py:52
    extra = dict(sd, **{"visual.extra": torch.zeros(1)})
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(extra, strict=True)
    wrong = dict(sd)
    wrong["visual.norm.weight"] = torch.zeros(7)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(wrong, strict=True)
    old_buffers = dict(sd, **{"text.embeddings.position_ids": torch.arange(64).unsqueeze(0)})
    m.load_state_dict(old_buffers, strict=True)  # transformers 4.34-era checkpoints carry this buffer
    assert AutoConfig.for_model("keep").__class__ is KEEPConfig


def test_save_and_reload_roundtrip(tmp_path):
    cfg = KEEPConfig(text_config=ko.TINY_TEXT_CONFIG, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128)
    m = KEEPModel(cfg)
    _, sd, _ = common.tiny_oracle(seed=5)
    m.load_state_dict(sd)
    torch.save(m.state_dict(), tmp_path / "pytorch_model.bin")
    cfg.save_pretrained(tmp_path)
    from transformers import AutoConfig, AutoModel

    cfg2 = AutoConfig.from_pretrained(str(tmp_path / "config.json"))   # keep_inference.py:80
    m2 = AutoModel.from_config(cfg2)                                     # :81
    m2.load_state_dict(torch.load(tmp_path / "pytorch_model.bin", map_location="cpu"), strict=True)  # :82-83
    m2.eval()
    assert torch.equal(m2.state_dict()["text.pooler.dense.weight"], sd["text.pooler.dense.weight"])


def test_from_pretrained_local_release_dir(tmp_path):
    """README.md:49-55 of the reference: AutoModel.from_pretrained(<repo>) - here on a local directory, through
    transformers' own loader (safetensors written by save_pretrained, and the release's pytorch_model.bin layout)."""
    from transformers import AutoModel

    cfg = KEEPConfig(text_config=ko.TINY_TEXT_CONFIG, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128)
    m = KEEPModel(cfg)
    _, sd, _ = common.tiny_oracle(seed=5)
    m.load_state_dict(sd)
    m.save_pretrained(tmp_path / "st")
    m2 = AutoModel.from_pretrained(str(tmp_path / "st"))
    assert type(m2) is KEEPModel and m2._dirty
    for k, v in sd.items():
        if k in m2.state_dict():
            assert torch.equal(m2.state_dict()[k], v), k
    (tmp_path / "bin").mkdir()
    cfg.save_pretrained(tmp_path / "bin")
    torch.save(m.state_dict(), tmp_path / "bin" / "pytorch_model.bin")
    m3 = KEEPModel.from_pretrained(str(tmp_path / "bin"))
    assert torch.equal(m3.state_dict()["visual.blocks.1.attn.qkv.weight"], sd["visual.blocks.1.attn.qkv.weight"])
    assert float(m3.logit_scale.detach()) == pytest.approx(3.2188758, abs=1e-6)


def test_no_cpu_fallback():
    m = KEEPModel(KEEPConfig(text_config=ko.TINY_TEXT_CONFIG, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128))
    with pytest.raises(KeepB200Error, match="no CPU fallback"):
        m.encode_image(torch.zeros(1, 3, 224, 224))
    with pytest.raises(KeepB200Error, match="no CPU fallback"):
        m.encode_text({"input_ids": torch.zeros(1, 8, dtype=torch.long)})
    from keep_b200 import ops

    with pytest.raises(KeepB200Error):
        ops.similarity(torch.zeros(4, 8), torch.zeros(8, 2))


def test_config_validation():
    with pytest.raises(ValueError, match="exact-erf GELU"):
        KEEPConfig(text_config=dict(ko.TINY_TEXT_CONFIG, hidden_act="relu")).text()
    cfg = KEEPConfig()  # reference default: vision_config=None, text_config=None, projection_dim=768
    assert cfg.projection_dim == 768 and cfg.vision()["width"] == 1024 and cfg.text()["hidden_size"] == 768


def test_product_path_never_imports_the_oracle():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "keep_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_slide_feature_files_round_trip(tmp_path):
    """keep_b200.io.load_slide: the reference's .h5 / .pt feature files (zeroshot_detection_WSI.py:29-31, utils.py:50-60)
    plus npz/npy; dtypes and shapes are normalised, mismatches are loud."""
    import numpy as np
    import pytest
    import torch

    from keep_b200 import io as kio

    g = np.random.default_rng(0)
    feats = g.standard_normal((37, 768)).astype(np.float32)
    coords = (g.integers(0, 90, (37, 2)) * 224).astype(np.int32)
    np.savez(tmp_path / "s.npz", features=feats, coords=coords)
    np.save(tmp_path / "s.npy", feats.astype(np.float64))
    torch.save(torch.from_numpy(feats), tmp_path / "bare.pt")
    torch.save({"features": torch.from_numpy(feats), "coords": torch.from_numpy(coords)}, tmp_path / "dict.pt")
    for name, has_coords in (("s.npz", True), ("s.npy", False), ("bare.pt", False), ("dict.pt", True)):
        f, c = kio.load_slide(str(tmp_path / name))
        assert f.dtype == torch.float32 and tuple(f.shape) == (37, 768) and np.allclose(f.numpy(), feats)
        assert (c is not None) == has_coords
        if has_coords:
            assert c.dtype == torch.int64 and np.array_equal(c.numpy(), coords)
    np.savez(tmp_path / "bad.npz", features=feats, coords=coords[:5])
    with pytest.raises(ValueError):
        kio.load_slide(str(tmp_path / "bad.npz"))
    with pytest.raises(ValueError):
        kio.load_slide(str(tmp_path / "slide.xyz"))
    try:
        import h5py
    except ImportError:
        with pytest.raises(ImportError):
            kio.load_slide(str(tmp_path / "s.h5"))
    else:
        with h5py.File(tmp_path / "s.h5", "w") as f:
            f["features"], f["coords"] = feats, coords
        f2, c2 = kio.load_slide(str(tmp_path / "s.h5"))
        assert np.allclose(f2.numpy(), feats) and np.array_equal(c2.numpy(), coords)
