"""One command for the day the released checkpoint is at hand (it cannot be fetched in the build sandbox):

    python tools/check_real_weights.py /path/to/KEEP_release [--image quick_start/example.tif] [--oracle-only]

`/path/to/KEEP_release` is the directory the reference's quick start loads (`keep_inference.py:79-87`): `config.json`,
`pytorch_model.bin` and the tokenizer files. The script repeats the reference demo (`keep_inference.py:95-104`: one
image, three prompts) twice - through the fp32 CPU oracle (the reference semantics) and through keep_b200 on cuda:0 -
and prints the per-embedding rel-L2 / cosine and the similarity rows, i.e. the parity gate of SURVEY.md section 8d on the
REAL weights instead of the seeded synthetic ones the test-suite has to use. Test/diagnostic tool: not product path.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PROMPTS = ["an H&E image of breast invasive carcinoma.", "an H&E image of normal tissue.",
           "an H&E image of lung adenocarcinoma."]  # keep_inference.py:96


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model_dir")
    ap.add_argument("--image", default=None, help="an RGB image file (default: a seeded random tile)")
    ap.add_argument("--oracle-only", action="store_true", help="CPU part only (no GPU needed)")
    a = ap.parse_args()

    from oracle import keep_oracle as ko
    from oracle import transform_oracle as to

    cfg = json.load(open(os.path.join(a.model_dir, "config.json")))
    text_cfg = cfg.get("text_config") or ko.DEFAULT_TEXT_CONFIG
    proj = cfg.get("projection_dim", 768)
    sd = torch.load(os.path.join(a.model_dir, "pytorch_model.bin"), map_location="cpu")
    oracle = ko.KEEPModel(text_cfg, proj, ko.DEFAULT_VISION_CONFIG).eval()
    missing, unexpected = oracle.load_state_dict(sd, strict=False)
    unexpected = [k for k in unexpected if not k.endswith(("position_ids", "token_type_ids"))]
    if missing or unexpected:
        raise SystemExit(f"state-dict mismatch: missing {missing[:4]} unexpected {unexpected[:4]}")

    if a.image:
        import numpy as np
        from PIL import Image
        raw = np.asarray(Image.open(a.image).convert("RGB"))
        tile = torch.from_numpy(to.to_tensor_normalize(to.resize_center_crop(raw)))[None]  # keep_inference.py:88-93
    else:
        tile = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    try:
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(a.model_dir, do_lower_case=True, local_files_only=True)  # :87
        text = dict(tok(PROMPTS, max_length=256, padding="max_length", truncation=True, return_tensors="pt"))  # :99
    except Exception as e:  # tokenizer files absent: same ids convention as the test-suite
        print(f"(tokenizer not loadable: {e}; using synthetic token ids)")
        text = ko.synthetic_text_inputs(3, seq_len=256, seed=0)

    with torch.no_grad():
        ref_i, ref_t = oracle.encode_image(tile), oracle.encode_text(text)
    print("oracle (fp32 CPU) similarity:", (ref_i @ ref_t.T).flatten().tolist())
    if a.oracle_only:
        return

    from keep_b200 import KEEPConfig, KEEPModel
    model = KEEPModel(KEEPConfig(text_config=text_cfg, projection_dim=proj))
    model.load_state_dict(sd, strict=True)  # the reference's strict load (keep_inference.py:83)
    model = model.to("cuda:0").eval()
    img = model.encode_image(tile.to("cuda:0")).cpu()
    txt = model.encode_text({k: v.to("cuda:0") for k, v in text.items()}).cpu()
    print("keep_b200 (B200) similarity:  ", (img @ txt.T).flatten().tolist())
    for name, got, ref in (("image", img, ref_i), ("text", txt, ref_t)):
        rel = ((got.double() - ref.double()).norm(dim=1) / ref.double().norm(dim=1)).max().item()
        cos = torch.nn.functional.cosine_similarity(got.double(), ref.double(), dim=1).min().item()
        print(f"{name}: max rel-L2 {rel:.3e}, min cosine {cos:.7f}   (gate: 2e-3 / 0.99999)")
    print("similarity max-abs difference:", ((img @ txt.T) - (ref_i @ ref_t.T)).abs().max().item(), "(gate 1e-3)")


if __name__ == "__main__":
    main()
