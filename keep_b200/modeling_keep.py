"""KEEPModel — host-side mirror of the reference model class (quick_start/keep_inference.py:25-76).

Same surface as the reference: `KEEPModel(config)`, `AutoModel.from_config(config)`,
`load_state_dict(state_dict, strict=True)` with the reference's 546 keys, `.to(device)`, `.eval()`,
`encode_image(Tensor[B,3,224,224]) -> Tensor[B,768]`, `encode_text(Mapping) -> Tensor[P,768]`,
`forward(image_inputs, text_inputs) -> {"vision_features", "text_features"}`; outputs are fp32, unit L2 norm,
on the input device.

Underneath there is no PyTorch math: the parameters are held only so that the usual checkpoint plumbing
works, and are uploaded once into `libkeep_b200.so` (include/keep_b200.h), whose hand-written sm_100a kernels
do all the arithmetic.  PyTorch supplies device memory, the CUDA stream and (elsewhere) torch.distributed.
There is deliberately no CPU fallback: calling encode_* with the model or inputs on the CPU raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Mapping

import torch
import torch.nn as nn
from transformers import AutoConfig, AutoModel, PreTrainedModel

from . import _lib
from .configuration_keep import KEEPConfig
from .weights import IGNORED_BUFFERS, state_dict_spec

TILES_F32_NCHW, TILES_U8_NHWC = 0, 1
OP_ENCODE_IMAGE, OP_ENCODE_TEXT = 0, 1


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    """Register `param` under root at the dotted path, creating bare container modules on the way."""
    *path, leaf = dotted.split(".")
    mod = root
    for part in path:
        if part not in mod._modules:
            mod.add_module(part, nn.Module())
        mod = mod._modules[part]
    mod.register_parameter(leaf, param)


class KEEPModel(PreTrainedModel):
    config_class = KEEPConfig
    base_model_prefix = ""
    _no_split_modules: list = []
    # image tiles per pass through the tower; bounds the activation workspace (~4.4 MB per tile)
    image_chunk = int(__import__("os").environ.get("KEEPB200_IMAGE_CHUNK", "512"))
    text_chunk_tokens = 1 << 18
    trim_text = True  # skip positions that no row attends (bit-identical result); False = always the padded length

    def __init__(self, config: KEEPConfig):
        super().__init__(config)
        self._spec = state_dict_spec(config)
        for name, shape in self._spec.items():
            if name == "logit_scale":  # keep_inference.py:52 — dead at inference but part of the state-dict
                p = nn.Parameter(torch.ones([]) * math.log(1 / 0.04), requires_grad=False)
            else:
                p = nn.Parameter(torch.zeros(shape, dtype=torch.float32), requires_grad=False)
            _attach(self, name, p)
        self._handle = None
        self._handle_device = None
        self._dirty = True
        self._ws = None
        # transformers' own bookkeeping (tied-weight tables etc.): `from_pretrained(<local release dir>)` needs it on
        # transformers >= 5; `_init_weights` is a no-op here, weights always come from a state-dict
        self.post_init()

    # ---- checkpoint plumbing -------------------------------------------------------------------------
    def _init_weights(self, module):  # weights always come from a state-dict
        pass

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {k: v for k, v in state_dict.items() if k not in IGNORED_BUFFERS}
        out = super().load_state_dict(sd, strict=strict, assign=assign)
        self._dirty = True
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._dirty = True
        return out

    def float(self):
        return self

    def half(self):
        raise _lib.KeepB200Error("precision is selected with config.operand_dtype; parameters stay fp32")

    bfloat16 = half

    # ---- library handle ---------------------------------------------------------------------------------
    def _c_config(self) -> _lib.KeepB200Config:
        v, t = self.config.vision(), self.config.text()
        od = {"float16": 0, "fp16": 0, "bfloat16": 1, "bf16": 1}.get(str(self.config.operand_dtype))
        if od is None:
            raise ValueError(f"operand_dtype must be 'float16' or 'bfloat16', got {self.config.operand_dtype!r}")
        c = _lib.KeepB200Config()
        c.struct_size = C.sizeof(_lib.KeepB200Config)
        c.img_size, c.patch_size = v["img_size"], v["patch_size"]
        c.vit_width, c.vit_depth, c.vit_heads, c.vit_mlp = v["width"], v["depth"], v["heads"], v["mlp"]
        c.vit_ln_eps = v["ln_eps"]
        c.proj_dim = self.config.projection_dim
        c.vocab_size, c.hidden = t["vocab_size"], t["hidden_size"]
        c.layers, c.heads = t["num_hidden_layers"], t["num_attention_heads"]
        c.intermediate, c.max_pos = t["intermediate_size"], t["max_position_embeddings"]
        c.type_vocab, c.bert_ln_eps = t["type_vocab_size"], t["layer_norm_eps"]
        c.operand_dtype = od
        return c

    def _release(self):
        if self._handle is not None:
            _lib.lib().keepb200_destroy(self._handle)
            self._handle = None
            self._handle_device = None
        self._ws = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _device(self) -> torch.device:
        return self.logit_scale.device

    def _sync(self) -> None:
        """(Re)create the device handle and upload the parameters if anything changed."""
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.KeepB200Error(
                f"KEEPModel is on {dev}; keep_b200 computes only on a CUDA sm_100a device (no CPU fallback): "
                "call model.to('cuda')")
        if self._handle is not None and not self._dirty and self._handle_device == dev:
            return
        L = _lib.lib()
        self._release()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        with torch.cuda.device(idx):
            h = C.c_void_p()
            cfg = self._c_config()
            _lib.check(L.keepb200_create(C.byref(cfg), idx, C.byref(h)), "create")
            self._handle, self._handle_device = h, dev
            stream = _lib.stream_ptr(dev)
            for name, p in self.state_dict().items():
                t = p.detach().to(device=dev, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(L.keepb200_load_weight(h, name.encode(), t.data_ptr(), shape, t.dim(), stream),
                           f"load_weight({name})")
            torch.cuda.current_stream(dev).synchronize()  # sources may be temporaries
            _lib.check(L.keepb200_finalize(h), "finalize")
        self._dirty = False

    def _workspace(self, nbytes: int, dev) -> torch.Tensor:
        # + 1024: _aligned() may move the base up by as much as 1023 bytes
        if self._ws is None or self._ws.numel() < nbytes + 1024 or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        return self._ws

    @staticmethod
    def _aligned(ws: torch.Tensor) -> int:
        return (ws.data_ptr() + 1023) // 1024 * 1024

    # ---- the reference API -----------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_image(self, image_inputs: torch.Tensor) -> torch.Tensor:
        """normalize(visual_head(visual(x)))  — keep_inference.py:54-58.

        `image_inputs`: float [B,3,224,224] ImageNet-normalised (reference contract), or uint8 [B,224,224,3]
        raw RGB (ToTensor+Normalize are then fused into the patch gather)."""
        self._sync()
        dev = self._device()
        x = image_inputs
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("encode_image expects a 4-D tensor")
        if x.device != dev:
            raise _lib.KeepB200Error(f"image_inputs on {x.device} but the model is on {dev}")
        if x.dtype == torch.uint8:
            if x.shape[3] != 3:
                raise ValueError(f"uint8 tiles must be [B,H,W,3] (NHWC), got {tuple(x.shape)}")
            layout, (H, W) = TILES_U8_NHWC, x.shape[1:3]
            x = x.contiguous()
        else:
            if x.shape[1] != 3:
                raise ValueError(f"float tiles must be [B,3,H,W] (NCHW), got {tuple(x.shape)}")
            layout, (H, W) = TILES_F32_NCHW, x.shape[2:4]
            x = x.to(torch.float32).contiguous()
        # the reference ViT is built with dynamic_img_size=True (keep_inference.py:39): any multiple of the patch size
        # is accepted and pos_embed is resampled to the new grid; 224x224 is the native grid. The attention kernels
        # serve up to 512 tokens, larger tiles are refused (never silently approximated).
        if H % 16 or W % 16 or H <= 0 or W <= 0:
            raise ValueError(f"tile height and width must be multiples of the 16-pixel patch, got {H}x{W}")
        if (H // 16) * (W // 16) + 1 > 512:
            raise NotImplementedError(f"encode_image supports tiles of up to 511 patches, got {H}x{W}")
        B = x.shape[0]
        out = torch.empty(B, self.config.projection_dim, dtype=torch.float32, device=dev)
        if B == 0:
            return out
        precision = _lib.PRECISION.get(str(getattr(self.config, "image_precision", "auto")))
        if precision is None:
            raise ValueError(f"image_precision must be 'auto', 'high', 'balanced' or 'fast', got {self.config.image_precision!r}")
        L = _lib.lib()
        with torch.cuda.device(dev):
            high = L.keepb200_image_precision_is_high(precision, B)
            need = L.keepb200_workspace_bytes_hw(self._handle, min(B, self.image_chunk), H, W, high)
            ws = self._workspace(need, dev)
            _lib.check(
                L.keepb200_encode_image_hw(self._handle, x.data_ptr(), layout, B, H, W, precision, out.data_ptr(),
                                           self._aligned(ws), need, _lib.stream_ptr(dev)),
                "encode_image")
        return out

    @torch.no_grad()
    def encode_text(self, text_inputs: Mapping) -> torch.Tensor:
        """normalize(BertModel(**text_inputs).pooler_output)  — keep_inference.py:60-62."""
        self._sync()
        dev = self._device()
        ids = text_inputs["input_ids"]
        if ids.device != dev:
            raise _lib.KeepB200Error(f"text_inputs on {ids.device} but the model is on {dev}")
        unknown = set(text_inputs.keys()) - {"input_ids", "token_type_ids", "attention_mask"}
        if unknown:
            raise TypeError(f"encode_text: unsupported BertModel arguments {sorted(unknown)}")
        ids = ids.to(torch.long).contiguous()
        P, S = ids.shape
        tt = text_inputs.get("token_type_ids")
        mask = text_inputs.get("attention_mask")
        tt = tt.to(device=dev, dtype=torch.long).contiguous() if tt is not None else None
        tcfg = self.config.text()
        hidden = tcfg["hidden_size"]
        out = torch.empty(P, hidden, dtype=torch.float32, device=dev)
        if P == 0:
            return out
        precision = _lib.PRECISION.get(str(getattr(self.config, "text_precision", "auto")))
        if precision is None:
            raise ValueError(f"text_precision must be 'auto', 'high', 'balanced' or 'fast', got {self.config.text_precision!r}")
        s_eff = S
        if mask is not None:
            mask = mask.to(device=dev, dtype=torch.long).contiguous()
        # One device->host read for all the facts the host needs. Out-of-range ids raise as nn.Embedding would (the
        # kernel clamps only to stay memory-safe). A row that attends to nothing is refused: BertModel would add
        # finfo.min to every score of that row and average V over all positions, pads included - no tokenizer
        # produces such a row, and silently returning something else is not an option here.
        facts = [ids.min(), ids.max(), tt.min() if tt is not None else ids.new_zeros(()),
                 tt.max() if tt is not None else ids.new_zeros(())]
        if mask is not None:
            attended = mask != 0
            facts += [attended.any(dim=1).all().to(torch.long),
                      (attended.any(dim=0).to(torch.long) * torch.arange(1, S + 1, device=dev)).max()]
        facts = torch.stack(facts).tolist()
        if facts[0] < 0 or facts[1] >= tcfg["vocab_size"]:
            raise IndexError(f"encode_text: input_ids outside [0, {tcfg['vocab_size']}) (min {facts[0]}, max {facts[1]})")
        if facts[2] < 0 or facts[3] >= tcfg["type_vocab_size"]:
            raise IndexError(f"encode_text: token_type_ids outside [0, {tcfg['type_vocab_size']})")
        if mask is not None:
            if not facts[4]:
                raise ValueError("encode_text: an attention_mask row has no attended position")
            if self.trim_text:
                # positions past the last attended key in EVERY row contribute exactly zero to the [CLS] output
                s_eff = max(int(facts[5]), 1)
        L = _lib.lib()
        with torch.cuda.device(dev):
            chunk = max(1, min(P, self.text_chunk_tokens // s_eff))
            need = L.keepb200_workspace_bytes(self._handle, OP_ENCODE_TEXT, chunk, s_eff)
            ws = self._workspace(need, dev)
            _lib.check(
                L.keepb200_encode_text(self._handle, ids.data_ptr(), _lib.ptr(tt), _lib.ptr(mask), P, S, s_eff, precision,
                                       out.data_ptr(), self._aligned(ws), need, _lib.stream_ptr(dev)),
                "encode_text")
        return out

    # ---- test / analysis hooks (include/keep_b200.h "debug") -----------------------------------------------------
    def debug_set_ln_fuse(self, mode: int) -> None:
        """LayerNorm placement in the ViT blocks (0 stand-alone, 1 default, 2 norm2 folded as well)."""
        self._sync()
        _lib.check(_lib.lib().keepb200_debug_set_ln_fuse(self._handle, int(mode)), "debug_set_ln_fuse")

    @torch.no_grad()
    def debug_layer_outputs(self, image_inputs=None, text_inputs=None):
        """Residual stream after every block of ONE tower for a small batch: list of [B, T, D] tensors (the last entry
        holds the CLS rows only, [B, D]) - what the per-layer parity table compares with hooks on the oracle."""
        self._sync()
        dev = self._device()
        L = _lib.lib()
        if (image_inputs is None) == (text_inputs is None):
            raise ValueError("give exactly one of image_inputs / text_inputs")
        if image_inputs is not None:
            v = self.config.vision()
            B = image_inputs.shape[0]
            T = (image_inputs.shape[-2] // 16) * (image_inputs.shape[-1] // 16) + 1
            depth, D = v["depth"], v["width"]
            if B > self.image_chunk:
                raise ValueError("debug_layer_outputs: the batch must fit one workspace chunk")
        else:
            t = self.config.text()
            depth, D = t["num_hidden_layers"], t["hidden_size"]
            B = text_inputs["input_ids"].shape[0]
            m = text_inputs.get("attention_mask")
            T = text_inputs["input_ids"].shape[1] if (m is None or not self.trim_text) else \
                max(int(((m != 0).any(dim=0).long() * torch.arange(1, m.shape[1] + 1, device=m.device)).max()), 1)
        buf = torch.zeros(depth, B * T * D, dtype=torch.float32, device=dev)
        _lib.check(L.keepb200_debug_layer_dump(self._handle, buf.data_ptr(), buf.numel() * 4), "debug_layer_dump")
        try:
            final = self.encode_image(image_inputs) if image_inputs is not None else self.encode_text(text_inputs)
            torch.cuda.synchronize(dev)
        finally:
            _lib.check(L.keepb200_debug_layer_dump(self._handle, None, 0), "debug_layer_dump")
        outs = [buf[i].view(B, T, D) for i in range(depth - 1)] + [buf[depth - 1, :B * D].view(B, D)]
        return outs, final

    def forward(self, image_inputs, text_inputs):
        """keep_inference.py:65-73."""
        return {"vision_features": self.encode_image(image_inputs), "text_features": self.encode_text(text_inputs)}


def register() -> None:
    """AutoConfig / AutoModel registration, as keep_inference.py:75-76."""
    try:
        AutoConfig.register("keep", KEEPConfig)
    except ValueError:
        pass
    try:
        AutoModel.register(KEEPConfig, KEEPModel)
    except ValueError:
        pass


register()
