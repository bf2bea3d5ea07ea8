"""Torch-tensor front ends of the handle-free C-ABI entry points (keepb200_op_*, keepb200_similarity,
keepb200_prompt_scores, keepb200_refine).  Tensors must live on a CUDA device; nothing here computes on the
host — arguments are validated, output tensors are allocated, and the library is called on the current stream.
"""
from __future__ import annotations

import torch

from . import _lib

EPI_BIAS_HALF, EPI_BIAS_GELU_HALF, EPI_RESID_F32, EPI_BIAS_F32, EPI_PATCH_F32 = range(5)
EPI_BIAS_GELU_HILO = 8
SPLIT_NONE, SPLIT_W, SPLIT_AW = 0, 1, 2


def _on(t):
    """Make the tensor's device current for the call: the library sizes its grids and caches kernel attributes per
    CURRENT device, so a process that drives several GPUs must launch with the right one selected."""
    return torch.cuda.device(t.device)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.KeepB200Error("keep_b200 ops need CUDA tensors (there is no CPU path)")


def _is_bf16(t) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float16:
        return 0
    raise TypeError(f"16-bit operand expected, got {t.dtype}")


def gemm(a, w, epi, bias=None, gamma=None, resid=None, out=None, pos=None, patches=0):
    """out = epilogue(a[M,K] @ w[N,K].T); a/w fp16 or bf16 (a may have a row pitch > K)."""
    _need_cuda(a, w, bias, gamma, resid, out, pos)
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    bf = _is_bf16(a)
    assert _is_bf16(w) == bf
    if out is None:
        if epi in (EPI_BIAS_HALF, EPI_BIAS_GELU_HALF):
            out = torch.empty(M, N, dtype=a.dtype, device=a.device)
        elif epi == EPI_PATCH_F32:
            out = torch.zeros((M // patches) * (patches + 1), N, dtype=torch.float32, device=a.device)
        else:
            out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    L = _lib.lib()
    with _on(a):
        _lib.check(
            L.keepb200_op_gemm(
                a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, epi, bf,
                _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(resid), resid.stride(0) if resid is not None else 0,
                out.data_ptr(), out.stride(0), _lib.ptr(pos), patches, _lib.stream_ptr(a.device),
            ),
            "op_gemm",
        )
    return out


def gemm_resid_stats(a, w, x, bias=None, gamma=None):
    """x[M,N] += gamma * (a @ w.T + bias) in place; returns (x16, stats[M, N/64, 2]) of the new x."""
    _need_cuda(a, w, x, bias, gamma)
    M, K = a.shape
    N = w.shape[0]
    assert x.shape == (M, N) and x.dtype == torch.float32 and x.is_contiguous() and a.is_contiguous() and w.is_contiguous()
    x16 = torch.empty(M, N, dtype=a.dtype, device=a.device)
    stats = torch.zeros(M, N // 64, 2, dtype=torch.float32, device=a.device)
    L = _lib.lib()
    with _on(a):
        _lib.check(
            L.keepb200_op_gemm_resid_stats(a.data_ptr(), w.data_ptr(), M, N, K, _is_bf16(a), _lib.ptr(bias), _lib.ptr(gamma),
                                           x.data_ptr(), x16.data_ptr(), stats.data_ptr(), _lib.stream_ptr(a.device)),
            "op_gemm_resid_stats",
        )
    return x16, stats


def fold_ln(w32, lnw, lnb, bias, dtype=torch.float16):
    """LayerNorm folded into the following Linear: returns (W16, s, c)."""
    _need_cuda(w32, lnw, lnb, bias)
    N, K = w32.shape
    w16 = torch.empty(N, K, dtype=dtype, device=w32.device)
    s = torch.empty(N, dtype=torch.float32, device=w32.device)
    c = torch.empty(N, dtype=torch.float32, device=w32.device)
    L = _lib.lib()
    with _on(w32):
        _lib.check(
            L.keepb200_op_fold_ln(w32.data_ptr(), N, K, lnw.data_ptr(), lnb.data_ptr(), _lib.ptr(bias), w16.data_ptr(),
                                  1 if dtype == torch.bfloat16 else 0, s.data_ptr(), c.data_ptr(), _lib.stream_ptr(w32.device)),
            "op_fold_ln",
        )
    return w16, s, c


def gemm_ln(x16, wf, s, c, stats, eps, gelu=False):
    """act(Linear(LayerNorm(x))) from the 16-bit copy of x, its row statistics and the folded weight."""
    _need_cuda(x16, wf, s, c, stats)
    M, K = x16.shape
    N = wf.shape[0]
    out = torch.empty(M, N, dtype=x16.dtype, device=x16.device)
    L = _lib.lib()
    with _on(x16):
        _lib.check(
            L.keepb200_op_gemm_ln(x16.data_ptr(), wf.data_ptr(), M, N, K, 1 if gelu else 0, _is_bf16(x16), c.data_ptr(),
                                  s.data_ptr(), stats.data_ptr(), eps, out.data_ptr(), _lib.stream_ptr(x16.device)),
            "op_gemm_ln",
        )
    return out


def pos_resample(pos, g0, gh, gw):
    """pos_embed [1 + g0*g0, D] fp32 -> [1 + gh*gw, D] (timm resample_abs_pos_embed semantics)."""
    _need_cuda(pos)
    D = pos.shape[-1]
    pos = pos.reshape(1 + g0 * g0, D).contiguous()
    out = torch.empty(1 + gh * gw, D, dtype=torch.float32, device=pos.device)
    L = _lib.lib()
    with _on(pos):
        _lib.check(L.keepb200_op_pos_resample(pos.data_ptr(), g0, gh, gw, D, out.data_ptr(), _lib.stream_ptr(pos.device)),
                   "op_pos_resample")
    return out


def layernorm(x, w, b, eps, out_dtype=torch.float16, want_f32=False, rows=None, row_stride=None, hilo=False):
    """hilo=True: y16 is the [rows, 2D] hi|lo operand of a split GEMM (lo = 16-bit(y - hi))."""
    _need_cuda(x, w, b)
    D = w.numel()
    if rows is None:
        rows = x.numel() // D
        row_stride = D
    y16 = torch.empty(rows, 2 * D if hilo else D, dtype=out_dtype, device=x.device) if out_dtype is not None else None
    y32 = torch.empty(rows, D, dtype=torch.float32, device=x.device) if want_f32 else None
    L = _lib.lib()
    with _on(x):
        _lib.check(
            L.keepb200_op_layernorm(
                x.data_ptr(), row_stride, rows, D, w.data_ptr(), b.data_ptr(), eps, _lib.ptr(y16),
                1 if out_dtype == torch.bfloat16 else 0, _lib.ptr(y32), 2 * D if hilo else D, D if hilo else 0,
                _lib.stream_ptr(x.device),
            ),
            "op_layernorm",
        )
    return y16, y32


def attention(qkv, B, S, H, key_mask=None, scale=0.125, hilo=False):
    """qkv [B*S, 3*H*64] 16-bit -> context [B*S, H*64] (hilo=True: [B*S, 2*H*64] = [hi | rounding remainder])."""
    _need_cuda(qkv, key_mask)
    out = torch.empty(B * S, (2 if hilo else 1) * H * 64, dtype=qkv.dtype, device=qkv.device)
    L = _lib.lib()
    with _on(qkv):
        _lib.check(
            L.keepb200_op_attention(
                qkv.data_ptr(), out.data_ptr(), B, S, H, _is_bf16(qkv), _lib.ptr(key_mask),
                key_mask.stride(0) if key_mask is not None else 0, scale, out.stride(0), H * 64 if hilo else 0,
                _lib.stream_ptr(qkv.device),
            ),
            "op_attention",
        )
    return out


def cast_hilo(w32, dtype=torch.float16):
    """fp32 [rows, K] -> 16-bit [rows, 2K] = [hi | lo], lo = 16-bit(w - hi): the operand layout of the split GEMMs."""
    _need_cuda(w32)
    w32 = w32.contiguous().float()
    rows, K = w32.shape
    out = torch.empty(rows, 2 * K, dtype=dtype, device=w32.device)
    with _on(w32):
        _lib.check(_lib.lib().keepb200_op_cast_hilo(w32.data_ptr(), out.data_ptr(), rows, K, 1 if dtype == torch.bfloat16 else 0,
                                                    _lib.stream_ptr(w32.device)), "op_cast_hilo")
    return out


def gemm_split(a, w, K, epi, split, bias=None, resid=None, hilo_out=False):
    """Split-operand GEMM: a [M, K or 2K], w [N, 2K] hi|lo (ops.cast_hilo); epi as ops.gemm or EPI_BIAS_GELU_HILO."""
    _need_cuda(a, w, bias, resid)
    M, N = a.shape[0], w.shape[0]
    bf = _is_bf16(a)
    if epi in (EPI_BIAS_HALF, EPI_BIAS_GELU_HALF, EPI_BIAS_GELU_HILO):
        out = torch.empty(M, 2 * N if epi == EPI_BIAS_GELU_HILO else N, dtype=a.dtype, device=a.device)
    else:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with _on(a):
        _lib.check(
            _lib.lib().keepb200_op_gemm_split(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, epi, bf, split,
                                              _lib.ptr(bias), _lib.ptr(resid), resid.stride(0) if resid is not None else 0,
                                              out.data_ptr(), out.stride(0), N if epi == EPI_BIAS_GELU_HILO else 0,
                                              _lib.stream_ptr(a.device)),
            "op_gemm_split",
        )
    return out


def visual_head(x, lnw, lnb, eps, w0, b0, w1, b1):
    """normalize(w1 @ gelu(w0 @ LayerNorm(x) + b0) + b1) in one fp32 kernel; w0 [N0, D], w1 [N1, N0] (torch layout)."""
    _need_cuda(x, lnw, lnb, w0, b0, w1, b1)
    x = x.contiguous().float()
    n, D = x.shape
    out = torch.empty(n, w1.shape[0], dtype=torch.float32, device=x.device)
    w0t, w1t = w0.t().contiguous().float(), w1.t().contiguous().float()
    with _on(x):
        _lib.check(
            _lib.lib().keepb200_op_visual_head(x.data_ptr(), D, n, D, lnw.data_ptr(), lnb.data_ptr(), eps, w0t.data_ptr(),
                                               b0.data_ptr(), w0.shape[0], w1t.data_ptr(), b1.data_ptr(), w1.shape[0],
                                               out.data_ptr(), _lib.stream_ptr(x.device)),
            "op_visual_head",
        )
    return out


def pooler(x, w, b):
    """normalize(tanh(w @ x + b)) in one fp32 kernel; w [D, D] (torch layout)."""
    _need_cuda(x, w, b)
    x = x.contiguous().float()
    n, D = x.shape
    out = torch.empty(n, D, dtype=torch.float32, device=x.device)
    wt = w.t().contiguous().float()
    with _on(x):
        _lib.check(_lib.lib().keepb200_op_pooler(x.data_ptr(), D, n, D, wt.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                 _lib.stream_ptr(x.device)), "op_pooler")
    return out


def act_l2norm(x, act=0):
    _need_cuda(x)
    y = torch.empty_like(x)
    L = _lib.lib()
    with _on(x):
        _lib.check(L.keepb200_op_act_l2norm(x.data_ptr(), x.shape[0], x.shape[1], act, y.data_ptr(), _lib.stream_ptr(x.device)), "op_act_l2norm")
    return y


def similarity(feats, cls, group=0, temp=10.0, want_probs=True, tensor_cores=True, want_logits=True, out_logits=None,
               out_probs=None, workspace=None):
    """logits = normalize(feats) @ cls ; probs = softmax(temp*logits) per `group` columns (0 = all).
    tensor_cores=True: TF32 tcgen05 kernel (needs D % 32 == 0); False: fp32 FMA kernel (bit-closer to fp32).
    out_logits / out_probs: preallocated [N, P] fp32 row blocks to write into (e.g. slices of a slide-sized buffer);
    want_logits=False skips the logits (group must divide 16): the kernel is bound by its output bytes."""
    _need_cuda(feats, cls)
    feats = feats.contiguous().float()
    cls = cls.contiguous().float()
    N, D = feats.shape
    P = cls.shape[1]
    assert cls.shape[0] == D
    for o in (out_logits, out_probs):
        assert o is None or (o.shape == (N, P) and o.dtype == torch.float32 and o.is_contiguous() and o.device == feats.device)
    logits = out_logits if out_logits is not None else (torch.empty(N, P, dtype=torch.float32, device=feats.device) if want_logits else None)
    probs = out_probs if out_probs is not None else (torch.empty(N, P, dtype=torch.float32, device=feats.device) if want_probs else None)
    L = _lib.lib()
    ws_bytes = L.keepb200_similarity_workspace_bytes(D, P) if tensor_cores else 0
    ws = None
    if tensor_cores:
        ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=feats.device)
    with _on(feats):
        _lib.check(
            L.keepb200_similarity(feats.data_ptr(), N, D, cls.data_ptr(), P, group, temp, _lib.ptr(logits), _lib.ptr(probs),
                                  _lib.ptr(ws), ws_bytes, _lib.stream_ptr(feats.device)),
            "similarity",
        )
    return logits, probs


def prompt_scores(feats, cls, K, C, workspace_mb=256, fused=False):
    """scores[k] = mean_n(top1 - top2 - |top1 + top2 - 1|) of normalize(feats) @ cls[:, k*C:(k+1)*C] (deterministic).
    fused=True: the margin is reduced inside the similarity epilogue, the logits never reach memory (C in {2,4,8,16})."""
    _need_cuda(feats, cls)
    feats = feats.contiguous().float()
    cls = cls.contiguous().float()
    N, D = feats.shape
    assert cls.shape == (D, K * C)
    scores = torch.empty(K, dtype=torch.float32, device=feats.device)
    L = _lib.lib()
    full = L.keepb200_prompt_scores_workspace_bytes(N, D, K, C)     # classifier copy + all logits
    minimum = L.keepb200_prompt_scores_workspace_bytes(64, D, K, C)  # ... + 64 rows of logits (chunked)
    ws_bytes = max(minimum, min(full, (workspace_mb << 20) + minimum))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
    with _on(feats):
        _lib.check(
            L.keepb200_prompt_scores(feats.data_ptr(), N, D, cls.data_ptr(), K, C, 1 if fused else 0, scores.data_ptr(),
                                     ws.data_ptr(), ws_bytes, _lib.stream_ptr(feats.device)),
            "prompt_scores",
        )
    return scores


def refine(coords, probs, patch_size, overlap):
    """Device refine_seg: returns (keep uint8 [N], refined fp32 [N,C])."""
    _need_cuda(coords, probs)
    coords = coords.contiguous().long()
    probs = probs.contiguous().float()
    N, Cc = probs.shape
    if N > 0:
        # the hash key holds 32 bits per component (csrc/refine.cu::pack_xy): level-0 pixel coordinates are far inside, a
        # coordinate outside would be silently dropped, so it is refused (one 16-byte host read per slide)
        lo, hi = (int(v) for v in torch.aminmax(coords))
        if lo < -(2 ** 31) or hi > 2 ** 31 - 2:
            raise ValueError(f"refine: tile coordinates outside [-2^31, 2^31 - 2] (min {lo}, max {hi})")
    keep = torch.empty(N, dtype=torch.uint8, device=probs.device)
    refined = torch.empty(N, Cc, dtype=torch.float32, device=probs.device)
    L = _lib.lib()
    ws_bytes = L.keepb200_refine_workspace_bytes(N)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=probs.device)
    with _on(probs):
        _lib.check(
            L.keepb200_refine(coords.data_ptr(), probs.data_ptr(), N, Cc, patch_size, 1 if overlap else 0, keep.data_ptr(),
                              refined.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(probs.device)),
            "refine",
        )
    return keep, refined
