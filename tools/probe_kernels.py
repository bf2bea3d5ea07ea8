"""First-contact probe for the CUDA kernels (run on a B200 via gpurun). Unlike the pytest suite it never
stops at the first failure and prints enough structure about a mismatch to diagnose a wrong descriptor or
layout from one run.  Output: stdout + gpurun_out/probe_*.npy for failing cases."""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import ops  # noqa: E402

OUT = "gpurun_out"
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
RESULTS = []


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    rel = err.max().item() / denom
    ok = bool(torch.isfinite(got).all().item()) and rel <= tol
    print(f"[{'ok' if ok else 'FAIL'}] {name}: max_abs={err.max().item():.4e} rel_to_max={rel:.3e} (tol {tol:.1e})", flush=True)
    RESULTS.append((name, ok))
    if not ok:
        np.save(f"{OUT}/probe_{name.replace(' ', '_').replace('/', '_')}_got.npy", got.cpu().numpy()[:512, :512])
        np.save(f"{OUT}/probe_{name.replace(' ', '_').replace('/', '_')}_ref.npy", ref.cpu().numpy()[:512, :512])
        bad = (err > tol * denom)
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:32].tolist()
        print(f"      bad fraction {bad.float().mean().item():.4f}; first bad rows {rows}; first bad cols {cols}")
        print(f"      got[0,:8]={got[0, :8].tolist()}\n      ref[0,:8]={ref[0, :8].tolist()}")
    return ok


def run(name, fn):
    try:
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        print(f"      ({name}: {time.time() - t0:.2f}s)", flush=True)
    except Exception as e:  # keep going: later probes may still be informative
        print(f"[EXC] {name}: {e}")
        traceback.print_exc()
        RESULTS.append((name, False))


def gemm_case(M, N, K, epi, dtype=torch.float16, ints=False, tag=""):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    if ints:
        a = torch.randint(-2, 3, (M, K), generator=g).to(dtype)
        w = torch.randint(-2, 3, (N, K), generator=g).to(dtype)
    else:
        a = (torch.randn(M, K, generator=g) * 0.5).to(dtype)
        w = (torch.randn(N, K, generator=g) * 0.05).to(dtype)
    bias = torch.randn(N, generator=g) * 0.1
    gamma = torch.rand(N, generator=g) + 0.5
    a, w, bias, gamma = a.to(dev), w.to(dev), bias.to(dev), gamma.to(dev)
    acc = a.float() @ w.float().T
    name = f"gemm{tag} M{M} N{N} K{K} epi{epi} {str(dtype)[6:]}"
    if epi == ops.EPI_BIAS_HALF:
        out = ops.gemm(a, w, epi, bias=bias)
        report(name, out, acc + bias, 2e-3 if dtype == torch.float16 else 1e-2)
    elif epi == ops.EPI_BIAS_GELU_HALF:
        out = ops.gemm(a, w, epi, bias=bias)
        report(name, out, torch.nn.functional.gelu(acc + bias), 2e-3 if dtype == torch.float16 else 1e-2)
    elif epi == ops.EPI_RESID_F32:
        resid = torch.randn(M, N, generator=g).to(dev)
        ref = resid + gamma * (acc + bias)
        out = resid.clone()
        ops.gemm(a, w, epi, bias=bias, gamma=gamma, resid=out, out=out)  # in place, as the model uses it
        report(name, out, ref, 1e-5)
    elif epi == ops.EPI_BIAS_F32:
        out = ops.gemm(a, w, epi, bias=bias)
        report(name, out, acc + bias, 1e-5)
    elif epi == ops.EPI_PATCH_F32:
        P = 196
        assert M % P == 0
        pos = torch.randn(P + 1, N, generator=g).to(dev)
        out = ops.gemm(a, w, epi, bias=bias, pos=pos, patches=P)
        ref = torch.zeros(M // P, P + 1, N, device=dev)
        ref[:, 1:] = (acc + bias).view(M // P, P, N) + pos[1:]
        report(name, out, ref.view(-1, N), 1e-5)


def probe_gemm():
    # exact integer products first: any error here is a layout/descriptor bug, not rounding
    run("gemm-int-1tile", lambda: gemm_case(128, 128, 64, ops.EPI_BIAS_F32, ints=True, tag="-int"))
    run("gemm-int-k256", lambda: gemm_case(128, 128, 256, ops.EPI_BIAS_F32, ints=True, tag="-int"))
    run("gemm-int-2x2", lambda: gemm_case(256, 256, 128, ops.EPI_BIAS_F32, ints=True, tag="-int"))
    run("gemm-tail", lambda: gemm_case(197 * 3, 768, 1024, ops.EPI_BIAS_F32))
    run("gemm-wide", lambda: gemm_case(128 * 160, 1024, 1024, ops.EPI_BIAS_F32))  # 256-wide tile path, persistent loop
    run("gemm-wide-qkv", lambda: gemm_case(197 * 64, 3072, 1024, ops.EPI_BIAS_HALF))
    run("gemm-wide-gelu", lambda: gemm_case(197 * 64, 4096, 1024, ops.EPI_BIAS_GELU_HALF))
    run("gemm-wide-fc2", lambda: gemm_case(197 * 64, 1024, 4096, ops.EPI_RESID_F32))
    run("gemm-patch", lambda: gemm_case(196 * 8, 1024, 768, ops.EPI_PATCH_F32))
    run("gemm-bf16", lambda: gemm_case(197 * 8, 3072, 1024, ops.EPI_BIAS_HALF, dtype=torch.bfloat16))
    run("gemm-bert", lambda: gemm_case(32 * 40, 2304, 768, ops.EPI_BIAS_HALF))


def probe_rows():
    def ln():
        x = torch.randn(1000, 1024, device=dev) * 3 + 1
        w = torch.rand(1024, device=dev) + 0.5
        b = torch.randn(1024, device=dev)
        y16, y32 = ops.layernorm(x, w, b, 1e-6, want_f32=True)
        ref = torch.nn.functional.layer_norm(x, (1024,), w, b, 1e-6)
        report("layernorm f32", y32, ref, 1e-5)
        report("layernorm f16", y16, ref, 1e-3)
        x2 = torch.randn(77, 768, device=dev)
        w2, b2 = torch.rand(768, device=dev) + 0.5, torch.randn(768, device=dev)
        _, y = ops.layernorm(x2, w2, b2, 1e-12, want_f32=True)
        report("layernorm 768", y, torch.nn.functional.layer_norm(x2, (768,), w2, b2, 1e-12), 1e-5)
        # strided rows (CLS gather)
        x3 = torch.randn(5, 197, 1024, device=dev)
        _, y = ops.layernorm(x3, w, b, 1e-6, want_f32=True, rows=5, row_stride=197 * 1024)
        report("layernorm cls-rows", y, torch.nn.functional.layer_norm(x3[:, 0], (1024,), w, b, 1e-6), 1e-5)

    def l2():
        x = torch.randn(33, 768, device=dev)
        report("l2norm", ops.act_l2norm(x, 0), torch.nn.functional.normalize(x, dim=-1), 1e-6)
        report("tanh+l2norm", ops.act_l2norm(x, 1), torch.nn.functional.normalize(torch.tanh(x), dim=-1), 1e-6)

    run("layernorm", ln)
    run("l2norm", l2)


def probe_attention():
    def case(B, S, H, masked, dtype=torch.float16):
        g = torch.Generator(device="cpu").manual_seed(S * 3 + H)
        qkv = (torch.randn(B * S, 3 * H * 64, generator=g)).to(dtype).to(dev)
        mask = None
        if masked:
            lens = torch.randint(1, S + 1, (B,), generator=g)
            mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
        out = ops.attention(qkv, B, S, H, key_mask=mask)
        q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        bias = None
        if masked:
            bias = torch.zeros(B, 1, 1, S, device=dev).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=0.125)
        ref = ref.transpose(1, 2).reshape(B * S, H * 64)
        report(f"attention B{B} S{S} H{H} mask{int(masked)} {str(dtype)[6:]}", out, ref, 3e-3 if dtype == torch.float16 else 2e-2)

    run("att-vit", lambda: case(3, 197, 16, False))
    run("att-bert", lambda: case(5, 256, 12, True))
    run("att-short", lambda: case(7, 32, 12, True))
    run("att-odd", lambda: case(2, 77, 4, True))
    run("att-bf16", lambda: case(2, 197, 16, False, torch.bfloat16))


def bench_fc1_epilogues():
    """Same GEMM (fc1 shape) with different epilogues: isolates the epilogue cost."""
    M, N, K = 197 * 256, 4096, 1024
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.zeros(N, device=dev)
    for epi, nm in ((0, "bias->fp16"), (1, "bias+gelu->fp16"), (3, "bias->fp32")):
        out = None
        for _ in range(3):
            out = ops.gemm(a, w, epi, bias=bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, epi, bias=bias, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[perf] fc1-shape epi={nm}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)


def bench_attention():
    for (B, S, H, masked) in [(512, 197, 16, False), (2048, 32, 12, True)]:
        qkv = torch.randn(B * S, 3 * H * 64, device=dev).half()
        mask = None
        if masked:
            lens = torch.randint(4, S + 1, (B,))
            mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
        for _ in range(3):
            ops.attention(qkv, B, S, H, key_mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attention(qkv, B, S, H, key_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * B * H * S * S * 64
        print(f"[perf] attention B={B} S={S} H={H} mask={int(masked)}: {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s", flush=True)


def bench_similarity():
    for (N, P, group) in [(10000, 32, 2), (50000, 256, 4), (200000, 2, 0)]:
        feats = torch.randn(N, 768, device=dev)
        cls = torch.nn.functional.normalize(torch.randn(768, P, device=dev), dim=0)
        for tc in (True, False):
            for _ in range(3):
                ops.similarity(feats, cls, group=group, tensor_cores=tc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.similarity(feats, cls, group=group, tensor_cores=tc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            byts = (N * 768 + 768 * P) * 4 + 2 * N * P * 4
            print(f"[perf] similarity N={N} P={P} {'tf32-tcgen05' if tc else 'fp32-fma'}: {ms * 1e3:.1f} us = {byts / ms / 1e6:.0f} GB/s "
                  f"(algorithmic bytes incl. logits+probs; includes the torch allocs of the wrapper)", flush=True)


def probe_sim():
    def sim():
        feats = torch.randn(1000, 768, device=dev) * 2
        cls = torch.nn.functional.normalize(torch.randn(768, 32, device=dev), dim=0)
        logits, probs = ops.similarity(feats, cls, group=2, temp=10.0)
        ref = torch.nn.functional.normalize(feats, dim=-1) @ cls
        report("similarity logits", logits, ref, 1e-5)
        report("similarity probs", probs, torch.softmax(ref.view(1000, 16, 2) * 10, -1).view(1000, 32), 1e-5)
        cls2 = torch.nn.functional.normalize(torch.randn(768, 2, device=dev), dim=0)
        lg, pr = ops.similarity(feats, cls2)
        ref2 = torch.nn.functional.normalize(feats, dim=-1) @ cls2
        report("similarity P2", pr, torch.softmax(ref2 * 10, 1), 1e-5)

    def scores():
        N, K, Cc = 3000, 70, 4
        feats = torch.randn(N, 768, device=dev)
        cls = torch.nn.functional.normalize(torch.randn(768, K * Cc, device=dev), dim=0)
        s = ops.prompt_scores(feats, cls, K, Cc, workspace_mb=1)
        lg = (torch.nn.functional.normalize(feats, dim=-1) @ cls).view(N, K, Cc)
        top = lg.topk(2, dim=2).values
        ref = ((top[..., 0] - top[..., 1]) - (top[..., 0] + top[..., 1] - 1).abs()).mean(0)
        report("prompt_scores", s[None], ref[None], 1e-4)

    def refine():
        g = torch.Generator().manual_seed(5)
        xy = torch.randint(0, 40, (5000, 2), generator=g) * 224
        probs = torch.softmax(torch.randn(5000, 2, generator=g), 1)
        keep, ref_out = ops.refine(xy.to(dev), probs.to(dev), 224, True)
        # python restatement of the reference dict walk
        first = {}
        for i, (x, y) in enumerate(xy.tolist()):
            first.setdefault((x, y), i)
        exp_keep = torch.zeros(5000, dtype=torch.uint8)
        exp = torch.zeros(5000, 2)
        pn = probs.numpy()
        for (x, y), i in first.items():
            exp_keep[i] = 1
            cur = [pn[first[c]] for c in ((x - 224, y - 224), (x, y - 224), (x - 224, y), (x, y)) if c in first]
            exp[i] = torch.from_numpy(np.array(cur).mean(0))
        ok = bool((keep.cpu() == exp_keep).all())
        print(f"[{'ok' if ok else 'FAIL'}] refine keep mask")
        RESULTS.append(("refine keep", ok))
        exact = bool((ref_out.cpu() == exp).all())
        print(f"      refine bit-exact: {exact}")
        report("refine probs", ref_out.cpu(), exp, 1e-7)

    run("similarity", sim)
    run("prompt_scores", scores)
    run("refine", refine)


def bench_gemm():
    """Quick throughput look at the layer shapes (B=64 tiles)."""
    M = 197 * 256
    for (N, K, epi, nm) in [(3072, 1024, 0, "qkv"), (1024, 1024, 2, "proj"), (4096, 1024, 1, "fc1"), (1024, 4096, 2, "fc2")]:
        a = (torch.randn(M, K, device=dev) * 0.5).half()
        w = (torch.randn(N, K, device=dev) * 0.05).half()
        bias = torch.zeros(N, device=dev)
        resid = torch.zeros(M, N, device=dev) if epi == 2 else None
        out = resid if epi == 2 else None
        for _ in range(3):
            out = ops.gemm(a, w, epi, bias=bias, resid=resid, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, epi, bias=bias, resid=resid, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * M * N * K / ms / 1e9
        # cuBLAS for context
        for _ in range(3):
            torch.matmul(a, w.T)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            torch.matmul(a, w.T)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        print(f"[perf] {nm}: M={M} N={N} K={K}: {ms:.3f} ms = {tf:.0f} TFLOP/s   (cuBLAS fp16 no-epilogue {2.0 * M * N * K / ms2 / 1e9:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "rows", "attention", "sim", "perf"]
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    if "gemm" in which:
        probe_gemm()
    if "rows" in which:
        probe_rows()
    if "attention" in which:
        probe_attention()
    if "sim" in which:
        probe_sim()
    if "perf" in which and all(ok for n, ok in RESULTS if n.startswith("gemm")):
        run("perf", bench_gemm)
    if "epiperf" in which:
        run("epiperf", bench_fc1_epilogues)
    if "simperf" in which:
        run("simperf", bench_similarity)
    if "attperf" in which:
        run("attperf", bench_attention)
    nfail = sum(1 for _, ok in RESULTS if not ok)
    print(f"SUMMARY: {len(RESULTS) - nfail}/{len(RESULTS)} ok")
    sys.exit(1 if nfail else 0)
