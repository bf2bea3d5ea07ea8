"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/keep_b200.h declares.
No compute entry point is called here (there is no GPU); create() must fail loudly, not fall back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from keep_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "keep_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(keepb200_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert path.exists()
    L = _lib.lib()
    assert L.keepb200_version() == _lib.ABI_VERSION == 2


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    declared = _declared()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/keep_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES and the header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.lib_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (keepb200_\w+)", out))
    assert exported == set(declared)


def test_library_contains_blackwell_kernels():
    """SASS evidence that the hot kernels are tcgen05/TMA code, not a recompiled legacy path."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.lib_path())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    # no legacy warp-level tensor path left: every MMA in the library is tcgen05 (UTC*MMA); HMMA would be mma.sync
    assert not re.search(r"\bHMMA\b", sass)
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", str(_lib.lib_path())], capture_output=True, text=True).stdout


def test_no_environment_switch_changes_the_numerics():
    """The shipped library reads no environment variable (numerics are selected by arguments only), and has one GEMM
    code path per shape: no multicast-cluster / dynamic-scheduler / wide-epilogue variants left in the binary."""
    import glob

    for src in glob.glob(os.path.join(ROOT, "keep_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "keep_b200", "csrc", "*.h")):
        assert "getenv" not in open(src).read(), src
    # (the statically linked CUDA runtime imports getenv for its own CUDA_* variables: the sources are what is checked)
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        sass_names = subprocess.run([cuobjdump, "-sass", str(_lib.lib_path())], capture_output=True, text=True).stdout
        names = set(re.findall(r"Function : (\S+)", sass_names))
        assert not [n for n in names if "gemm2d" in n], "dynamic-scheduler GEMM variant still compiled"
        pair = [n for n in names if re.search(r"\d+gemm2_kernelI", n)]
        single = [n for n in names if re.search(r"\d+gemm_kernelI", n)]
        # one instantiation per (epilogue, 16-bit format, full / ragged row tiles) and main-loop variant: which one runs is a
        # function of the problem (shape and dtype) only
        assert len(pair) == 9 * 4 and len(single) == 9 * 4, (len(pair), len(single))


def test_config_struct_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "keep_b200.h")).read()
    body = text[text.index("typedef struct KeepB200Config {"):text.index("} KeepB200Config;")]
    fields = re.findall(r"^\s*(int32_t|float)\s+(\w+);", body, flags=re.M)
    assert [n for _, n in fields] == [n for n, _ in _lib.KeepB200Config._fields_]
    assert C.sizeof(_lib.KeepB200Config) == 4 * len(fields)


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    cfg = _lib.KeepB200Config()
    cfg.struct_size = C.sizeof(cfg)
    cfg.img_size, cfg.patch_size, cfg.vit_width, cfg.vit_depth, cfg.vit_heads, cfg.vit_mlp = 224, 16, 128, 1, 2, 256
    cfg.vit_ln_eps, cfg.proj_dim = 1e-6, 128
    cfg.vocab_size, cfg.hidden, cfg.layers, cfg.heads, cfg.intermediate, cfg.max_pos, cfg.type_vocab = 100, 128, 1, 2, 256, 64, 2
    cfg.bert_ln_eps = 1e-12
    h = C.c_void_p()
    rc = L.keepb200_create(C.byref(cfg), 0, C.byref(h))
    assert rc < 0 and h.value is None
    assert b"no CPU path" in L.keepb200_last_error() or b"CUDA" in L.keepb200_last_error()
    cfg.patch_size = 14
    assert L.keepb200_create(C.byref(cfg), 0, C.byref(h)) == -1  # argument errors are reported before device errors
