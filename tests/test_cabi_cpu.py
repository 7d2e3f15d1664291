"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/prl_b200.h declares, validates arguments without a GPU, and the host layer fails loudly
(no CPU fallback)."""

import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from prime_rl_cse_b200 import _lib

    return _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "prl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/prl_b200.h but not exported"
    assert sorted(lib.exported_symbols()) == declared, "ctypes signature table and header disagree"


def test_ctypes_signatures_match_the_header_argument_by_argument(lib):
    """Every prototype of include/prl_b200.h against the ctypes table: same number of parameters, and per parameter the same
    class (pointer / 64-bit integer / 32-bit integer / float / double / size_t) — an ABI drift between the header and the
    Python binding would otherwise only show up as garbage arguments on the GPU box."""
    import ctypes

    text = open(os.path.join(ROOT, "include", "prl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(prl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text)
    assert len(protos) == len(lib._SIGNATURES)

    def klass(decl):
        decl = decl.strip()
        if decl in ("void", ""):
            return None
        if "*" in decl:
            return "ptr"
        for key, val in (("int64_t", "i64"), ("int32_t", "i32"), ("size_t", "size"), ("float", "f32"), ("double", "f64"), ("int", "i32")):
            if re.search(rf"\b{key}\b", decl):
                return val
        raise AssertionError(f"unrecognised parameter {decl!r}")

    ct = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int64: "i64", ctypes.c_int32: "i32", ctypes.c_size_t: "size",
          ctypes.c_float: "f32", ctypes.c_double: "f64"}
    for ret, name, params in protos:
        res, args = lib._SIGNATURES[name]
        want = [k for k in (klass(p) for p in params.split(",")) if k is not None]
        assert [ct[a] for a in args] == want, f"{name}: header {want} vs ctypes {[ct[a] for a in args]}"
        assert ct[res] == klass(ret + (" x" if "*" not in ret else "")), f"{name}: return type"


def test_version_and_strerror(lib):
    L = lib.load()
    assert L.prl_version() >= 100
    assert L.prl_strerror(0) == b"ok"
    assert b"argument" in L.prl_strerror(-1)
    assert b"workspace" in L.prl_strerror(-4)


def test_argument_validation_needs_no_gpu(lib):
    L = lib.load()
    # n % group != 0 -> PRL_ERR_BAD_ARG before any CUDA call (advantage.py:90's assert)
    assert L.prl_advantages_f32(None, None, 3, 2, 0, None, None, None, None) == -1
    assert L.prl_advantages_f32(None, None, 4, 2, 9, None, None, None, None) == -1
    assert L.prl_advantages_f32(None, None, 0, 2, 0, None, None, None, None) == 0
    assert L.prl_lmhead_fwd(None, None, None, 8, 12, 100, 0, 100, 1.0, 1, 1, None, None, None, None, None, 0, None) == -1
    assert L.prl_loss_packed_fwd(None, None, None, None, None, 1, 8, 7, 0, 8.0, 1.0, None, None, None, None, None, None, 0, None) == -1
    assert L.prl_lmhead_fwd(None, None, None, 8, 16, 100, 0, 100, -1.0, 1, 1, None, None, None, None, None, 0, None) == -1  # temperature <= 0
    assert L.prl_lmhead_fwd_workspace_bytes(16384, 4096, 151936) >= 64 * 16384 * 16
    assert L.prl_lmhead_bwd_workspace_bytes(16384, 4096, 151936, 8192) >= 16384 * 8192 * 2
    # saved-logits pair: the row pitch covers V and keeps rows 16-byte aligned; a pitch below V or off alignment is refused
    for V in (1, 8, 1003, 151936):
        ld = L.prl_lmhead_saved_logits_ld(V)
        assert ld >= V and ld % 8 == 0
    assert L.prl_lmhead_fwd_save(None, None, None, 8, 16, 100, 0, 100, 1.0, 1, None, None, None, None, None, 96, None, 0, None) == -1
    assert L.prl_lmhead_fwd_save(None, None, None, 8, 16, 100, 0, 100, 1.0, 1, None, None, None, None, 16, 100, None, 0, None) == -1
    assert L.prl_lmhead_bwd_saved(None, None, None, None, None, 8, 16, 100, 0, 1.0, 1, None, 104, None, 0, None, 0, 0, 0, None, 0, None) == -1
    assert L.prl_lmhead_bwd_saved(None, None, None, None, None, 8, 16, 100, 0, 0.0, 1, None, 104, None, 0, None, 0, 0, 0, None, 0, None) == -1
    assert L.prl_lmhead_bwd_saved(None, None, None, None, None, 0, 16, 100, 0, 1.0, 1, None, 104, None, 0, None, 0, 0, 0, None, 0, None) == 0
    # fused compute + collective entry points: the peer tables are validated on the host before anything is launched
    import ctypes

    two = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    scat = lambda T, rows, self_idx=0, peers=2, tab=two: L.prl_lmhead_bwd_scatter(  # noqa: E731
        None, None, None, None, None, T, 64, 1000, 0, 1.0, 1, 1, tab, peers, rows, self_idx, None, 1, 1, 0, 0, None, 0, None)
    assert scat(80, 40) == -2            # a 32-row epilogue slab must not straddle two owners: PRL_ERR_UNSUPPORTED
    assert scat(96, 32) == -1            # T != n_peers * rows_per_peer
    assert scat(64, 32, self_idx=2) == -1 and scat(64, 32, tab=None) == -1 and scat(288, 32, peers=9) == -1
    assert scat(64, 32, tab=(ctypes.c_void_p * 2)(0x1000, 0)) == -1   # a peer that was never mapped
    dws = lambda V, rows, peers=2, tab=two, local=None: L.prl_lmhead_bwd_dw_scatter(  # noqa: E731
        None, None, None, None, None, 8, 64, V, 1.0, 1, 1, None, 0, tab, peers, rows, 0, local, 0, 0, None, 0, None)
    assert dws(1000, 400) == -1          # rows_per_peer * n_peers must cover V (ceil(V / n_peers) rows per owner)
    assert dws(40, 20) == -1             # fewer than 32 vocab rows per owner
    assert dws(1000, 500, tab=(ctypes.c_void_p * 2)(0x1000, 0)) == -1
    assert dws(1000, 500, local=ctypes.c_void_p(0x1004)) == -1        # the local accumulator is read with 16-byte loads
    # the optional workspace is only asked for where a K-split of the d_hidden GEMM pays off (few output tiles)
    assert L.prl_lmhead_bwd_saved_workspace_bytes(16384, 4096, 151936) == 0
    assert L.prl_lmhead_bwd_saved_workspace_bytes(4096, 1024, 151936) >= 4096 * 1024 * 4


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """The boundary is a C ABI, not a C++ or torch one: a C99 translation unit that includes the header compiles with
    -pedantic, links against the shared library and gets the documented error code back (no GPU involved)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text(
        '#include "prl_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  int rc = prl_advantages_f32(NULL, NULL, 3, 2, PRL_ADV_DRGRPO, NULL, NULL, NULL, NULL);\n"
        '  printf("%d %d %s|%s\\n", prl_version(), rc, prl_strerror(rc), prl_last_error_detail());\n'
        "  return (rc == PRL_ERR_BAD_ARG && prl_lmhead_saved_logits_ld(1003) >= 1003) ? 0 : 1;\n}\n")
    libdir = os.path.dirname(lib.LIB_PATH)
    exe = tmp_path / "client"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:" + os.path.basename(lib.LIB_PATH), "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "samples_per_problem" in r.stdout, r.stdout + r.stderr


def test_no_cpu_fallback():
    import prime_rl_cse_b200 as P

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.get_response_lengths(torch.arange(8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.compute_loss([torch.zeros(4)], [torch.zeros(4)], [torch.zeros(4)], [torch.ones(4, dtype=torch.bool)], P.LossConfig(), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.fused_lm_head_logprobs_entropy(torch.zeros(1, 4, 8, dtype=torch.bfloat16), torch.zeros(16, 8, dtype=torch.bfloat16),
                                         torch.zeros(1, 4, dtype=torch.long), 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "prime_rl_cse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_loss_config_mirror():
    import prime_rl_cse_b200 as P

    c = P.LossConfig()
    assert (c.ratio_type, c.ratio_length_norm, c.clip_ratio) == ("token", False, 8.0)  # rl/config.py:22-27
    with pytest.raises(ValueError):
        P.LossConfig(ratio_type="gae")
    with pytest.raises(ValueError):
        P.LossConfig(clip_ratio=-1.0)
