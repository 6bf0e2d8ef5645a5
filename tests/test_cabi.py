"""CPU-side checks of the C-ABI boundary: the shared object loads, exports every symbol that
include/numpitron_b200.h declares, the ctypes table covers them all, and host-only entry points
(error text, workspace sizing, argument validation) behave.  No kernel is launched here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from numpitron_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "numpitron_b200.h"


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(npt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.declared_symbols()) == header_functions()


def test_struct_layouts_match_header():
    # sizes the C compiler gives the two argument structs (LP64): see include/numpitron_b200.h
    assert C.sizeof(_lib.GemmArgs) == 8 * 4 + 4 * 32 + 8 + 8 + 8 + 8 + 8 + 3 * 8 + 16 + 24 + 8 + 8 + 8  # + C_bf16_peer, peer_rows_lo/hi, skip_splitk_reduce (padded)
    assert C.sizeof(_lib.AdamTensor) == 64 + 16  # + g_splits, g_split_stride


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.npt_version() >= 100
    rc = lib.npt_gemm(None, None)
    assert rc != 0
    assert b"null args" in lib.npt_last_error()


def test_gemm_argument_validation_and_workspace_sizing():
    lib = _lib.load()
    g = _lib.GemmArgs()
    g.mode, g.M, g.N, g.K, g.batch, g.batch_inner = 0, 0, 8, 8, 1, 1
    assert lib.npt_gemm(C.byref(g), None) != 0 and b"bad shape" in lib.npt_last_error()
    # dW-shaped problem (small M x N, long K) asks for split-K scratch; a big square one does not
    g.M, g.N, g.K = 384, 1152, 16384
    need = lib.npt_gemm_workspace_bytes(C.byref(g))
    assert need > 0 and need % (384 * 1152 * 4) == 0
    g.M, g.N, g.K = 16384, 1152, 384
    assert lib.npt_gemm_workspace_bytes(C.byref(g)) == 0
    g.mode = 1
    g.M, g.N, g.K = 384, 1152, 16384
    assert lib.npt_gemm_workspace_bytes(C.byref(g)) > 0


def test_workspace_queries():
    lib = _lib.load()
    assert lib.npt_layernorm_bwd_workspace_bytes(16384, 384) >= 2 * 384 * 4
    assert lib.npt_colsum_workspace_bytes(16384, 1536) >= 1536 * 4
    assert lib.npt_embedding_scatter_add_workspace_bytes(16384, 65) >= 16384 * 4 + 2 * 66 * 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NptError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_compute_refuses_to_run_without_gpu():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from numpitron_b200 import distributed as npdist
    from numpitron_b200.nn import LayerNorm

    npdist.init(tp_size=1)
    ln = LayerNorm(32, rng=np.random.default_rng(0))
    with pytest.raises(_lib.NptError, match="no CPU fallback"):
        ln(np.zeros((1, 2, 32), np.float32))


def test_token_ids_must_fit_uint16():
    import numpy as np

    from numpitron_b200.nn.core import _host_to_device

    with pytest.raises(ValueError, match="uint16"):
        _host_to_device(np.array([[0, 70000]], dtype=np.int64))
    with pytest.raises(ValueError, match="uint16"):
        _host_to_device(np.array([[-1, 3]], dtype=np.int64))


@pytest.mark.parametrize("M,N", [(384, 1152), (384, 384), (384, 1536), (1536, 384)])
def test_weight_gradient_split_factor_is_one_wave(M, N):
    """npt_gemm_splits (host-only): the weight-gradient GEMMs of the Shakespeare shape (K = 16 384 tokens)
    must split into a number of units that fits ONE wave of the 148 SMs.  (The first version rounded the
    factor up — 18 tiles x 9 splits = 162 units = two waves — which cost 6 % of the step.)"""
    lib = _lib.load()
    g = _lib.GemmArgs()
    g.mode, g.trans_a, g.M, g.N, g.K, g.batch, g.batch_inner = 1, 1, M, N, 16384, 1, 1
    splits = lib.npt_gemm_splits(C.byref(g))
    assert splits > 1
    need = lib.npt_gemm_workspace_bytes(C.byref(g))
    assert need == splits * M * N * 4
    # tile widths the kernel can pick: 128 / 192 / 256 columns, 128 rows
    tiles_by_width = {bn: -(-M // 128) * -(-N // bn) for bn in (128, 192, 256)}
    assert any(t * splits <= 148 for t in tiles_by_width.values()), (splits, tiles_by_width)
    assert splits * 4 <= 16384 // 64          # at least 4 k-blocks per split
    # a launch that already fills the GPU is not split
    g.trans_a, g.M, g.N, g.K = 0, 16384, 1536, 384
    assert lib.npt_gemm_splits(C.byref(g)) == 1
