"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/clk.h declares, and
the product path fails loudly (no fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "clk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clk_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("clk_conv3x3_fprop", "clk_conv3x3_dgrad", "clk_conv3x3_wgrad", "clk_convT2x2_fprop",
                 "clk_ce_kd_loss", "clk_confusion_matrix", "clk_adam_multi_tensor", "clk_bn_finalize"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/clk.h but not exported by libclk.so"


def test_binding_table_matches_header(lib_built):
    from continual_learning_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared_symbols()
    lib = _lib.load()
    assert lib.clk_version() >= 100


def test_built_for_sm100a_with_tcgen05_and_tma(lib_built):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass  # tcgen05.mma
    assert "UTMALDG" in sass                      # TMA loads
    assert "LDTM" in sass                         # tcgen05.ld


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(lib_built):
    import continual_learning_b200 as clk
    from continual_learning_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ensure_device()
    m = clk.UNet(21)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clk.CrossEntropyDistillLoss()(torch.zeros(1, 21, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        clk.metrics.eval_metrics(torch.zeros(1, 4, 4, dtype=torch.long), torch.zeros(1, 4, 4, dtype=torch.long), 22)


def test_bad_arguments_return_status_codes(lib_built):
    from continual_learning_b200 import _lib
    lib = _lib.load()
    assert lib.clk_set_tuning(b"no_such_key", 1) == -1
    assert b"unknown tuning key" in lib.clk_last_error()
    # argument validation happens before any CUDA call
    assert lib.clk_conv3x3_fprop(None, 64, None, 0, None, None, None, None, None, 1, 16, 16, 64, 1, None) == -1
    assert lib.clk_im2col3x3_stem(1, 1, 1, 9, 16, 16, None) == -2  # Cin*9 > 64 is an unsupported shape
    assert lib.clk_confusion_matrix(None, None, 0, 64, 1, None, None) == -2  # nc > 36
