"""CPU: the C-ABI library loads, exports every symbol include/usl.h declares,
its struct layouts agree with the ctypes mirrors, and argument validation
returns error codes (no compute: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

from uncertainty_model_b200 import _build, _lib

HEADER = os.path.join(ROOT, 'include', 'usl.h')


@pytest.fixture(scope='module')
def lib():
    _build.build()
    return _lib.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(usl_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f'{name} not exported by libusl.so'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings(lib):
    assert lib.usl_version() == 200
    assert lib.usl_strerror(0) == b'ok'
    assert lib.usl_strerror(-1) == b'invalid argument'
    assert lib.usl_strerror(-3) == b'unsupported shape'
    assert lib.usl_strerror(-99) == b'unknown error'


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / 'layout.c'
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "usl.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(UslLossConfig), offsetof(UslLossConfig, coef),
         sizeof(UslLossScale), offsetof(UslLossScale, grad_recon_out));
  printf("%zu %zu %zu %zu\\n", offsetof(UslLossScale, images),
         offsetof(UslLossScale, err_in), offsetof(UslLossScale, recon_out),
         offsetof(UslLossScale, grad_disp));
  printf("%d %d %d\\n", USL_MAX_SCALES, USL_NUM_TERMS, USL_VERSION);
  return 0;
}''')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'),
                    str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True,
                         text=True).stdout.split()
    got = [int(v) for v in out]
    S, Cfg = _lib.UslLossScale, _lib.UslLossConfig
    want = [C.sizeof(Cfg), Cfg.coef.offset, C.sizeof(S),
            S.grad_recon_out.offset, S.images.offset, S.err_in.offset,
            S.recon_out.offset, S.grad_disp.offset, _lib.USL_MAX_SCALES,
            _lib.USL_NUM_TERMS, 200]
    assert got == want


def test_argument_validation_returns_codes(lib):
    assert lib.usl_pyramid(None, 1, 6, 8, 8, 0, 0, 4, None, None) == -1
    assert lib.usl_warp_fwd(None, 0, 1.0, None, 0, 0, 1, 3, 8, 8, None, 0, 0,
                            None) == -1
    s = _lib.UslLossScale()
    s.B, s.h, s.w = 1, 2, 2
    assert lib.usl_loss_fwd_ctas(C.byref(s)) == -1
    s.h, s.w = 64, 128
    assert lib.usl_loss_fwd_ctas(C.byref(s)) >= 1
    cfg = _lib.UslLossConfig()
    assert lib.usl_loss_fwd(C.byref(cfg), C.byref(s), 1, None, None) == -1
    assert lib.usl_loss_bwd(C.byref(cfg), C.byref(s), 1, None, None, 0,
                            None) == -1
    assert lib.usl_spars_workspace_bytes(2, 8, 8, 11, 0) == 0
    n = (40 - 10) * (56 - 10)
    assert lib.usl_spars_workspace_bytes(4, 40, 56, 11, 0) >= 4 * n * 16
    assert lib.usl_spars_curve(None, None, 1, 40, 56, 11, None, 100, None,
                               None, None, None, None, 0, None) == -1
    assert lib.usl_spars_ause(None, None, 100, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.UslError, match='no CPU fallback'):
        _lib.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import
    or execute it."""
    pkg = os.path.join(ROOT, 'uncertainty_model_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text,
                                     flags=re.M), f
                assert 'emu_harness' not in text, f
