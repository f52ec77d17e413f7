"""CPU: the C-ABI library builds, loads and exports every symbol include/pwr.h
declares; argument validation happens before any CUDA call."""
import ctypes
import os
import re

import pytest

from pixelwiseregression_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pwr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pwr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libpwr_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_signature_arity_matches_header():
    text = open(os.path.join(ROOT, "include", "pwr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, params in re.findall(r"\b(pwr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert len(_lib.SIGNATURES[name]) == n, name


def test_version_and_error_strings(lib):
    assert lib.pwr_version() == _lib.ABI_VERSION == 201
    assert lib.pwr_error_string(0) == b"ok"
    assert b"NULL" in lib.pwr_error_string(-1)
    assert b"aligned" in lib.pwr_error_string(-3)


def test_dispatch_options_are_explicit_not_environment(lib):
    assert lib.pwr_set_option(99, 1) == -4
    assert _lib.set_option("bwd_direct", 1) == 0
    assert _lib.set_option("bwd_direct", 0) == 1
    with _lib.option("fwd_pipe", 1):
        assert lib.pwr_set_option(_lib.OPTIONS["fwd_pipe"], 1) == 1
    assert lib.pwr_set_option(_lib.OPTIONS["fwd_pipe"], 0) == 0
    text = "".join(open(os.path.join(ROOT, "pixelwiseregression_b200", "csrc", f)).read() for f in ("decoder.cu", "sfr.cu"))
    assert "getenv" not in text                       # nothing on the launch path reads the environment


def test_stale_library_is_detected_by_content(lib, tmp_path, monkeypatch):
    assert not build.needs_build()
    stamp = open(build.STAMP).read()
    try:
        open(build.STAMP, "w").write("0" * 64 + "\n")
        assert build.needs_build()
    finally:
        open(build.STAMP, "w").write(stamp)
    assert not build.needs_build()


def test_argument_errors_without_a_gpu(lib):
    null = None
    fake = ctypes.c_void_p(0x1000)          # aligned, never dereferenced: validation fails first
    odd = ctypes.c_void_p(0x1004)           # misaligned
    # NULL required pointer
    assert lib.pwr_decoder_fwd(null, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 14, 0, 0, null) == -1
    # misaligned map pointer
    assert lib.pwr_decoder_fwd(odd, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 14, 0, 0, null) == -3
    # unknown method / bad shapes
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 14, 7, 0, null) == -4
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 0, 0, 0, null) == -2
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, -1, 14, 0, 0, null) == -2
    # unknown conv dtype, and caller-supplied heat maps exist in float32 only
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 14, 0, 5, null) == -4
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 1, 14, 2, 1, null) == -4
    assert lib.pwr_sfr_com(null, 480, 640, fake, 1, null) == -1
    assert lib.pwr_sfr_com(fake, 0, 640, fake, 1, null) == -2
    assert lib.pwr_reduce_partials(null, fake, 1, 1, 1, null) == -1
    # B == 0 is a successful no-op
    assert lib.pwr_decoder_fwd(fake, fake, fake, fake, fake, null, null, null, null, fake, fake, fake, null, 0, 14, 0, 0, null) == 0
    # one-pass last stage: (z, w, D, L, m, heat_gt, dmap_gt, uvd_gt, taps, alpha, lambda_h, lambda_d, loss_scale,
    #                       loss_scale_dev, n_mean, H, uvd, gz, gD, gw_partial, loss_partial, B, J, method, dtype, stream)
    fused = lambda **kw: lib.pwr_decoder_fwd_bwd_loss(
        kw.get("z", fake), fake, kw.get("D", fake), fake, fake, kw.get("heat", fake), fake, kw.get("uvd_gt", fake),
        kw.get("taps", null), 1.0, 1.0, 0.01, 1.0, null, 0, fake, fake, fake, fake, fake, fake, kw.get("B", 1), 14,
        kw.get("method", 0), 0, null)
    assert fused(z=null) == -1 and fused(D=null) == -1 and fused(uvd_gt=null) == -1      # depth branch is mandatory
    assert fused(heat=null) == -1                        # dense targets missing and no taps either
    assert fused(z=odd) == -3
    assert fused(method=2) == -4                         # PWR_METHOD_GIVEN has no last-stage loss
    assert fused(B=-1) == -2 and fused(B=0) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setenv("PWR_LIB_PATH", str(tmp_path / "nope.so"))      # an explicit path is never auto-built
    with pytest.raises(_lib.PwrError, match="no CPU fallback"):
        _lib.load()
