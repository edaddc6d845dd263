"""CPU: the C-ABI library builds, loads and exports every symbol include/lgcn.h declares; host-side argument
checks work without a GPU (no compute calls here)."""
import ctypes
import subprocess

from lanegcn_b200 import _C


def test_header_symbols_are_bound_and_exported(lib):
    declared = _C.header_symbols()
    assert declared == sorted(_C._SIGS), "every function in include/lgcn.h needs a ctypes signature"
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", _C.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    assert set(declared) <= exported


def test_header_is_plain_c():
    p = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", _C.HEADER_PATH], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_version_and_sizes(lib):
    assert lib.lgcn_version() >= 100
    assert lib.lgcn_laneconv_wpack_floats(14) == 15 * 128 * 128 + 128 * 128 + 4 * 128
    assert lib.lgcn_att_wpack_floats() == 3 * 128 + 8 * 128 * 128 + 10 * 128
    assert lib.lgcn_laneconv_workspace_bytes(1000, 14) >= 1000 * 16 * 128 * 4
    assert lib.lgcn_att_workspace_bytes(10, 100) >= (3 * 100 + 2 * 10) * 512
    assert lib.lgcn_csr_workspace_bytes(10, 100) >= 440
    assert lib.lgcn_pairs_workspace_bytes(10, 2) >= 4 * (11 + 4 + 10)


def test_argument_errors_are_reported_not_fatal(lib):
    # rejected on the host before any CUDA call: bad ks, bad n_src, flags with several output blocks
    one = ctypes.c_void_p(16)
    rc = lib.lgcn_linear128(one, None, None, None, None, None, 1, one, 3, one, 1, None, None, None, 0, one, 128, 4, None)
    assert rc < 0 and b"ks" in lib.lgcn_last_error()
    rc = lib.lgcn_linear128(one, None, None, None, None, None, 4, None, 0, one, 1, None, None, None, 0, one, 128, 4, None)
    assert rc < 0 and b"n_src" in lib.lgcn_last_error()
    rc = lib.lgcn_linear128(one, None, None, None, None, None, 1, None, 0, one, 15, one, one, None, 1, one, 1920, 4, None)
    assert rc < 0 and b"n_out_blocks" in lib.lgcn_last_error()
    rc = lib.lgcn_offset_indices(one, 3, one, one, 1, 1, one, None)
    assert rc < 0 and b"idx_bytes" in lib.lgcn_last_error()
    rc = lib.lgcn_csr_build(None, None, None, 99, 10, one, one, one, one, None)
    assert rc < 0 and b"n_keys" in lib.lgcn_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _C.lib()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected a RuntimeError")
