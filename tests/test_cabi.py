"""The C-ABI library loads and exports every symbol include/snb200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

from structurednets_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "snb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_version_error_string_and_launch_counter(built_lib):
    lib = _lib.lib()
    assert lib.sn_version() == 100
    assert isinstance(lib.sn_last_error_string(), bytes)
    _lib.reset_launch_count()
    assert _lib.launch_count() == 0


def test_argument_errors_are_reported_not_thrown(built_lib):
    lib = _lib.lib()
    rc = lib.sn_sss_pack(None, None, None, None)
    assert rc != 0 and b"plan" in lib.sn_last_error_string()
    rc = lib.sn_psm_forward(None, 0, None, 0, None, 0, None, None, 4, 3, 2, None)
    assert rc != 0 and b"psm" in lib.sn_last_error_string()


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_lib.SnSssStage) == 64
    assert ctypes.sizeof(_lib.SnSssChunk) == 64
    assert ctypes.sizeof(_lib.SnSssPlan) == 12 * 4 + 2 * 8
    assert ctypes.sizeof(_lib.SnPsmEll) == 8 + 4 * 8
    assert ctypes.sizeof(_lib.SnPsmFactor) == 16 + 2 * 40 + 5 * 8
