"""The C-ABI library loads and exports every symbol include/vorbis_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from vorbis_b200 import abi, lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vorbis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vb200_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libvorbis_b200.so does not export %s" % n
    assert sorted(lib.EXPORTS) == names, "vorbis_b200/lib.py EXPORTS out of sync with the header"


def test_struct_sizes_match_header():
    # compile-time layout check through the oracle library, which is built from the same header
    from oracle import pyoracle
    pyoracle.build()
    assert ctypes.sizeof(abi.BlockDesc) == 16
    assert ctypes.sizeof(abi.PhaseAIO) == 10 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(abi.EncodeIO) == 128          # static_assert in vorbis_b200/csrc/vb200.cu
    # vb200_psy_setup: 2 ints, 7 floats(2+3+1+1), int, 40 floats, float, 3 ints, pad, double, 4 ints, float, pad, 5 ptrs
    assert ctypes.sizeof(abi.PsySetup) % 8 == 0


def test_bad_arguments_return_error_codes_not_aborts():
    L = lib.load()
    s = abi.Setup()
    s.blocksizes[0], s.blocksizes[1] = 100, 2048       # not a power of two
    s.channels = 2
    h = ctypes.c_void_p()
    rc = L.vb200_ctx_create(ctypes.byref(s), 0, ctypes.byref(h))
    assert rc == -131, rc                                # OV_EINVAL
    assert b"power" in L.vb200_last_error()
    assert L.vb200_mdct_forward(None, 0, 1, None, None) == -131


def test_envelope_apply_marks_is_plain_host_code():
    """vb200_envelope_apply_marks (lib/envelope.c:254-264 replayed from trigger bits) needs no GPU:
    against the oracle's restatement on random bits, and against the reference's marks on the fixtures"""
    import numpy as np
    from conftest import CONFIG_NAMES, load_npz, load_setup
    from oracle import pyoracle
    L = lib.load()
    L.vb200_envelope_apply_marks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.vb200_envelope_apply_marks.restype = None
    rng = np.random.default_rng(2)
    ret = rng.integers(0, 8, 300).astype(np.uint8) * (rng.random(300) < 0.2)
    ret = np.ascontiguousarray(ret, np.uint8)
    for first in (0, 7):
        mark = np.zeros(first + len(ret) + 2, np.int32)
        L.vb200_envelope_apply_marks(ret.ctypes.data, first, len(ret), mark.ctypes.data)
        o = pyoracle.Oracle(load_setup(CONFIG_NAMES[0]))
        assert np.array_equal(mark, o.envelope_marks(ret, first_step=first))
    for name in CONFIG_NAMES:
        env = load_npz("envelope", name)
        o = pyoracle.Oracle(load_setup(name))
        steps = int(env["steps"])
        r, _ = o.envelope_search(env["stream"][None], 0, steps)
        mark = np.zeros(steps + 2, np.int32)
        L.vb200_envelope_apply_marks(np.ascontiguousarray(r[0]).ctypes.data, 0, steps, mark.ctypes.data)
        assert np.array_equal(mark, env["marks"])
