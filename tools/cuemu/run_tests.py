#!/usr/bin/env python
"""Run the `-m gpu` parity tests (or any pytest selection) against the HOST EMULATION build of the CUDA
sources (tools/cuemu): a development aid for checking kernel logic in a GPU-less container.
Nothing here is product code; the product library is never replaced on disk.

usage: python tools/cuemu/run_tests.py [pytest args...]     (default: tests -m gpu -x -q)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import build_emu  # noqa: E402

path = build_emu.build(verbose=False)
from vorbis_b200 import lib  # noqa: E402

lib.LIB_PATH = path            # this process only
import pytest  # noqa: E402

args = sys.argv[1:] or ["tests", "-m", "gpu", "-x", "-q"]
sys.exit(pytest.main(args))
