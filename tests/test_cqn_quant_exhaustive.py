"""The quantiser shortcut of k_cqn (vorbis_b200/csrc/vb200_cqn.cuh cqn_quant: sqrtf + one exact fp64 comparison instead
of rint(sqrt((double)ve)), lib/psy.c:959-963) checked EXHAUSTIVELY: all 1 249 902 592 floats in [0, 2^22).
tools/cqn_quant_check.c restates the two expressions on the host (IEEE sqrtf, as -prec-sqrt=true on the device)."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cqn_quant_shortcut_equals_double_sqrt_for_every_float_below_2p22(tmp_path):
    exe = str(tmp_path / "cqn_quant_check")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", os.path.join(ROOT, "tools", "cqn_quant_check.c"),
                           "-lm", "-o", exe])
    out = json.loads(subprocess.check_output([exe]).decode())
    assert out["stride"] == 1 and out["checked"] == 1249902592 and out["mismatches"] == 0
