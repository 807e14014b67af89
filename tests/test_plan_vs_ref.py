"""Block planning: the oracle's restatement of what vorbis_analysis_blockout decides per block
(W, lW, nW, blocktype, position) must equal the reference's own block sequence, captured while the unmodified
reference encodes a stream through its public API (oracle/_ref).  Needs /root/reference (oracle/_ref)."""
import numpy as np
import pytest

from conftest import probe_signal
from oracle import pyoracle, pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (no /root/reference)")

GRID = [(2, 44100, .5), (1, 44100, .4), (2, 44100, .1), (1, 22050, .3), (6, 48000, .2), (2, 48000, .9), (2, 32000, 0.)]


def burst_signal(ch, rate, secs, seed):
    rng = np.random.default_rng(seed)
    ns = int(rate * secs)
    t = np.arange(ns)
    pcm = np.stack([0.1 * rng.uniform(-1, 1, ns) + 0.4 * np.sin(2 * np.pi * (300 + 70 * c) * t / rate)
                    for c in range(ch)]).astype(np.float32)
    for _ in range(6):
        a = int(rng.integers(2000, ns - 3000))
        pcm[:, a:a + 200] *= 0.02
        pcm[:, a + 200:a + 260] = rng.uniform(-0.9, 0.9, (ch, 60))
    return pcm


def reference_stream(ch, rate, q, pcm, fields=("pcm",)):
    r = pyref.Ref(ch, rate, q)
    cap = r.encode_capture(pcm, fields=fields, timeline=True)
    return r, cap


@pytest.mark.parametrize("mode", ["probe", "bursts"])
@pytest.mark.parametrize("ch,rate,q", GRID)
def test_oracle_plan_equals_reference_block_sequence(ch, rate, q, mode):
    pcm = probe_signal(ch, rate, 1.5, 7) if mode == "probe" else burst_signal(ch, rate, 1.5, 7)
    r, cap = reference_stream(ch, rate, q, pcm)
    o = pyoracle.Oracle(r.setup())
    tl = cap["timeline"]
    mark, nsteps = o.timeline_marks(tl[None])
    plan, nb = o.plan_blocks(mark, nsteps, [tl.shape[1]], [cap["eof"]])
    k = cap["nblocks"]
    assert nb[0] == k and (cap["W"] == 0).sum() >= 5           # the signals do switch block sizes
    for name in ("W", "lW", "nW", "blocktype"):
        assert np.array_equal(plan[0, :k][name], cap[name][:k]), name
    for b in range(k):                                          # positions: the block is that slice of the timeline
        N = r.bs[cap["W"][b]]
        p = plan[0, b]["pos"]
        assert np.array_equal(cap["pcm"][b][:, :N], tl[:, p:p + N]), "block %d position" % b


def test_config1_plumbing_numbers():
    """BASELINE config 1: 1 s mono 44.1 kHz 440 Hz sine (0.8 amplitude, float), q=0.4 through the
    reference API.  The committed driver (oracle/ref_driver.c ref_encode_capture: 1024-sample
    vorbis_analysis_wrote calls, then wrote(0); audio packets only) gives 46 blocks = 2 short + 44 long and
    1705 packet bytes.  (SURVEY §8d quotes 47 / 2+45 / 1851 from a survey-time probe whose source was not
    kept; block count and bytes depend on the write chunking - 45..46 blocks, 1615..1705 bytes for chunks
    of 256..44100 samples - so the pinned numbers are the ones this repository can reproduce.)"""
    t = np.arange(44100)
    pcm = (0.8 * np.sin(2 * np.pi * 440.0 * t / 44100.0)).astype(np.float32)[None]
    r, cap = reference_stream(1, 44100, .4, pcm)
    assert cap["nblocks"] == 46
    assert int((cap["W"] == 0).sum()) == 2 and int((cap["W"] == 1).sum()) == 44
    assert cap["bytes"] == 1705
    o = pyoracle.Oracle(r.setup())
    mark, nsteps = o.timeline_marks(cap["timeline"][None])
    plan, nb = o.plan_blocks(mark, nsteps, [cap["timeline"].shape[1]], [cap["eof"]])
    assert nb[0] == 46 and np.array_equal(plan[0, :46]["W"], cap["W"])
