#!/usr/bin/env python
"""Golden fixtures for the envelope / block-switch detector, from the REFERENCE itself.

Runs only in the build container (needs oracle/_ref/libvorbis_ref.so).  For each configuration the
reference's own _ve_envelope_search (lib/envelope.c:216) analyses a fresh vorbis_dsp_state that was fed
a probe signal through vorbis_analysis_buffer/_wrote; recorded are
  stream   the stream buffer the detector saw (blocksizes[1]/2 samples of preamble - rewritten by
           _preextrapolate_helper - followed by the input), float32 [ch][samples]
  marks    ve->mark[0 .. steps+VE_POST)
  state    envelope_lookup.stretch + envelope_filter_state[ch*VE_BANDS] after the search
and the setup fixtures gain the four vorbis_info_psy_global fields the detector reads (env_* keys).

usage:  python tests/golden/make_golden_envelope.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from vorbis_b200 import abi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CONFIGS = {"44k_stereo_q5": (2, 44100, 0.5, 24000), "44k_stereo_q1": (2, 44100, 0.1, 12000),
           "44k_mono_q4": (1, 44100, 0.4, 24000), "22k_mono_q3": (1, 22050, 0.3, 16000),
           "48k_6ch_q2": (6, 48000, 0.2, 8000)}


def signal(ch, rate, ns, seed):
    """noise + sines with a near-silent stretch, two bursts and a digital-silence gap: pre-echo and
    post-echo triggers in several bands, stretch resets, the near-DC refresh every 15 steps"""
    rng = np.random.default_rng(seed)
    t = np.arange(ns)
    pcm = np.stack([0.2 * rng.uniform(-1, 1, ns) + 0.4 * np.sin(2 * np.pi * (330 + 170 * c) * t / rate)
                    for c in range(ch)]).astype(np.float32)
    a = ns // 4
    pcm[:, a:a + 1500] *= 0.002
    pcm[:, a + 1500:a + 1700] = rng.uniform(-0.9, 0.9, (ch, 200)).astype(np.float32)
    pcm[:, 2 * a:2 * a + 900] = 0
    pcm[0, 3 * a:3 * a + 64] += 0.5
    return pcm


def main():
    for name, (ch, rate, q, ns) in CONFIGS.items():
        r = pyref.Ref(ch, rate, q)
        path = os.path.join(OUT, "setup_%s.npz" % name)
        old = abi.SetupHolder.load(path).arrays
        new = r.setup()
        for k, v in old.items():                       # nothing but the env_* keys may change
            assert np.array_equal(np.asarray(v), np.asarray(new.arrays[k])), (name, k)
        new.save(path)
        marks, steps, state, stream = r.envelope_marks(signal(ch, rate, ns, seed=77))
        np.savez_compressed(os.path.join(OUT, "envelope_%s.npz" % name), stream=stream, marks=marks,
                            steps=np.int32(steps), state=state)
        print(name, "steps", steps, "marks set", int(marks.sum()), "stretch", int(state[0]))
        r.close()


if __name__ == "__main__":
    main()
