#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Runs only in the build container (needs oracle/_ref/libvorbis_ref.so, i.e. the
unmodified reference sources compiled by oracle/Makefile).  The reference's own
tests hold no golden vectors for this path, so these files are the
pin: every vector below was produced by the reference's real API loop
(vorbis_analysis_blockout -> vorbis_analysis, vorbis_synthesis ->
vorbis_synthesis_blockin) and recorded at the call boundaries of mapping0.c.

  setup_<cfg>.npz    lookup tables of vorbis_encode_init_vbr(ch, rate, q)
  encode_<cfg>.npz   per-block vectors of mapping0_forward for a few blocks
  decode_<cfg>.npz   spectra entering mdct_backward, block flags, finished PCM

usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # name: (channels, rate, quality, seconds, max long blocks kept, max short blocks kept)
    "44k_stereo_q5": (2, 44100, 0.5, 1.5, 5, 6),
    "44k_stereo_q1": (2, 44100, 0.1, 1.0, 3, 4),   # noise normalisation active
    "44k_mono_q4": (1, 44100, 0.4, 1.0, 3, 2),     # BASELINE config 1 shape (uncoupled)
    "22k_mono_q3": (1, 22050, 0.3, 1.0, 3, 2),     # 512/1024 blocks
    "48k_6ch_q2": (6, 48000, 0.2, 0.5, 2, 2),      # 5.1: four coupling steps, noise norm active
}


def signal(ch, rate, secs, seed):
    """0.25*uniform + 0.5*sine (the survey's probe signal) with a quiet gap and a burst so the
    envelope detector emits short blocks."""
    rng = np.random.default_rng(seed)
    ns = int(rate * secs)
    t = np.arange(ns)
    pcm = np.stack([0.25 * rng.uniform(-1, 1, ns) + 0.5 * np.sin(2 * np.pi * (440 + 110 * c) * t / rate)
                    for c in range(ch)]).astype(np.float32)
    for c in range(1, ch):                        # correlated channels so that coupling has work to do
        pcm[c] = (0.6 * pcm[0] + 0.4 * pcm[c]).astype(np.float32)
    a = ns // 2
    pcm[:, a:a + 300] *= 0.01
    pcm[:, a + 300:a + 400] = rng.uniform(-0.9, 0.9, (ch, 100)).astype(np.float32)
    return pcm


def main():
    for name, (ch, rate, q, secs, nlong, nshort) in CONFIGS.items():
        r = pyref.Ref(ch, rate, q)
        setup = r.setup()
        setup.save(os.path.join(OUT, "setup_%s.npz" % name))
        bs = r.bs
        pcm = signal(ch, rate, secs, seed=1234)
        if name == "44k_mono_q4":
            pcm[:, :3000] = 0                     # digital silence: floor1_fit returns NULL there
        cap = r.encode_capture(pcm)
        W = cap["W"]
        longs = np.where(W == 1)[0]
        shorts = np.where(W == 0)[0]
        # keep: the first blocks (stream start, ampmax = -9999), a transition, steady state
        keep_l = list(longs[:2]) + list(longs[len(longs) // 2: len(longs) // 2 + nlong - 2])
        keep_s = list(shorts[:nshort])
        enc = {"bs": np.array(bs, np.int32), "channels": np.int32(ch)}
        for tag, idx, N in (("L", keep_l, bs[1]), ("S", keep_s, bs[0])):
            idx = np.array(sorted(idx), np.int64)
            n = N // 2
            enc[tag + "_index"] = idx
            for k in ("W", "lW", "nW", "blocktype", "ampmax_in", "ampmax_out", "global_ampmax",
                      "local_ampmax", "nonzero_in", "nonzero_out"):
                enc[tag + "_" + k] = cap[k][idx]
            for k in ("pcm", "windowed", "fft"):
                enc[tag + "_" + k] = cap[k][idx][:, :, :N]
            for k in ("mdct_raw", "logfft", "logmdct", "noise", "tone", "logmask", "mdct_m1",
                      "ilogmask", "iwork_out"):
                enc[tag + "_" + k] = cap[k][idx][:, :, :n]
            # floor 1: what floor1_fit returned ([0] = -1 where it returned NULL) and what
            # floor1_encode left in post[] after its quantise / predict pass
            enc[tag + "_fit_posts"] = cap["fit_posts"][idx]
            enc[tag + "_enc_posts"] = cap["enc_posts"][idx]
        # the whole ampmax chain (cheap): lets the stream-mode test replay it
        enc["chain_W"] = W
        enc["chain_ampmax_in"] = cap["ampmax_in"]
        enc["chain_ampmax_out"] = cap["ampmax_out"]
        np.savez_compressed(os.path.join(OUT, "encode_%s.npz" % name), **enc)

        # decode: first 14 packets (contains short/long mixes near the start for these signals)
        d = r.decode_capture(cap["nblocks"] + 4, pcm.shape[1] + 4 * bs[1])
        # choose a window of blocks that includes long->short->long transitions
        Wd = d["W"]
        k0 = 0
        sh = np.where(Wd == 0)[0]
        if len(sh):
            k0 = max(0, int(sh[0]) - 3)
        k1 = min(len(Wd), k0 + 14)
        dec = {"bs": np.array(bs, np.int32), "channels": np.int32(ch), "W": Wd[k0:k1]}
        coefs = []
        for k in range(k0, k1):
            N = bs[Wd[k]]
            coefs.append(d["dec_coef"][k][:, :N // 2].reshape(-1))
        dec["coef"] = np.concatenate(coefs)
        # finished PCM of those blocks: block k finishes bs[W[k-1]]/4+bs[W[k]]/4 samples
        fin = np.zeros(len(Wd), np.int64)
        for k in range(1, len(Wd)):
            fin[k] = bs[Wd[k - 1]] // 4 + bs[Wd[k]] // 4
        start = int(fin[:k0 + 1].sum())      # samples finished by blocks <= k0
        stop = int(fin[:k1].sum())
        dec["pcm"] = d["pcm"][:, start:stop]
        dec["imdct_first"] = d["dec_imdct"][k0][:, :bs[Wd[k0]]]
        np.savez_compressed(os.path.join(OUT, "decode_%s.npz" % name), **dec)
        print(name, "blocks", cap["nblocks"], "long kept", len(keep_l), "short kept", len(keep_s),
              "decode blocks", k1 - k0, "pcm", dec["pcm"].shape)
        r.close()


if __name__ == "__main__":
    main()
