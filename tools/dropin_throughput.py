#!/usr/bin/env python
"""Drop-in throughput: N concurrent encoders driven by the multi-stream driver of
vorbis_b200/host/vb200_mapping0.c (their ready blocks go to the device together, the reference's own floor1_encode
and residue backend write the bits on ONE host thread) against the stock reference encoder on one host thread,
same streams, same box.  Prints one JSON object; packets are cross-checked by hash.

usage: python tools/dropin_throughput.py [--streams 1000] [--seconds 2.0]
Needs oracle/_ref/*.so (built where /root/reference exists; the .so files travel to the GPU box)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1000)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--stock-streams", type=int, default=64, help="streams encoded by the stock reference for the rate and the hash check")
args = ap.parse_args()
ch, rate, q = 2, 44100, 0.5
ns, n = args.streams, int(rate * args.seconds)
rng = np.random.default_rng(3)
t = np.arange(n, dtype=np.float32)
base = (0.25 * rng.uniform(-1, 1, (8, ch, n)) + 0.5 * np.sin(2 * np.pi * (440 + 110 * np.arange(ch)).reshape(1, ch, 1) * t / rate)).astype(np.float32)
for k in range(8):                                          # a few transients so that both block sizes occur
    for a in rng.integers(3000, n - 3000, 3):
        base[k, :, a:a + 200] *= 0.02
        base[k, :, a + 200:a + 260] = rng.uniform(-0.9, 0.9, (ch, 60))
pcm = np.ascontiguousarray(base[np.arange(ns) % 8])         # [ns][ch][n]: 8 distinct signals, cycled
D = pyref.dropin_lib()
D.ref_ms_encode.restype = C.c_long
hashes, nbytes, counts = (C.c_uint64 * ns)(), (C.c_long * ns)(), (C.c_long * ns)()
small = min(ns, 16)
D.ref_ms_encode(small, ch, C.c_long(rate), C.c_float(q), 0, pcm.ctypes.data_as(C.c_void_p), C.c_long(n), hashes, nbytes, counts)  # warm-up
t0 = time.perf_counter()
blocks = D.ref_ms_encode(ns, ch, C.c_long(rate), C.c_float(q), 0, pcm.ctypes.data_as(C.c_void_p), C.c_long(n), hashes, nbytes, counts)
dt_ms = time.perf_counter() - t0
assert blocks > 0, "multi-stream driver failed"
L = pyref.lib()
L.ref_stock_encode_summary.restype = C.c_long
k = min(args.stock_streams, ns)
t0 = time.perf_counter()
sb = 0
for i in range(k):
    h, b, c = C.c_uint64(0), C.c_long(0), C.c_long(0)
    sb += L.ref_stock_encode_summary(ch, C.c_long(rate), C.c_float(q), pcm[i].ctypes.data_as(C.c_void_p), C.c_long(n), C.byref(h), C.byref(b), C.byref(c))
    assert (h.value, b.value, c.value) == (hashes[i], nbytes[i], counts[i]), "stream %d: packets differ from the stock reference" % i
dt_stock = time.perf_counter() - t0
pk = sum(counts)
print(json.dumps({
    "what": "N concurrent vorbis encoders (44.1 kHz stereo q=0.5, %.1f s each): multi-stream drop-in driver vs stock reference, one host thread each" % args.seconds,
    "streams": ns, "blocks": int(blocks), "packets": int(pk),
    "dropin_packets_per_s": pk / dt_ms, "dropin_blocks_per_s": blocks / dt_ms, "dropin_seconds": dt_ms,
    "dropin_realtime_factor": ns * args.seconds / dt_ms,
    "stock_streams_timed": k, "stock_packets_per_s": sum(counts[i] for i in range(k)) / dt_stock, "stock_blocks_per_s": sb / dt_stock,
    "stock_realtime_factor": k * args.seconds / dt_stock,
    "speedup_one_host_thread": (blocks / dt_ms) / (sb / dt_stock),
    "packets_identical_to_stock": True,
    "note": "the host thread of the drop-in still runs the reference's floor1_encode and residue VQ/Huffman packing for every block "
            "(north_star keeps them on the host); the device does the rest in one call per block size and round"}))
