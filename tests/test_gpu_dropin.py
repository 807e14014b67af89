"""Drop-in proof: the reference's own encoder and decoder (unmodified sources, real API loop)
with lib/mapping0.c's hot callees AND lib/block.c's envelope search (the block-size decisions of
vorbis_analysis_blockout) bound to the CUDA library through the reference-signature shims
(vorbis_b200/host/vb200_ref_shim.c) must produce byte-identical packets and bit-identical decoded PCM.  Needs oracle/_ref/*.so (built in the container where /root/reference exists;
the .so files travel to the GPU box)."""
import numpy as np
import pytest

from conftest import probe_signal
from oracle import pyref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ch,rate,q", [(2, 44100, 0.5), (1, 44100, 0.4), (2, 44100, 0.1), (6, 48000, 0.2)])
def test_encoder_packets_identical(cuda_ok, ch, rate, q):
    if not (pyref.available() and pyref.dropin_available()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    pcm = probe_signal(ch, rate, 0.6, seed=5)
    ref = pyref.Ref(ch, rate, q)
    ref.encode_capture(pcm, fields=())
    want = ref.packets()
    got_enc = pyref.Ref(ch, rate, q, dropin=True)
    l0 = got_enc.L.vb200shim_launches()
    got_enc.encode_capture(pcm, fields=())
    got = got_enc.packets()
    assert got_enc.L.vb200shim_launches() - l0 > 100, "the CUDA path did not run"
    assert len(got) == len(want) and len(want) > 10
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "packet %d differs" % i
    # decode the same packets through the shimmed mdct_backward
    d_ref = ref.decode_capture(len(want) + 4, pcm.shape[1] + 8192, fields=())
    d_got = got_enc.decode_capture(len(want) + 4, pcm.shape[1] + 8192, fields=())
    assert d_ref["pcm"].shape == d_got["pcm"].shape
    assert np.array_equal(d_ref["pcm"].view(np.uint32), d_got["pcm"].view(np.uint32))
    got_enc.close()
    ref.close()
