"""CPU oracle (oracle/vb_oracle.c) against the golden vectors recorded from the reference.
Bit-exact.  Runs anywhere gcc exists."""
import numpy as np
import pytest

from conftest import CONFIG_NAMES, assert_bits_equal, load_npz, load_setup, make_desc
from vorbis_b200 import lib as vlib


@pytest.fixture(scope="module", params=CONFIG_NAMES)
def cfg(request, oracle_lib):
    name = request.param
    setup = load_setup(name)
    return name, setup, oracle_lib.Oracle(setup), load_npz("encode", name), load_npz("decode", name)


@pytest.mark.parametrize("tag", ["L", "S"])
def test_transforms(cfg, tag):
    name, setup, o, enc, _ = cfg
    W = 1 if tag == "L" else 0
    N = setup.blocksize(W)
    pcm = enc[tag + "_pcm"].reshape(-1, N)
    lW = np.repeat(enc[tag + "_lW"], setup.channels)
    nW = np.repeat(enc[tag + "_nW"], setup.channels)
    win = o.apply_window(W, pcm, lW, nW)
    assert_bits_equal(win, enc[tag + "_windowed"].reshape(-1, N), "window")
    assert_bits_equal(o.mdct_forward(W, win), enc[tag + "_mdct_raw"].reshape(-1, N // 2), "mdct_forward")
    assert_bits_equal(o.drft_forward(W, win), enc[tag + "_fft"].reshape(-1, N), "drft_forward")


@pytest.mark.parametrize("tag", ["L", "S"])
def test_psy_stages_isolated(cfg, tag):
    name, setup, o, enc, _ = cfg
    W = 1 if tag == "L" else 0
    n = setup.blocksize(W) // 2
    ch = setup.channels
    for bt in (0, 1):
        sel = np.where(enc[tag + "_blocktype"] == bt)[0]
        if not len(sel):
            continue
        look = bt + 2 * W
        assert_bits_equal(o.noisemask(look, enc[tag + "_logmdct"][sel].reshape(-1, n)),
                          enc[tag + "_noise"][sel].reshape(-1, n), "noise look %d" % look)
        g = np.repeat(enc[tag + "_global_ampmax"][sel], ch)
        l = enc[tag + "_local_ampmax"][sel].reshape(-1)
        assert_bits_equal(o.tonemask(look, enc[tag + "_logfft"][sel].reshape(-1, n), g, l),
                          enc[tag + "_tone"][sel].reshape(-1, n), "tone look %d" % look)
        lm, m1 = o.offset_and_mix(look, 1, enc[tag + "_noise"][sel].reshape(-1, n),
                                  enc[tag + "_tone"][sel].reshape(-1, n),
                                  enc[tag + "_mdct_raw"][sel].reshape(-1, n),
                                  enc[tag + "_logmdct"][sel].reshape(-1, n))
        assert_bits_equal(lm, enc[tag + "_logmask"][sel].reshape(-1, n), "logmask")
        assert_bits_equal(m1, enc[tag + "_mdct_m1"][sel].reshape(-1, n), "mdct after M1")


@pytest.mark.parametrize("tag", ["L", "S"])
def test_phaseA_end_to_end(cfg, tag):
    name, setup, o, enc, _ = cfg
    W = 1 if tag == "L" else 0
    out = o.phaseA(W, enc[tag + "_pcm"], make_desc(enc, tag), taps=True)
    for k, g in (("mdct_raw", "mdct_raw"), ("logfft", "logfft"), ("noise", "noise"), ("tone", "tone"),
                 ("logmdct", "logmdct"), ("logmask", "logmask"), ("mdct", "mdct_m1")):
        assert_bits_equal(out[k], enc[tag + "_" + g], "phaseA " + k)
    assert_bits_equal(out["ampmax_out"], enc[tag + "_ampmax_out"], "ampmax_out")


def test_ampmax_chain(cfg):
    """in[k] = decay(max(in[k-1], out[k-1])) reproduces the recorded chain (lib/block.c:626-628)."""
    name, setup, o, enc, _ = cfg
    W, ain, aout = enc["chain_W"], enc["chain_ampmax_in"], enc["chain_ampmax_out"]
    g = np.float32(-9999.0)
    prev = g
    for k in range(len(W)):
        g = max(g, prev)
        g = np.float32(o.ampmax_decay(g, int(W[k])))
        assert g == ain[k], (k, g, ain[k])
        prev = aout[k]


@pytest.mark.parametrize("tag", ["L", "S"])
def test_couple_quantize_normalize(cfg, tag):
    name, setup, o, enc, _ = cfg
    W = 1 if tag == "L" else 0
    for bt in (0, 1):
        sel = np.where(enc[tag + "_blocktype"] == bt)[0]
        if not len(sel):
            continue
        iw, nz = o.couple_quantize_normalize(W, bt, 7, enc[tag + "_mdct_m1"][sel], enc[tag + "_ilogmask"][sel],
                                             enc[tag + "_nonzero_in"][sel])
        assert np.array_equal(iw, enc[tag + "_iwork_out"][sel])
        assert np.array_equal(nz, enc[tag + "_nonzero_out"][sel])


def test_decode(cfg):
    name, setup, o, _, dec = cfg
    bs = [setup.blocksize(0), setup.blocksize(1)]
    Wseq = dec["W"][None, :]
    coef_off, pcm_off, coef_len, pcm_len = vlib.synthesis_layout(Wseq, bs, setup.channels)
    assert coef_len == dec["coef"].size
    pcm = o.synthesis(Wseq, coef_off, dec["coef"], pcm_off, pcm_len)
    assert_bits_equal(pcm[0], dec["pcm"], "decoded pcm")
    W0 = int(dec["W"][0])
    first = o.mdct_backward(W0, dec["coef"][:setup.channels * bs[W0] // 2])
    assert_bits_equal(first, dec["imdct_first"], "mdct_backward")


def test_decouple_inverts_the_encoder_coupling(cfg):
    """oracle's de-coupling (lib/mapping0.c:754-779) undoes the encoder's lossless square-polar
    coupling (lib/psy.c:1128-1146): (A,B) -> (iM,iA) -> (A,B) for every integer pair."""
    name, setup, o, _, _ = cfg
    if setup.channels != 2 or int(setup.c.coupling_steps[1]) != 1:
        pytest.skip("needs a single-step stereo coupling")
    n = setup.blocksize(1) // 2
    rng = np.random.default_rng(0)
    A = rng.integers(-40, 41, n)
    B = rng.integers(-40, 41, n)
    A[:20] = 0; B[10:30] = 0
    iM, iA = A.copy(), B.copy()
    big = np.abs(A) > np.abs(B)
    iA = np.where(big, np.where(A > 0, A - B, B - A), np.where(B > 0, A - B, B - A))
    iM = np.where(big, A, B)
    flip = iA >= np.abs(iM) * 2
    iA = np.where(flip, -iA, iA)
    iM = np.where(flip, -iM, iM)
    res = np.zeros((1, 2, n), np.float32)
    res[0, int(setup.c.coupling_mag[1][0])] = iM
    res[0, int(setup.c.coupling_ang[1][0])] = iA
    out = o.decouple(1, res)
    assert np.array_equal(out[0, int(setup.c.coupling_mag[1][0])], A.astype(np.float32))
    assert np.array_equal(out[0, int(setup.c.coupling_ang[1][0])], B.astype(np.float32))


def floor_expect(enc, tag):
    """golden floor vectors as the API lays them out: NULL fits -> zero rows, fit_nonzero 0"""
    fit = enc[tag + "_fit_posts"].reshape(-1, enc[tag + "_fit_posts"].shape[-1]).astype(np.int32).copy()
    nz = (fit[:, 0] != -1).astype(np.int32)
    fit[nz == 0] = 0
    encp = enc[tag + "_enc_posts"].reshape(fit.shape).astype(np.int32)
    return fit, nz, encp


@pytest.mark.parametrize("tag", ["L", "S"])
def test_floor1_golden(cfg, tag):
    """floor1_fit (lib/floor1.c:576) and floor1_encode's quantise/predict/render (:765-945)."""
    name, setup, o, enc, _ = cfg
    W = 1 if tag == "L" else 0
    n = setup.blocksize(W) // 2
    fit, nz, encp = floor_expect(enc, tag)
    posts, got_nz = o.floor1_fit(W, enc[tag + "_logmdct"], enc[tag + "_logmask"])
    assert np.array_equal(got_nz, nz), "fit_nonzero"
    assert np.array_equal(posts, fit), "fit posts"
    p2, ilog, nz2 = o.floor1_render(W, posts, got_nz)
    assert np.array_equal(p2[nz == 1], encp[nz == 1]), "encode posts"
    assert np.array_equal(ilog, enc[tag + "_ilogmask"].reshape(-1, n)), "ilogmask"
    assert np.array_equal(nz2, enc[tag + "_nonzero_in"].reshape(-1)), "nonzero"


def test_floor1_golden_has_null_fit():
    enc = load_npz("encode", "44k_mono_q4")
    assert (enc["L_fit_posts"][:, :, 0] == -1).any(), "fixture should hold a silent block (floor1_fit == NULL)"


@pytest.mark.parametrize("name", CONFIG_NAMES)
@pytest.mark.parametrize("tag", ["L", "S"])
def test_oracle_encode_chain_golden(name, tag, oracle_lib):
    """the composed chain the one-call GPU entry point (vb200_encode_dsp) is checked against:
    Phase A -> floor1_fit -> floor render -> couple/quantise/normalise with per-block psy look"""
    setup = load_setup(name)
    enc = load_npz("encode", name)
    if not len(enc[tag + "_blocktype"]):
        pytest.skip("no such blocks")
    o = oracle_lib.Oracle(setup)
    r = o.encode_dsp(1 if tag == "L" else 0, enc[tag + "_pcm"], make_desc(enc, tag))
    assert np.array_equal(r["iwork"], enc[tag + "_iwork_out"])
    assert np.array_equal(r["nonzero"], enc[tag + "_nonzero_out"])
    ep = enc[tag + "_enc_posts"].astype(np.int32).copy()
    ep[enc[tag + "_fit_posts"][..., 0] == -1] = 0                       # NULL fit: the API returns a zero row
    assert np.array_equal(r["posts"], ep)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_oracle_envelope_golden(name, oracle_lib):
    """envelope / block-switch detector (lib/envelope.c): the oracle's trigger bits, replayed into marks,
    and its final filter state equal what the reference's _ve_envelope_search left behind"""
    setup = load_setup(name)
    env = load_npz("envelope", name)
    o = oracle_lib.Oracle(setup)
    steps = int(env["steps"])
    ret, state = o.envelope_search(env["stream"][None], 0, steps)
    assert np.array_equal(o.envelope_marks(ret[0])[:steps + 2], env["marks"])
    assert np.array_equal(state[0], env["state"])
    assert env["marks"].sum() > 0 and (ret & 2).any() and (ret & 1).any()
    # the search may be cut anywhere: state carries over
    a = steps // 2 + 1
    r1, s1 = o.envelope_search(env["stream"][None], 0, a)
    r2, s2 = o.envelope_search(env["stream"][None], a, steps - a, state=s1)
    assert np.array_equal(np.concatenate([r1, r2], 1), ret) and np.array_equal(s2, state)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_oracle_encode_decode_round_trip(name, oracle_lib):
    """the encode chain's posts + quantised residue, decoded by the decode chain (de-couple, floor multiply,
    IMDCT, overlap-add), reconstruct the PCM: a whole-codec property no single stage test covers"""
    from vorbis_b200 import abi
    setup = load_setup(name)
    o = oracle_lib.Oracle(setup)
    N, ch = setup.blocksize(1), setup.channels
    hop, ns, bps = N // 2, 2, 6
    stride = (bps - 1) * hop + N
    rng = np.random.default_rng(8)
    t = np.arange(stride)
    pcm = np.stack([[0.05 * rng.standard_normal(stride) + 0.4 * np.sin(2 * np.pi * (300 + 90 * c + 40 * s) * t / setup.rate)
                     for c in range(ch)] for s in range(ns)]).astype(np.float32)
    desc = np.zeros(ns * bps, abi.BLOCKDESC_DTYPE)
    desc["lW"] = 1; desc["nW"] = 1; desc["blocktype"] = 1
    blocks = np.stack([pcm[s, :, k * hop:k * hop + N] for s in range(ns) for k in range(bps)])
    enc = o.encode_dsp(1, blocks, desc, streams=(ns, bps))
    Wseq = np.ones((ns, bps), np.int32)
    coef_off, pcm_off, coef_len, pcm_len = vlib.synthesis_layout(Wseq, [setup.blocksize(0), N], ch)
    present = ((enc["posts"][..., 0] != 0) | (enc["posts"][..., 1] != 0) | (enc["nonzero"] != 0)).astype(np.int32)
    got = o.decode_dsp(Wseq, coef_off, enc["iwork"].astype(np.float32).reshape(-1), enc["posts"], present, pcm_off, pcm_len)
    ref = pcm[:, :, N // 2:N // 2 + pcm_len]
    err = got - ref
    assert 10 * np.log10((ref ** 2).sum() / (err ** 2).sum()) > 10.0
