#!/usr/bin/env python
"""bench.py — encode Phase-A (window + MDCT + FFT + noise/tone masks + mix) throughput.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (default N=1)
  python bench.py --impl reference ...                   the reference's CPU code on the host cores

Workload (BASELINE.json configs[2]): 44.1 kHz stereo, vorbis_encode_init_vbr q=0.5, 100 000
long blocks (N=2048 samples -> 1024 spectral lines per channel: the "N=1024" of the metric;
SURVEY.md §8d) per GPU per step, synthetic PCM, independent blocks with ampmax given per block
(drop-in semantics).  One step = one pass of the hot path over that batch.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
METRIC = "stereo_blocks_per_sec_mdct_psy"
UNIT = "blocks/s"
W_LONG = 1


def workload_name(nblocks, N, ch):
    return ("mapping0_forward Phase A (window+MDCT+FFT+noise/tone mask+mix), 44.1kHz stereo q=0.5, "
            "%d long blocks x %d ch x N=%d samples (n=%d lines/ch)" % (nblocks, ch, N, N // 2))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def synth_pcm_torch(torch, nblocks, ch, N, rate, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(N, device=device, dtype=torch.float32)
    pcm = torch.rand((nblocks, ch, N), generator=g, device=device, dtype=torch.float32)
    pcm.mul_(0.5).sub_(0.25)                                     # 0.25*uniform(-1,1)
    f = 440.0 + 110.0 * torch.arange(ch, device=device, dtype=torch.float32).view(1, ch, 1)
    ph = torch.rand((nblocks, 1, 1), generator=g, device=device) * 6.2831853
    pcm.add_(0.5 * torch.sin(2 * np.pi * f * t.view(1, 1, N) / rate + ph))
    return pcm


def make_desc(nblocks):
    from vorbis_b200 import abi
    d = np.zeros(nblocks, abi.BLOCKDESC_DTYPE)
    d["lW"] = 1; d["nW"] = 1; d["blocktype"] = 1          # steady-state long blocks (psy look 3)
    d["ampmax"] = -6.0
    return d


def cpu_reference_rate(setup_name, W, pcm_np, desc_np, threads):
    """Reference CPU implementation of the same Phase A on `threads` host threads (the library is
    single threaded; blocks are independent, one reference instance per thread).  Returns
    (blocks/s, kind).  Uses oracle/_ref (the compiled reference) when present, else the oracle port."""
    from concurrent.futures import ThreadPoolExecutor
    from vorbis_b200 import abi
    nb = pcm_np.shape[0]
    shards = [s for s in np.array_split(np.arange(nb), threads) if len(s)]
    from oracle import pyref
    if pyref.available():
        kind = "reference"
        ch = pcm_np.shape[1]
        insts = [pyref.Ref(ch, 44100, 0.5) for _ in shards]

        def run(i):
            insts[i].phaseA_batch(W, pcm_np[shards[i]], desc_np[shards[i]])
    else:
        kind = "port"
        from oracle import pyoracle
        setup = abi.SetupHolder.load(os.path.join(GOLD, "setup_%s.npz" % setup_name))
        insts = [pyoracle.Oracle(setup) for _ in shards]

        def run(i):
            insts[i].phaseA(W, pcm_np[shards[i]], desc_np[shards[i]])
    with ThreadPoolExecutor(len(shards)) as ex:
        list(ex.map(run, range(len(shards))))           # warm caches / page in
        t0 = time.perf_counter()
        list(ex.map(run, range(len(shards))))
        dt = time.perf_counter() - t0
    return nb / dt, kind, len(shards)


# ------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vorbis_b200 import abi
    setup = abi.SetupHolder.load(os.path.join(GOLD, "setup_44k_stereo_q5.npz"))
    N, ch = setup.blocksize(W_LONG), setup.channels
    cores = os.cpu_count() or 1
    per_core = args.ref_blocks_per_core
    nb = per_core * cores
    rng = np.random.default_rng(1)
    t = np.arange(N, dtype=np.float32)
    pcm = (0.25 * rng.uniform(-1, 1, (nb, ch, N)) +
           0.5 * np.sin(2 * np.pi * (440 + 110 * np.arange(ch)).reshape(1, ch, 1) * t / 44100.0
                        + rng.uniform(0, 6.28, (nb, 1, 1)))).astype(np.float32)
    desc = make_desc(nb)
    rates = []
    kind = used = None
    for i in range(args.warmup + args.steps):
        r, kind, used = cpu_reference_rate("44k_stereo_q5", W_LONG, pcm, desc, cores)
        if i >= args.warmup:
            rates.append(r)
    total_t = sum(nb / r for r in rates)
    value = nb * len(rates) / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(rates),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(100000, N, ch), "l2": "n/a (CPU)",
                   "note": "each step is a bounded sample of the workload: %d blocks" % nb},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": "%d long stereo blocks (%d per thread) per step" % (nb, per_core)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def extra_configs(torch, lib, abi, device, peak):
    """BASELINE configs[1] (batched mdct_forward N=1024 x 65536) and configs[3] (decode: IMDCT +
    overlap-add, mixed 256/2048 blocks), device resident, CUDA events.  Informational."""
    out = {}
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # config 2: mdct_init(1024), 65536 vectors uniform(-1,1)
    s22 = abi.SetupHolder.load(os.path.join(GOLD, "setup_22k_mono_q3.npz"))
    c22 = lib.Context(s22, device=device)
    N = s22.blocksize(1)
    nv = 65536
    g = torch.Generator(device=dev); g.manual_seed(12345)
    x = torch.rand((nv, N), generator=g, device=dev) * 2 - 1
    y = torch.empty((nv, N // 2), device=dev)
    ms = timed(lambda: c22.mdct_forward_dev(1, nv, x.data_ptr(), y.data_ptr(), stream))
    out["mdct_forward_N1024_x65536"] = {"ms": ms, "transforms_per_s": nv / ms * 1e3,
                                         "algorithmic_GBps": 6 * N * nv / ms / 1e6,
                                         "frac_of_hbm_peak": 6 * N * nv / ms / 1e6 / peak}
    c22.close()
    # config 4: decode, 4096 stereo streams x 33 blocks, one run of 8 short blocks per 24 long
    s44 = abi.SetupHolder.load(os.path.join(GOLD, "setup_44k_stereo_q5.npz"))
    c44 = lib.Context(s44, device=device)
    bs = [s44.blocksize(0), s44.blocksize(1)]
    ns, nblk = 4096, 33
    Wrow = np.ones(nblk, np.int32); Wrow[12:20] = 0
    Wseq = np.tile(Wrow, (ns, 1))
    coef_off, pcm_off, coef_len, pcm_len = lib.synthesis_layout(Wseq, bs, s44.channels)
    coef = (torch.rand(coef_len, generator=g, device=dev) * 2 - 1) * 1e-2
    pcm = torch.zeros((ns, s44.channels, pcm_len), device=dev)
    dW = torch.from_numpy(Wseq).to(dev); dco = torch.from_numpy(coef_off).to(dev); dpo = torch.from_numpy(pcm_off).to(dev)
    ms = timed(lambda: c44.synthesis_dev(ns, nblk, dW.data_ptr(), dco.data_ptr(), coef.data_ptr(), dpo.data_ptr(),
                                         pcm.data_ptr(), pcm_len, stream))
    byts = 4 * (coef_len + ns * s44.channels * pcm_len)
    out["decode_4096streams_x33blocks_mixed"] = {"ms": ms, "stereo_blocks_per_s": ns * nblk / ms * 1e3,
                                                  "algorithmic_GBps": byts / ms / 1e6,
                                                  "frac_of_hbm_peak": byts / ms / 1e6 / peak}
    # the whole per-block encode DSP on the device - Phase A, floor1_fit, floor render,
    # couple/quantise/normalise - 20000 long stereo blocks, buffers resident
    nb, N, chn = 20000, bs[1], s44.channels
    n = N // 2
    pcm_e = synth_pcm_torch(torch, nb, chn, N, 44100, dev, 99)
    d_desc = torch.from_numpy(make_desc(nb).view(np.uint8).reshape(-1, 16).copy()).to(dev)
    o_m = torch.empty((nb, chn, n), device=dev); o_lm = torch.empty_like(o_m); o_mask = torch.empty_like(o_m)
    o_amp = torch.empty(nb, device=dev)
    posts = torch.empty((nb * chn, abi.FLOOR1_STRIDE), dtype=torch.int32, device=dev)
    fnz = torch.empty(nb * chn, dtype=torch.int32, device=dev)
    iwork = torch.empty((nb, chn, n), dtype=torch.int32, device=dev)
    nz = torch.empty(nb * chn, dtype=torch.int32, device=dev)
    io = abi.PhaseAIO()
    io.pcm, io.desc = pcm_e.data_ptr(), d_desc.data_ptr()
    io.mdct, io.logmdct, io.logmask, io.ampmax_out = o_m.data_ptr(), o_lm.data_ptr(), o_mask.data_ptr(), o_amp.data_ptr()
    stages = {
        "phaseA": lambda: c44.phaseA_dev(1, nb, io, stream=stream),
        "floor1_fit": lambda: c44.floor1_fit_dev(1, nb * chn, o_lm.data_ptr(), o_mask.data_ptr(), posts.data_ptr(),
                                                 fnz.data_ptr(), stream=stream),
        "floor1_render": lambda: c44.floor1_render_dev(1, nb * chn, posts.data_ptr(), fnz.data_ptr(), iwork.data_ptr(),
                                                       nz.data_ptr(), stream=stream),
    }

    def fit_render():                   # render rewrites posts in place: time it behind a fresh fit
        stages["floor1_fit"]()
        stages["floor1_render"]()

    def render_cqn():                   # CQN rewrites iwork in place: time it behind a fresh fit + render
        fit_render()
        c44.couple_quantize_normalize_dev(1, 1, 7, nb, o_m.data_ptr(), iwork.data_ptr(), nz.data_ptr(), stream=stream)
    chain = {}
    chain["phaseA_ms"] = timed(stages["phaseA"], reps=3)
    chain["floor1_fit_ms"] = timed(stages["floor1_fit"], reps=3)
    fr = timed(fit_render, reps=3)
    chain["floor1_render_ms"] = fr - chain["floor1_fit_ms"]
    stages["couple_quantize_normalize"] = lambda: c44.couple_quantize_normalize_dev(
        1, 1, 7, nb, o_m.data_ptr(), iwork.data_ptr(), nz.data_ptr(), stream=stream)
    chain["couple_quantize_normalize_ms"] = timed(render_cqn, reps=3) - fr

    def whole():
        for fn in stages.values():
            fn()
    ms = timed(whole, reps=3)
    chain["whole_chain_ms"] = ms
    chain["stereo_blocks_per_s"] = nb / ms * 1e3
    chain["blocks"] = nb
    out["encode_chain_phaseA_floor1_cqn_20000_long_stereo"] = chain
    c44.close()
    return out


def run_ours(args):
    import torch
    from vorbis_b200 import abi, lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    setup = abi.SetupHolder.load(os.path.join(GOLD, "setup_44k_stereo_q5.npz"))
    ctx = lib.Context(setup, device=local)       # raises if the CUDA library is missing
    N, ch = setup.blocksize(W_LONG), setup.channels
    n = N // 2
    nb = args.blocks
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    # every rank owns its own shard of independent blocks (weak scaling: nb per GPU, no collective)
    pcm = synth_pcm_torch(torch, nb, ch, N, setup.rate, dev, seed=1000 + rank)
    desc_np = make_desc(nb)
    desc = torch.from_numpy(desc_np.view(np.uint8).reshape(nb, 16).copy()).to(dev)
    mdct = torch.empty((nb, ch, n), device=dev, dtype=torch.float32)
    logmdct = torch.empty_like(mdct)
    logmask = torch.empty_like(mdct)
    amp = torch.empty(nb, device=dev, dtype=torch.float32)
    io = abi.PhaseAIO()
    io.pcm, io.desc = pcm.data_ptr(), desc.data_ptr()
    io.mdct, io.logmdct, io.logmask, io.ampmax_out = mdct.data_ptr(), logmdct.data_ptr(), logmask.data_ptr(), amp.data_ptr()

    def step():
        ctx.phaseA_dev(W_LONG, nb, io, stream=sptr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * nb * args.steps / (ms_max * 1e-3)

    # per-kernel durations (CUDA events on the launching stream, inside the library) for the roofline
    ctx.set_profiling(True)
    kms = np.zeros(3)
    reps = max(3, args.steps)
    for _ in range(reps):
        step()
        torch.cuda.synchronize()
        kms += np.array(ctx.phaseA_kernel_ms())
    kms /= reps
    ctx.set_profiling(False)

    line = None
    if rank == 0:
        peak, peak_src = load_peaks()
        names = ["k_phaseA_transform", "k_ampmax", "k_phaseA_psy"]
        # algorithmic bytes per channel-block: transform reads 4N, writes mdct 2N + logfft 2N;
        # psy reads mdct 2N + logfft 2N, writes mdct' 2N + logmdct 2N + logmask 2N.  Phase A fused: 10N.
        alg = [8 * N, 0, 10 * N]
        dom = int(np.argmax(kms))
        achieved = alg[dom] * ch * nb / (kms[dom] * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bytes per
        # (block,channel) row, profiles/summary.json), scaled to this launch
        traffic = None
        try:
            summ = json.load(open(os.path.join(ROOT, "profiles", "summary.json")))
            per_row = summ["kernels"][names[dom]]["dram_bytes_per_row"]
            traffic = per_row * ch * nb
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": {k: float(v) for k, v in zip(names, kms)},
                "phaseA_algorithmic_GBps": 10 * N * ch * nb / (kms.sum() * 1e-3) / 1e9}

        # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside)
        nb_e = min(nb, args.e2e_blocks)
        hp = torch.empty((nb_e, ch, N), dtype=torch.float32).pin_memory()
        hp.copy_(pcm[:nb_e].cpu())
        outs = [torch.empty((nb_e, ch, n), dtype=torch.float32).pin_memory() for _ in range(3)]
        hamp = torch.empty(nb_e, dtype=torch.float32).pin_memory()
        hdesc = desc_np[:nb_e].copy()
        hio = abi.PhaseAIO()
        hio.pcm, hio.desc = hp.data_ptr(), hdesc.ctypes.data
        hio.mdct, hio.logmdct, hio.logmask = (o.data_ptr() for o in outs)
        hio.ampmax_out = hamp.data_ptr()
        L = lib.load()

        def e2e_step():
            rc = L.vb200_analysis_phaseA(ctx.h, W_LONG, nb_e, C.byref(hio))
            if rc:
                raise RuntimeError("vb200_analysis_phaseA failed: %d" % rc)
        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        dt = time.perf_counter() - t0
        e2e = {"value": nb_e * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(hp.numel() * 4 + hdesc.nbytes),
               "d2h_bytes_per_step": int(sum(o.numel() for o in outs) * 4 + hamp.numel() * 4),
               "blocks_per_step": nb_e}

        # ---- CPU baseline on a bounded sample of the same workload
        cores = os.cpu_count() or 1
        nb_c = args.ref_blocks_per_core * cores
        pcm_c = pcm[:min(nb, nb_c)].cpu().numpy()
        if pcm_c.shape[0] < nb_c:
            pcm_c = np.concatenate([pcm_c] * (nb_c // pcm_c.shape[0] + 1))[:nb_c]
        rate_c, kind, used = cpu_reference_rate("44k_stereo_q5", W_LONG, pcm_c, make_desc(nb_c), cores)
        cpu = {"value": rate_c, "unit": UNIT, "cores": used, "kind": kind,
               "sample": "%d long stereo blocks (%d per thread)" % (nb_c, args.ref_blocks_per_core)}

        extra = None
        if not args.no_extra:
            try:
                extra = extra_configs(torch, lib, abi, local, peak)
            except Exception as e:  # the headline line must still be printed
                extra = {"error": repr(e)}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(nb, N, ch), "blocks_per_gpu": nb,
                       "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2" % (18 * N * ch * nb / 1e9),
                       "sharding": "independent blocks per rank, no collective"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=100000, help="stereo blocks per GPU per step")
    ap.add_argument("--e2e-blocks", type=int, default=50000)
    ap.add_argument("--no-extra", action="store_true", help="skip the informational configs 2/4")
    ap.add_argument("--ref-blocks-per-core", type=int, default=512)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
