// vb200.cu — kernels (__global__) and the C ABI of include/vorbis_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -Xcompiler -fPIC -shared   (see __graft_entry__.build()).
// No torch, no oracle/ dependency: this is the product library.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "vorbis_b200.h"
#include "vb200_tables.h"
#include "vb200_kernels.cuh"
#include "vb200_cqn.cuh"
#include "vb200_psy2.cuh"
#include "vb200_psy3.cuh"
#include "vb200_floor1.cuh"
#include "vb200_env.cuh"
#include "vb200_res.cuh"
#include "vb200_streams.cuh"
#include "vb200_managed.cuh"
#include "floor1_db_table.h"

using namespace vb200;

// ======================================================================== //
// error plumbing
static thread_local std::string g_err;
static int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
  return code;
}
#define CU(call)                                                         \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) return fail(VB200_EFAULT, #call, e_);         \
  } while (0)

extern "C" const char *vb200_last_error(void) { return g_err.c_str(); }

// ======================================================================== //
// context
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct vb200_ctx {
  int device = 0;
  int sm_count = 148;
  vb200_setup setup;                 // scalar copy (pointers not used after create)
  HostXform hx[2];
  XformDev dx[2];
  WinDev dwin;
  int n_psy = 0;
  HostPsyFlow hflow[4];
  PsyDev dpsy[4];
  std::vector<void *> owned;         // device allocations freed at destroy
  std::atomic<uint64_t> launches{0};
  // grow-only scratch for the host-buffer entry points and phase A intermediates
  DevBuf scratch[16];
  DevBuf lane_buf[2][10];            // per-lane device buffers of the pipelined host Phase-A path
  cudaStream_t s_lane[2] = {nullptr, nullptr};
  const ResDev *d_res[2] = {nullptr, nullptr};        // [VB200_MAX_SUBMAPS] residue class parameters per block size
  int res_partvals[2] = {0, 0};
  EnvDev env;                        // envelope detector tables (N = 128 transform, windows, thresholds)
  DevBuf env_buf[4];                 // scratch of vb200_envelope_search[_dev]
  int grid_div = 1;                  // see grid_for
  cudaStream_t s_split[2] = {nullptr, nullptr};       // vb200_encode_dsp_dev: two concurrent half-batches
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  DevBuf enc_buf[8];                 // scratch of vb200_encode_dsp_dev
  DevBuf str_buf[40];                // scratch of vb200_plan_blocks / vb200_encode_streams[_dev]
  DevBuf mgd_buf[24];                // scratch + host-call staging of vb200_encode_dsp_managed[_dev]
  // vb200_encode_dsp (host buffers): chunks rotate over ENC_SETS buffer sets; one stream per copy direction and
  // two compute streams, ordered by events (see there)
  DevBuf enc_lane[4][17];
  cudaStream_t s_enc[3] = {nullptr, nullptr, nullptr};   // compute 0, compute 1, host->device
  cudaStream_t s_d2h = nullptr;
  cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr}, ev_cmp[4] = {nullptr, nullptr, nullptr, nullptr},
              ev_d2h[4] = {nullptr, nullptr, nullptr, nullptr};
  int psy_ctas_per_sm = 5;
  int psy_carveout_ctas = -1;        // CTAs/SM the generic psy kernel's shared-memory carve-out was last set for (per device)
  // The *_dev entry points keep their intermediates in per-context scratch: two calls in flight on different
  // user streams would share it.  Every such call waits for the previous one's event and records its own.
  cudaEvent_t ev_scratch = nullptr;
  bool scratch_busy = false;
  const float *d_fromdB = nullptr;
  const int *d_mag[2] = {nullptr, nullptr}, *d_ang[2] = {nullptr, nullptr};
  const Floor1Dev *d_floor[2] = {nullptr, nullptr};   // [VB200_MAX_SUBMAPS] per block size
  const unsigned char *d_chmux[2] = {nullptr, nullptr};
  cudaStream_t s_main = nullptr;
  std::mutex mu;
  // optional per-kernel timing of the last Phase-A call (bench roofline evidence)
  bool profiling = false;
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

static int scratch_begin(vb200_ctx *c, cudaStream_t st) {
  if (c->scratch_busy) CU(cudaStreamWaitEvent(st, c->ev_scratch, 0));
  return 0;
}
static int scratch_end(vb200_ctx *c, cudaStream_t st) {
  CU(cudaEventRecord(c->ev_scratch, st));
  c->scratch_busy = true;
  return 0;
}

template <class T>
static int upload(vb200_ctx *c, const T *src, size_t count, const T **dst) {
  void *d = nullptr;
  size_t bytes = sizeof(T) * (count ? count : 1);
  CU(cudaMalloc(&d, bytes));
  c->owned.push_back(d);
  if (count) CU(cudaMemcpy(d, src, sizeof(T) * count, cudaMemcpyHostToDevice));
  *dst = reinterpret_cast<const T *>(d);
  return 0;
}

static int ensure(vb200_ctx *c, int slot, size_t bytes, void **out) {
  DevBuf &b = c->scratch[slot];
  if (b.cap < bytes) {
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
  }
  *out = b.p;
  return 0;
}

extern "C" int vb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" void vb200_ctx_destroy(vb200_ctx *c);
// everything of vb200_ctx_create that can fail after the context exists; the caller destroys *c on failure
static int ctx_build(vb200_ctx *c, const vb200_setup *s, int device) {
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
  for (auto &st : c->s_lane) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto &st : c->s_enc) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  for (auto &e : c->ev_h2d) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : c->ev_cmp) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : c->ev_d2h) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &st : c->s_split) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming));
  for (auto &e : c->ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

  for (int w = 0; w < 2; w++) {
    HostXform &h = c->hx[w];
    build_xform(h, s->blocksizes[w], s->window[w]);
    XformDev &d = c->dx[w];
    memset(&d, 0, sizeof(d));
    d.N = h.N; d.log2n = h.log2n; d.nst = h.log2n - 6; d.nf = h.nf; d.scale = h.scale;
    for (int i = 0; i < h.nf && i < 8; i++) d.fac[i] = h.fac[i];
    for (size_t i = 0; i < h.stage_off.size() && i < 8; i++) d.stage_off[i] = h.stage_off[i];
    int rc;
    if ((rc = upload(c, h.trig.data(), h.trig.size(), &d.trig))) return rc;
    if ((rc = upload(c, h.bitrev.data(), h.bitrev.size(), &d.bitrev))) return rc;
    const float *tw = nullptr;
    if ((rc = upload(c, h.stage_tw.data(), h.stage_tw.size(), &tw))) return rc;
    d.stage_tw = reinterpret_cast<const float2 *>(tw);
    if ((rc = upload(c, h.win.data(), h.win.size(), &d.win))) return rc;
    if ((rc = upload(c, h.wa.data(), h.wa.size(), &d.wa))) return rc;
    c->dwin.N[w] = h.N;
    c->dwin.win[w] = d.win;
  }
  for (int w = 0; w < 2; w++) {      // residue classification parameters (lib/backends.h:103-118)
    ResDev hr[VB200_MAX_SUBMAPS];
    memset(hr, 0, sizeof(hr));
    for (int sm = 0; sm < VB200_MAX_SUBMAPS; sm++) {
      const vb200_residue_setup &r = s->residue[w][sm];
      ResDev &d = hr[sm];
      d.type = -1;
      if (r.type < 0 || r.grouping <= 0) continue;         // not provided (zero-initialised setups: grouping 0)
      if (r.type > 2 || r.begin < 0 || r.end < r.begin || r.partitions < 1 || r.partitions > 64)
        return fail(VB200_EINVAL, "residue setup");
      d.type = r.type; d.begin = r.begin; d.end = r.end; d.grouping = r.grouping; d.partitions = r.partitions;
      d.partvals = (r.end - r.begin) / r.grouping;
      d.scale = (float)(100. / r.grouping);
      for (int k = 0; k < 64; k++) { d.cm1[k] = r.classmetric1[k]; d.cm2[k] = r.classmetric2[k]; }
      if (d.partvals > c->res_partvals[w]) c->res_partvals[w] = d.partvals;
    }
    int rc;
    if ((rc = upload(c, hr, (size_t)VB200_MAX_SUBMAPS, &c->d_res[w]))) return rc;
  }
  {
    // envelope detector lookups, _ve_envelope_init (lib/envelope.c:31-74)
    HostXform h;
    build_xform(h, ENV_N, nullptr);
    EnvDev &E = c->env;
    memset(&E, 0, sizeof(E));
    XformDev &d = E.X;
    d.N = h.N; d.log2n = h.log2n; d.nst = h.log2n - 6; d.nf = h.nf; d.scale = h.scale;
    for (size_t i = 0; i < h.stage_off.size() && i < 8; i++) d.stage_off[i] = h.stage_off[i];
    int rc;
    if ((rc = upload(c, h.trig.data(), h.trig.size(), &d.trig))) return rc;
    if ((rc = upload(c, h.bitrev.data(), h.bitrev.size(), &d.bitrev))) return rc;
    const float *tw = nullptr;
    if ((rc = upload(c, h.stage_tw.data(), h.stage_tw.size(), &tw))) return rc;
    d.stage_tw = reinterpret_cast<const float2 *>(tw);
    float win[ENV_N];
    for (int i = 0; i < ENV_N; i++) {
      win[i] = (float)std::sin(i / (ENV_N - 1.) * M_PI);
      win[i] = win[i] * win[i];
    }
    if ((rc = upload(c, win, (size_t)ENV_N, &E.win))) return rc;
    static const int B[VB200_VE_BANDS] = {2, 4, 6, 9, 13, 17, 22}, En[VB200_VE_BANDS] = {4, 5, 6, 8, 8, 8, 8};
    float bwin[VB200_VE_BANDS * 8] = {0};
    for (int j = 0; j < VB200_VE_BANDS; j++) {
      float total = 0.f;
      for (int i = 0; i < En[j]; i++) {
        bwin[j * 8 + i] = (float)std::sin((i + .5) / En[j] * M_PI);
        total = total + bwin[j * 8 + i];
      }
      E.total[j] = (float)(1. / (double)total);
      E.begin[j] = B[j]; E.end[j] = En[j];
      E.preecho[j] = s->preecho_thresh[j]; E.postecho[j] = s->postecho_thresh[j];
    }
    if ((rc = upload(c, bwin, (size_t)VB200_VE_BANDS * 8, &E.bwin))) return rc;
    E.stretch_penalty = s->stretch_penalty; E.minenergy = s->preecho_minenergy;
  }
  {
    int rc;
    if ((rc = upload(c, VB_FLOOR1_FROMDB, (size_t)256, &c->d_fromdB))) return rc;
    for (int w = 0; w < 2; w++) {
      if (s->coupling_steps[w] < 0 || s->coupling_steps[w] > VB200_MAX_COUPLING) return fail(VB200_EINVAL, "coupling_steps");
      if ((rc = upload(c, (const int *)s->coupling_mag[w], (size_t)s->coupling_steps[w], &c->d_mag[w]))) return rc;
      if ((rc = upload(c, (const int *)s->coupling_ang[w], (size_t)s->coupling_steps[w], &c->d_ang[w]))) return rc;
    }
  }
  c->n_psy = s->n_psy;
  for (int i = 0; i < s->n_psy; i++) {
    const vb200_psy_setup &p = s->psy[i];
    if (!p.ath || !p.octave || !p.bark || !p.tonecurves || !p.noiseoffset)
      return fail(VB200_EINVAL, "psy lookup tables missing");
    if (p.n != s->blocksizes[i >> 1] / 2) return fail(VB200_EINVAL, "psy n != blocksize/2");
    HostPsyFlow &f = c->hflow[i];
    build_psy_flow(f, p);
    PsyDev &d = c->dpsy[i];
    memset(&d, 0, sizeof(d));
    d.n = p.n; d.total = p.total_octave_lines; d.linesper = p.eighth_octave_lines;
    d.firstoc = p.firstoc; d.shiftoc = p.shiftoc;
    d.noisewindowfixed = p.noisewindowfixed;
    d.bark_first_extra = f.bark_first_extra; d.fixed_first_extra = f.fixed_first_extra;
    d.nruns = (int)f.run_lo.size(); d.ngrp = (int)f.grp.size() / 4; d.tail_lin0 = f.tail_lin0;
    d.ath_adjatt = p.ath_adjatt; d.ath_maxatt = p.ath_maxatt; d.tone_abs_limit = p.tone_abs_limit;
    d.noisemaxsupp = p.noisemaxsupp; d.max_curve_dB = p.max_curve_dB; d.m_val = p.m_val;
    for (int j = 0; j < 3; j++) d.tone_masteratt[j] = p.tone_masteratt[j];
    int rc;
    if ((rc = upload(c, p.ath, p.n, &d.ath))) return rc;
    if ((rc = upload(c, p.octave, p.n, &d.octave))) return rc;
    if ((rc = upload(c, p.bark, p.n, &d.bark))) return rc;
    if ((rc = upload(c, p.tonecurves, (size_t)VB200_P_BANDS * VB200_P_LEVELS * (VB200_EHMER_MAX + 2),
                     &d.tonecurves))) return rc;
    if ((rc = upload(c, p.noiseoffset, (size_t)VB200_P_NOISECURVES * p.n, &d.noiseoffset))) return rc;
    if ((rc = upload(c, p.noisecompand, (size_t)VB200_COMPAND_LEVELS, &d.noisecompand))) return rc;
    if ((1 << f.linesper_log2) != p.eighth_octave_lines)
      return fail(VB200_EIMPL, "eighth_octave_lines must be a power of two (lib/psy.h:108)");
    const int *dr = nullptr, *dg = nullptr, *dc = nullptr, *ds = nullptr;
    if ((rc = upload(c, f.runinfo.data(), f.runinfo.size(), &dr))) return rc;
    if ((rc = upload(c, f.grp.data(), f.grp.size(), &dg))) return rc;
    if ((rc = upload(c, f.cls_run.data(), f.cls_run.size(), &dc))) return rc;
    if ((rc = upload(c, f.slot_rng.data(), f.slot_rng.size(), &ds))) return rc;
    d.runinfo = reinterpret_cast<const int4 *>(dr);
    d.grps = reinterpret_cast<const int4 *>(dg);
    d.cls_run = reinterpret_cast<const int2 *>(dc);
    d.slot_rng = reinterpret_cast<const int2 *>(ds);
    d.linesper_log2 = f.linesper_log2;
    { const int *drr = nullptr, *dlg = nullptr;
      if ((rc = upload(c, f.runrec.data(), f.runrec.size(), &drr))) return rc;
      if ((rc = upload(c, f.long_grp.data(), f.long_grp.size(), &dlg))) return rc;
      d.runrec = reinterpret_cast<const int4 *>(drr); d.long_grp = dlg; d.nlong = (int)f.long_grp.size(); }
    { const short *db = nullptr; if ((rc = upload(c, f.bin_grp.data(), f.bin_grp.size(), &db))) return rc; d.bin_grp = db; }
    if (p.eighth_octave_lines > 32) return fail(VB200_EIMPL, "eighth_octave_lines > 32");
    if (p.total_octave_lines >= 2048) return fail(VB200_EIMPL, "total_octave_lines >= 2048");
    { const int *dco = nullptr; if ((rc = upload(c, f.cls_off.data(), f.cls_off.size(), &dco))) return rc; d.cls_off = dco; }
    d.max_cls_len = 0;
    for (size_t k = 0; k + 1 < f.cls_off.size(); k++) d.max_cls_len = std::max(d.max_cls_len, f.cls_off[k + 1] - f.cls_off[k]);
  }
  // floor 1 lookups: what floor1_look (lib/floor1.c:205-253) derives from the post list
  for (int w = 0; w < 2; w++) {
    Floor1Dev hf[VB200_MAX_SUBMAPS];
    memset(hf, 0, sizeof(hf));
    if (s->submaps[w] < 0 || s->submaps[w] > VB200_MAX_SUBMAPS) return fail(VB200_EINVAL, "submaps");
    for (int sm = 0; sm < VB200_MAX_SUBMAPS; sm++) {
      const vb200_floor1_setup &f = s->floor1[w][sm];
      Floor1Dev &d = hf[sm];
      if (f.posts == 0) continue;
      const int P = f.posts;
      if (P < 2 || P > VB200_VIF_POSIT + 2) return fail(VB200_EINVAL, "floor1 posts");
      if (f.mult < 1 || f.mult > 4) return fail(VB200_EINVAL, "floor1 mult");
      if (f.n < 1 || f.n > s->blocksizes[w] / 2 || f.postlist[0] != 0 || f.postlist[1] != f.n)
        return fail(VB200_EINVAL, "floor1 post list must start 0, n with n <= blocksize/2");
      for (int i = 0; i < P; i++) {
        if (f.postlist[i] < 0 || f.postlist[i] > f.n) return fail(VB200_EINVAL, "floor1 post out of range");
        for (int j = 0; j < i; j++)
          if (f.postlist[j] == f.postlist[i]) return fail(VB200_EINVAL, "floor1 posts must be distinct");
      }
      d.posts = P; d.n = f.n; d.mult = f.mult;
      d.maxover = f.maxover; d.maxunder = f.maxunder; d.maxerr = f.maxerr;
      d.twofitweight = f.twofitweight; d.twofitatten = f.twofitatten;
      d.int_thresh = f.maxover == (float)(int)f.maxover && f.maxunder == (float)(int)f.maxunder &&
                     fabsf(f.maxover) < 65536.f && fabsf(f.maxunder) < 65536.f;
      d.maxover_i = d.int_thresh ? (int)f.maxover : 0;
      d.maxunder_i = d.int_thresh ? (int)f.maxunder : 0;
      int order[VB200_VIF_POSIT + 2];
      for (int i = 0; i < P; i++) order[i] = i;
      std::sort(order, order + P, [&](int x, int y) { return f.postlist[x] < f.postlist[y]; });
      for (int i = 0; i < P; i++) {
        d.postlist[i] = (short)f.postlist[i];
        d.fwd[i] = (short)order[i];
        d.rev[order[i]] = (short)i;
        d.sorted[i] = (short)f.postlist[order[i]];
      }
      for (int i = 0; i < P - 2; i++) {                 // nearest already-coded posts on both sides
        int lo = 0, hi = 1, lx = 0, hx = f.n;
        const int cur = f.postlist[i + 2];
        for (int j = 0; j < i + 2; j++) {
          const int x = f.postlist[j];
          if (x > lx && x < cur) { lo = j; lx = x; }
          if (x < hx && x > cur) { hi = j; hx = x; }
        }
        d.lo[i] = (short)lo; d.hi[i] = (short)hi;
        d.prcp[i] = 1.f / (float)(hx - lx);             // render_point's divisor for post i+2 is static
      }
      {                                                 // dependency levels of the prediction passes
        int level[VB200_VIF_POSIT + 2], nl = 0, w = 0;
        level[0] = level[1] = -1;
        for (int i = 2; i < P; i++) {
          const int a = level[d.lo[i - 2]], b = level[d.hi[i - 2]];
          level[i] = (a > b ? a : b) + 1;
          if (level[i] + 1 > nl) nl = level[i] + 1;
        }
        d.nlevels = nl;
        for (int lv = 0; lv < nl; lv++) {
          d.lvl_start[lv] = (unsigned char)w;
          for (int i = 2; i < P; i++) if (level[i] == lv) d.lvl_order[w++] = (unsigned char)i;
        }
        d.lvl_start[nl] = (unsigned char)w;
      }
    }
    for (int k = 0; k < s->channels; k++) {
      const int sm = s->chmux[w][k];
      if (sm >= VB200_MAX_SUBMAPS) return fail(VB200_EINVAL, "chmux");
    }
    int rc;
    if ((rc = upload(c, hf, (size_t)VB200_MAX_SUBMAPS, &c->d_floor[w]))) return rc;
    if ((rc = upload(c, (const unsigned char *)s->chmux[w], (size_t)VB200_MAX_CHANNELS + 1, &c->d_chmux[w]))) return rc;
  }
  return 0;
}


extern "C" int vb200_ctx_create(const vb200_setup *s, int device, vb200_ctx **out) {
  if (!s || !out) return fail(VB200_EINVAL, "null argument");
  for (int w = 0; w < 2; w++)
    if (!pow2(s->blocksizes[w]) || s->blocksizes[w] < 64 || s->blocksizes[w] > 8192)
      return fail(VB200_EINVAL, "block sizes must be powers of two in [64,8192] (lib/info.c:227-228)");
  if (s->blocksizes[0] > s->blocksizes[1]) return fail(VB200_EINVAL, "blocksizes[0] > blocksizes[1]");
  if (s->n_psy != 0 && s->n_psy != 4) return fail(VB200_EIMPL, "n_psy must be 0 or 4");
  if (s->channels < 1 || s->channels > VB200_MAX_CHANNELS) return fail(VB200_EINVAL, "channels");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(VB200_EINVAL, "no such CUDA device");
  CU(cudaSetDevice(device));
  vb200_ctx *c = new vb200_ctx();
  c->device = device;
  c->setup = *s;
  const int rc = ctx_build(c, s, device);
  if (rc) { vb200_ctx_destroy(c); return rc; }      // streams, events and every table uploaded so far are released
  *out = c;
  return 0;
}

extern "C" void vb200_ctx_destroy(vb200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (void *p : c->owned) cudaFree(p);
  for (auto &b : c->scratch) if (b.p) cudaFree(b.p);
  for (auto &l : c->lane_buf) for (auto &b : l) if (b.p) cudaFree(b.p);
  for (auto &st : c->s_lane) if (st) cudaStreamDestroy(st);
  for (auto &st : c->s_split) if (st) cudaStreamDestroy(st);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (auto &e : c->ev_join) if (e) cudaEventDestroy(e);
  for (auto &b : c->enc_buf) if (b.p) cudaFree(b.p);
  for (auto &b : c->env_buf) if (b.p) cudaFree(b.p);
  for (auto &l : c->enc_lane) for (auto &b : l) if (b.p) cudaFree(b.p);
  for (auto &b : c->mgd_buf) if (b.p) cudaFree(b.p);
  for (auto &b : c->str_buf) if (b.p) cudaFree(b.p);
  for (auto &st : c->s_enc) if (st) cudaStreamDestroy(st);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  for (auto &e : c->ev_h2d) if (e) cudaEventDestroy(e);
  for (auto &e : c->ev_cmp) if (e) cudaEventDestroy(e);
  for (auto &e : c->ev_d2h) if (e) cudaEventDestroy(e);
  if (c->s_main) cudaStreamDestroy(c->s_main);
  for (auto &e : c->ev) if (e) cudaEventDestroy(e);
  if (c->ev_scratch) cudaEventDestroy(c->ev_scratch);
  delete c;
}

extern "C" int vb200_ctx_table(vb200_ctx *c, int W, int which, void *dst, int cap) {
  if (!c || W < 0 || W > 1 || !dst) return VB200_EINVAL;
  const HostXform &h = c->hx[W];
  const void *src = nullptr; size_t cnt = 0, el = 4;
  switch (which) {
    case 0: src = h.trig.data(); cnt = h.trig.size(); break;
    case 1: src = h.bitrev.data(); cnt = h.bitrev.size(); break;
    case 2: src = h.win.data(); cnt = h.win.size(); break;
    case 3: src = h.wa.data(); cnt = h.wa.size(); break;
    default: return VB200_EINVAL;
  }
  if ((size_t)cap < cnt) return VB200_EINVAL;
  memcpy(dst, src, cnt * el);
  return (int)cnt;
}

extern "C" uint64_t vb200_launch_count(vb200_ctx *c) { return c ? c->launches.load() : 0; }

extern "C" int vb200_set_profiling(vb200_ctx *c, int on) {
  if (!c) return fail(VB200_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  if (on) for (auto &e : c->ev) if (!e) CU(cudaEventCreate(&e));
  c->profiling = on != 0;
  return 0;
}

// dev aid: per-phase cycle sums of k_phaseA_psy2 accumulated while VB200_PHASE_TIMING is set
extern "C" int vb200_debug_phase_cycles(vb200_ctx *c, unsigned long long *out16, int reset) {
  if (!c || !out16) return fail(VB200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  void *p; int rc;
  if ((rc = ensure(c, 10, 16 * sizeof(unsigned long long), &p))) return rc;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out16, p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) CU(cudaMemset(p, 0, 16 * sizeof(unsigned long long)));
  return 0;
}

extern "C" int vb200_phaseA_kernel_ms(vb200_ctx *c, float *ms3) {
  if (!c || !ms3 || !c->ev[0]) return fail(VB200_EINVAL, "profiling not enabled");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ev[3]));
  for (int i = 0; i < 3; i++) CU(cudaEventElapsedTime(ms3 + i, c->ev[i], c->ev[i + 1]));
  return 0;
}

extern "C" int vb200_encode_dsp_kernel_ms(vb200_ctx *c, float *ms6) {
  if (!c || !ms6 || !c->ev[0]) return fail(VB200_EINVAL, "profiling not enabled");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ev[6]));
  for (int i = 0; i < 6; i++) CU(cudaEventElapsedTime(ms6 + i, c->ev[i], c->ev[i + 1]));
  return 0;
}

// ======================================================================== //
// kernels

// config 2: batched mdct_forward.  One CTA per vector (grid-stride), the vector
// staged in shared memory with 128-bit coalesced loads.
template <int NC>
__global__ void __launch_bounds__(256)
k_mdct_forward(XformDev X, int nvec, const float *__restrict__ in, float *__restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int N = NC ? NC : X.N, tid = threadIdx.x, nt = blockDim.x;
  // two input buffers: the NEXT vector of this CTA travels global -> shared with cp.async (LDGSTS, no registers)
  // while this one is transformed; the padded hand-over buffer keeps the bit-reverse gather conflict free
  float *sx[2] = {sm, sm + N};
  float *sw = sm + 2 * N, *pad = sm + 3 * N;
  auto stage = [&](float *dst, int v) {
    const float *src = in + (size_t)v * N;
    const unsigned d = smem_u32(dst);
    for (int i = tid; i < (N >> 2); i += nt) cp_async16(d + 16u * (unsigned)i, src + 4 * i);
    cp_async_commit();
  };
  int v = blockIdx.x, buf = 0;
  if (v < nvec) stage(sx[0], v);
  for (; v < nvec; v += gridDim.x, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();
    if (v + (int)gridDim.x < nvec) stage(sx[buf ^ 1], v + gridDim.x);
    dev_mdct_forward<NC>(X, sx[buf], sw, out + (size_t)v * (N >> 1), tid, nt, pad);
    __syncthreads();
  }
}

template <int NC>
__global__ void __launch_bounds__(256)
k_mdct_backward(XformDev X, int nvec, const float *__restrict__ in, float *__restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int N = NC ? NC : X.N, n2 = N >> 1, tid = threadIdx.x, nt = blockDim.x;
  float *sin_ = sm, *so = sm + n2;         // n2 coefficients, N outputs
  for (int v = blockIdx.x; v < nvec; v += gridDim.x) {
    const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)v * n2);
    for (int i = tid; i < (n2 >> 2); i += nt) reinterpret_cast<float4 *>(sin_)[i] = __ldg(src + i);
    __syncthreads();
    dev_mdct_backward<NC>(X, sin_, so, tid, nt);
    float4 *dst = reinterpret_cast<float4 *>(out + (size_t)v * N);
    for (int i = tid; i < (N >> 2); i += nt) dst[i] = reinterpret_cast<float4 *>(so)[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_apply_window(WinDev Wd, int W, int nvec, const int *__restrict__ lW, const int *__restrict__ nW,
               float *__restrict__ data) {
  const int N = Wd.N[W];
  for (int v = blockIdx.x; v < nvec; v += gridDim.x) {
    const int l = lW ? lW[v] : 0, r = nW ? nW[v] : 0;
    float *d = data + (size_t)v * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      bool z;
      const float g = dev_window_gain(Wd, W, l, r, i, z);
      d[i] = z ? 0.f : d[i] * g;
    }
  }
}

__global__ void __launch_bounds__(256)
k_drft_forward(XformDev X, int nvec, float *__restrict__ data) {
  extern __shared__ __align__(16) float sm[];
  const int N = X.N, tid = threadIdx.x, nt = blockDim.x;
  float *sa = sm, *sb = sm + fft_buf_floats(N);
  for (int v = blockIdx.x; v < nvec; v += gridDim.x) {
    float4 *g = reinterpret_cast<float4 *>(data + (size_t)v * N);
    for (int i = tid; i < (N >> 2); i += nt) reinterpret_cast<float4 *>(sa)[i] = g[i];
    __syncthreads();
    const float *r = dev_drft_forward<0>(X, sa, sb, tid, nt);
    float *gs = data + (size_t)v * N;
    for (int i = tid; i < N; i += nt) gs[i] = r[fft_idx(i)];
    __syncthreads();
  }
}

// ---- Phase A, kernel 1: per (block, channel) window + MDCT + FFT + log spectrum.
// Reads 4N bytes of PCM, writes mdct (2N), logfft (2N) and one local_ampmax.
// (first per-channel loop of mapping0_forward, lib/mapping0.c:254-360)
// where a row's N samples come from: block layout (fmt 0), or a contiguous per-stream buffer
// from which block k is the window starting at k*hop (lib/block.c:630-643), float planar or
// interleaved int16 (examples/encoder_example.c:196-201)
struct PcmSrc {
  const void *base;
  int fmt;                 // 0 blocks [row][N] f32, VB200_PCM_F32_PLANAR, VB200_PCM_S16_INTERLEAVED
  int bps, hop;
  long long stride;        // samples per channel per stream
  const int2 *blk_src;     // optional [blocks]: (stream, first sample) of every block instead of (blk / bps, k * hop)
};

// stream and first sample of block `blk`: equal-size runs (k * hop) or the planner's table
__device__ __forceinline__ void blk_origin(const PcmSrc &src, int blk, long long &st, long long &off) {
  if (src.blk_src) { const int2 o = __ldg(src.blk_src + blk); st = o.x; off = o.y; }
  else { const int s = blk / src.bps; st = s; off = (long long)(blk - s * src.bps) * src.hop; }
}

// the FFT ping-pong buffer is idle during the MDCT: it takes the padded fly output
__device__ __forceinline__ float *mdct_pad_buffer(float *sf) {
#ifdef VB200_NO_FLY_PAD
  (void)sf; return nullptr;
#else
  return sf;
#endif
}

#ifndef XF_MINB
#define XF_MINB 5
#endif
template <int NC>
__global__ void __launch_bounds__(256, XF_MINB)
k_phaseA_transform(XformDev X, WinDev Wd, int W, int ch, int nrows,
                   PcmSrc src, const vb200_block_desc *__restrict__ desc,
                   float *__restrict__ mdct, float *__restrict__ logfft, float *__restrict__ lmax) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_red[8];
  const int N = NC ? NC : X.N, n = N >> 1, tid = threadIdx.x, nt = blockDim.x;
  float *sx = sm, *sw = sm + fft_buf_floats(N), *sf = sw + N;   // sx and sf double as the padded FFT ping-pong buffers
  const float scale = 4.f / (float)N;
  const float scale_dB = add345(todB_dev(scale));
  // The raw samples of the NEXT row travel global -> shared with cp.async (16-byte LDGSTS, no registers)
  // while this row is transformed; the window is applied on the shared -> shared pass that follows.
  // mode 0: nothing staged (unaligned source or channel layout without a vector path): direct loads
  // mode 1: N floats staged;  mode 2: N stereo int16 frames staged (this row keeps its channel)
  float *sraw = sf + fft_buf_floats(N);
  auto stage = [&](int r) -> int {
    const int b2 = r / ch, c = r - b2 * ch;
    const char *g = nullptr;
    int mode = 0;
    if (src.fmt == VB200_PCM_S16_INTERLEAVED) {
      long long st, off; blk_origin(src, b2, st, off);
      g = reinterpret_cast<const char *>(reinterpret_cast<const short *>(src.base) + (st * src.stride + off) * ch);
      mode = ch == 2 ? 2 : 0;
    } else {
      const float *pf = reinterpret_cast<const float *>(src.base);
      if (src.fmt == VB200_PCM_F32_PLANAR) {
        long long st, off; blk_origin(src, b2, st, off);
        pf += (st * ch + c) * src.stride + off;
      } else {
        pf += (size_t)r * N;
      }
      g = reinterpret_cast<const char *>(pf);
      mode = 1;
    }
    if (reinterpret_cast<uintptr_t>(g) & 15) mode = 0;
    if (mode) {
      const unsigned d = smem_u32(sraw);
      for (int v = tid; v < (N >> 2); v += nt) cp_async16(d + 16u * v, g + 16 * (size_t)v);
    }
    cp_async_commit();
    return mode;
  };
  int mode = blockIdx.x < nrows ? stage(blockIdx.x) : 0;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int blk = row / ch;
    const int lW = desc[blk].lW, nW = desc[blk].nW;
    cp_async_wait_all();
    __syncthreads();
    if (mode == 1) {
      for (int v = tid; v < (N >> 2); v += nt)
        *reinterpret_cast<float4 *>(sx + 4 * v) = dev_window4(Wd, W, lW, nW, 4 * v, *reinterpret_cast<const float4 *>(sraw + 4 * v));
    } else if (mode == 2) {
      // stereo: four frames = one 128-bit word, this row keeps its channel's four samples
      const int sh = 16 * (row - blk * ch);
      for (int v = tid; v < (N >> 2); v += nt) {
        const int4 u = *reinterpret_cast<const int4 *>(sraw + 4 * v);
        const float4 x = make_float4((float)(short)((unsigned)u.x >> sh) / 32768.f, (float)(short)((unsigned)u.y >> sh) / 32768.f,
                                     (float)(short)((unsigned)u.z >> sh) / 32768.f, (float)(short)((unsigned)u.w >> sh) / 32768.f);
        *reinterpret_cast<float4 *>(sx + 4 * v) = dev_window4(Wd, W, lW, nW, 4 * v, x);
      }
    } else if (src.fmt == VB200_PCM_S16_INTERLEAVED) {
      const int c = row - blk * ch;
      long long st, off; blk_origin(src, blk, st, off);
      const short *p16 = reinterpret_cast<const short *>(src.base) + (st * src.stride + off) * ch + c;
      for (int i = tid; i < N; i += nt) {
        bool z;
        const float g = dev_window_gain(Wd, W, lW, nW, i, z);
        const float v = (float)__ldg(p16 + (long long)i * ch) / 32768.f;
        sx[i] = z ? 0.f : v * g;
      }
    } else {
      const float *pf = reinterpret_cast<const float *>(src.base);
      if (src.fmt == VB200_PCM_F32_PLANAR) {
        const int c = row - blk * ch;
        long long st, off; blk_origin(src, blk, st, off);
        pf += (st * ch + c) * src.stride + off;
      } else {
        pf += (size_t)row * N;
      }
      for (int i = tid; i < N; i += nt) {          // unaligned source: scalar loads
        bool z;
        const float g = dev_window_gain(Wd, W, lW, nW, i, z);
        sx[i] = z ? 0.f : __ldg(pf + i) * g;
      }
    }
    __syncthreads();
    mode = row + gridDim.x < nrows ? stage(row + gridDim.x) : 0;
    dev_mdct_forward<NC>(X, sx, sw, mdct + (size_t)row * n, tid, nt, mdct_pad_buffer(sf));
    const float *f = dev_drft_forward<NC>(X, sx, sf, tid, nt);
    // log spectrum + local maximum (lib/mapping0.c:310-345)
    float mx = -1e30f;
    float *lf = logfft + (size_t)row * n;
    for (int k = tid; k < n; k += nt) {
      float v;
      if (k == 0) {
        v = add345(scale_dB + todB_dev(f[fft_idx(0)]));
      } else {
        const float2 c = *reinterpret_cast<const float2 *>(f + fft_idx(2 * k - 1));   // aligned: shifted by one, padded
        const float re = c.x, im = c.y;
        const float t = re * re + im * im;
        v = add345(scale_dB + .5f * todB_dev(t));
      }
      lf[k] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
      float m = s_red[0];
      for (int w = 1; w < (nt >> 5); w++) m = fmaxf(m, s_red[w]);
      if (m > 0.f) m = 0.f;
      lmax[row] = m;
    }
    __syncthreads();
  }
}

// ---- ampmax: per-block global_ampmax.
// mode 0 (independent blocks): g[blk] = max(desc[blk].ampmax, locals)   (lib/mapping0.c:244,346)
// mode 1 (streams): the decay chain of vorbis_analysis_blockout (lib/block.c:626-628,
//   lib/psy.c:837-848), sequential per stream: one thread per stream.
__global__ void k_ampmax(int mode, int nstreams, int bps, int ch, const vb200_block_desc *__restrict__ desc,
                         const float *__restrict__ lmax, const float *__restrict__ amp0,
                         float secs_att, float *__restrict__ gmax) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nstreams) return;
  if (mode == 0) {
    float g = desc[s].ampmax;
    for (int c = 0; c < ch; c++) g = fmaxf(g, lmax[(size_t)s * ch + c]);
    gmax[s] = g;
    return;
  }
  float g = amp0 ? amp0[s] : -9999.f;
  float prev = g;
  for (int k = 0; k < bps; k++) {
    const size_t blk = (size_t)s * bps + k;
    if (prev > g) g = prev;
    g = g + secs_att;                      // amp += secs*ampmax_att_per_sec
    if (g < -9999.f) g = -9999.f;
    float o = g;
    for (int c = 0; c < ch; c++) o = fmaxf(o, lmax[blk * ch + c]);
    gmax[blk] = o;
    prev = o;
  }
}

// ---- Phase A, kernel 2: per (block, channel) logmdct, noise mask, tone mask, mix.
// (second per-channel loop of mapping0_forward, lib/mapping0.c:366-470)

// shared-memory carve-up for the psy kernels (floats)
struct PsySmem {
  float *logmdct, *noise, *scan, *fft;
  ToneSmem T;
};
__host__ __device__ inline size_t psy_smem_floats(int n, int total, int nruns) {
  const int tp = (total + 7) & ~7, rp = (nruns + 3) & ~3;
  size_t runs = 4 * (size_t)rp;                      // run_rec (int4), also hosts rec (tp shorts)
  if (runs < (size_t)tp / 2) runs = tp / 2;
  return (size_t)3 * n + 5 * (size_t)(n + 4) + 2 * (size_t)tp + tp / 2 + runs;
}
__device__ __forceinline__ PsySmem psy_carve(float *sm, int n, int total, int nruns) {
  const int tp = (total + 7) & ~7, rp = (nruns + 3) & ~3;
  PsySmem s;
  s.logmdct = sm; s.noise = s.logmdct + n; s.scan = s.noise + n;
  s.fft = s.scan + 5 * (n + 4);
  s.T.seed = s.fft + n;
  s.T.astk = s.T.seed + tp;
  s.T.pstk = reinterpret_cast<short *>(s.T.astk + tp);
  s.T.run_rec = reinterpret_cast<int4 *>(s.T.astk + tp + tp / 2);
  s.T.rec = reinterpret_cast<short *>(s.T.run_rec);
  (void)rp;
  return s;
}

#define PSY_THREADS 128
__global__ void __launch_bounds__(PSY_THREADS)
k_phaseA_psy(PsyDev P0, PsyDev P1, int ch, int nrows, PhaseA2Args A) {
  extern __shared__ __align__(16) float sm[];
  const int n = P0.n, ns = n + 4, tid = threadIdx.x, nt = PSY_THREADS;
  const int total = P0.total > P1.total ? P0.total : P1.total;
  const int nruns = P0.nruns > P1.nruns ? P0.nruns : P1.nruns;
  const PsySmem S = psy_carve(sm, n, total, nruns);
  const int warp = tid >> 5, lane = tid & 31;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int blk = row / ch;
    const PsyDev &P = A.desc[blk].blocktype ? P1 : P0;
    const float *gm = A.mdct_in + (size_t)row * n;
    const float *lf = A.logfft + (size_t)row * n;
    const float g = A.gmax[blk], lmax = A.lmax[row];
    for (int i = tid; i < n; i += nt) {
      const float l = add345(todB_dev(gm[i]));          // lib/mapping0.c:384-385
      S.logmdct[i] = l;
      A.logmdct[(size_t)row * n + i] = l;
      S.fft[i] = lf[i];
    }
    __syncthreads();
    // all warps: run peaks / curve choice, and the first pass' per-bin sum terms
    dev_tone_runs(P, S.fft, g, lmax, S.T, tid, nt);
    dev_noise_terms(n, S.logmdct, nullptr, 140.f, S.scan, ns, tid, nt);
    __syncthreads();
    if (!(A.dbg_skip & 4)) dev_tone_slots(P, S.T, tid, nt);
    __syncthreads();
    // warp 0: the sequential seed_chase + gather; warps 1-3: the noise mask (two sequential
    // prefix-sum passes on five lanes + regressions).  The two chains are independent.
    if (warp == 0) { if (!(A.dbg_skip & 1)) dev_tone_chase_gather(P, S.fft, lmax, S.T, lane); }
    else if (!(A.dbg_skip & 2)) dev_noisemask(P, S.logmdct, S.noise, S.scan, ns, tid - 32, nt - 32, 1, true);
    __syncthreads();
    const float *noff = P.noiseoffset + n;               // offset_select 1
    for (int i = tid; i < n; i += nt) {
      float m = gm[i];
      const float nz = S.noise[i], tn = S.fft[i];
      const float lm = dev_mix_bin(P, 1, nz, tn, __ldg(noff + i), S.logmdct[i], m);
      A.logmask[(size_t)row * n + i] = lm;
      A.mdct_out[(size_t)row * n + i] = m;
      if (A.tap_noise) A.tap_noise[(size_t)row * n + i] = nz;
      if (A.tap_tone) A.tap_tone[(size_t)row * n + i] = tn;
    }
    if (tid == 0 && (row % ch) == 0) A.ampmax_out[blk] = g;   // lib/mapping0.c:576
    __syncthreads();
  }
}

// ---- stage-isolated psy kernels (parity tests feed them the oracle's upstream vectors)
__global__ void __launch_bounds__(PSY_THREADS)
k_noisemask(PsyDev P, int nvec, const float *__restrict__ logmdct, float *__restrict__ noise) {
  extern __shared__ __align__(16) float sm[];
  const int n = P.n, ns = n + 4, tid = threadIdx.x, nt = PSY_THREADS;
  const PsySmem S = psy_carve(sm, n, P.total, P.nruns);
  for (int v = blockIdx.x; v < nvec; v += gridDim.x) {
    for (int i = tid; i < n; i += nt) S.logmdct[i] = logmdct[(size_t)v * n + i];
    __syncthreads();
    dev_noisemask(P, S.logmdct, S.noise, S.scan, ns, tid, nt, 0, false);
    for (int i = tid; i < n; i += nt) noise[(size_t)v * n + i] = S.noise[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PSY_THREADS)
k_tonemask(PsyDev P, int nvec, const float *__restrict__ logfft, const float *__restrict__ gmax,
           const float *__restrict__ lmax, float *__restrict__ tone) {
  extern __shared__ __align__(16) float sm[];
  const int n = P.n, tid = threadIdx.x, nt = PSY_THREADS;
  const PsySmem S = psy_carve(sm, n, P.total, P.nruns);
  for (int v = blockIdx.x; v < nvec; v += gridDim.x) {
    for (int i = tid; i < n; i += nt) S.fft[i] = logfft[(size_t)v * n + i];
    __syncthreads();
    dev_tone_runs(P, S.fft, gmax[v], lmax[v], S.T, tid, nt);
    __syncthreads();
    dev_tone_slots(P, S.T, tid, nt);
    __syncthreads();
    if (tid < 32) dev_tone_chase_gather(P, S.fft, lmax[v], S.T, tid);
    __syncthreads();
    for (int i = tid; i < n; i += nt) tone[(size_t)v * n + i] = S.fft[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_offset_and_mix(PsyDev P, int nvec, int sel, const float *__restrict__ noise,
                 const float *__restrict__ tone, float *__restrict__ mdct,
                 const float *__restrict__ logmdct, float *__restrict__ logmask) {
  const int n = P.n;
  const size_t tot = (size_t)nvec * n;
  const float *noff = P.noiseoffset + (size_t)sel * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % n);
    float m = mdct[e];
    logmask[e] = dev_mix_bin(P, sel, noise[e], tone[e], __ldg(noff + i), logmdct[e], m);
    if (sel == 1) mdct[e] = m;
  }
}

// ---- decode: mdct_backward (lib/mapping0.c:792-795) + the windowed overlap-add of
// vorbis_synthesis_blockin (lib/block.c:767-823).  One CTA walks one (stream, channel)
// block by block; the previous block's second half stays in shared memory, so the only
// HBM traffic is the spectra in (2N) and the finished samples out (2N per channel-block).
// finished-sample sinks: planar float, or interleaved int16 as examples/decoder_example.c:250-262
struct SinkF32 {
  float *p;
  __device__ __forceinline__ void put(int i, float v) const { p[i] = v; }
};
struct SinkS16 {
  short *p; int ch;
  __device__ __forceinline__ void put(int i, float v) const {
    int val = (int)floorf(v * 32767.f + .5f);
    if (val > 32767) val = 32767;
    if (val < -32768) val = -32768;
    p[(long long)i * ch] = (short)val;
  }
};

template <bool S16>
__global__ void __launch_bounds__(256)
k_synthesis(XformDev X0, XformDev X1, WinDev Wd, int ch, int nstreams, int nblk,
            const int *__restrict__ Wseq, const long long *__restrict__ coef_off,
            const float *__restrict__ coef, const long long *__restrict__ pcm_off,
            void *__restrict__ pcm_out, long long pcm_stride) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n0 = X0.N >> 1, n1 = X1.N >> 1;
  float *s_in = sm, *s_out = sm + n1, *s_prev = s_out + 2 * n1;
  const float *w0 = Wd.win[0], *w1 = Wd.win[1];
  const int off = n1 / 2 - n0 / 2;
  for (int task = blockIdx.x; task < nstreams * ch; task += gridDim.x) {
    const int st = task / ch, c = task - st * ch;
    int lW = 0;
    for (int k = 0; k < nblk; k++) {
      const int W = Wseq[(size_t)st * nblk + k];
      const XformDev &X = W ? X1 : X0;
      const int N = X.N, n2 = N >> 1;
      const float4 *src = reinterpret_cast<const float4 *>(coef + coef_off[(size_t)st * nblk + k] + (size_t)c * n2);
      for (int i = tid; i < (n2 >> 2); i += nt) reinterpret_cast<float4 *>(s_in)[i] = __ldg(src + i);
      __syncthreads();
      dev_mdct_backward<0>(X, s_in, s_out, tid, nt);
      if (k > 0) {
        const long long o = pcm_off[(size_t)st * nblk + k];
        typename std::conditional<S16, SinkS16, SinkF32>::type dst;
        if constexpr (S16) {
          dst.p = reinterpret_cast<short *>(pcm_out) + ((long long)st * pcm_stride + o) * ch + c;
          dst.ch = ch;
        } else {
          dst.p = reinterpret_cast<float *>(pcm_out) + ((size_t)st * ch + c) * pcm_stride + o;
        }
        const float *R = s_prev, *Lh = s_out;
        if (lW && W) {
          for (int i = tid; i < n1; i += nt) dst.put(i, R[i] * __ldg(w1 + n1 - i - 1) + Lh[i] * __ldg(w1 + i));
        } else if (lW && !W) {
          for (int i = tid; i < off; i += nt) dst.put(i, R[i]);
          for (int i = tid; i < n0; i += nt) dst.put(off + i, R[off + i] * __ldg(w0 + n0 - i - 1) + Lh[i] * __ldg(w0 + i));
        } else if (!lW && W) {
          for (int i = tid; i < n1 / 2 + n0 / 2; i += nt)
            dst.put(i, i < n0 ? R[i] * __ldg(w0 + n0 - i - 1) + Lh[off + i] * __ldg(w0 + i) : Lh[off + i]);
        } else {
          for (int i = tid; i < n0; i += nt) dst.put(i, R[i] * __ldg(w0 + n0 - i - 1) + Lh[i] * __ldg(w0 + i));
        }
      }
      __syncthreads();
      for (int i = tid; i < n2; i += nt) s_prev[i] = s_out[n2 + i];
      lW = W;
      __syncthreads();
    }
  }
}

// ---- decode: undo channel coupling, lib/mapping0.c:754-779 (square polar -> L/R), in place
__global__ void __launch_bounds__(256)
k_decouple(int n, int ch, int steps, const int *__restrict__ mag, const int *__restrict__ ang,
           long long total, float *__restrict__ res) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long blk = e / n;
    const int j = (int)(e - blk * n);
    float *base = res + blk * (long long)ch * n + j;
    for (int s = steps - 1; s >= 0; s--) {
      float *pM = base + (long long)mag[s] * n, *pA = base + (long long)ang[s] * n;
      const float m = *pM, a = *pA;
      if (m > 0.f) {
        if (a > 0.f) { *pM = m; *pA = m - a; }
        else         { *pA = m; *pM = m + a; }
      } else {
        if (a > 0.f) { *pM = m; *pA = m + a; }
        else         { *pA = m; *pM = m - a; }
      }
    }
  }
}

// ======================================================================== //
// launch helpers
static int threads_for(int N) {
  int div = 8;
  { const char *e = getenv("VB200_XF_DIV"); if (e && atoi(e) > 0) div = atoi(e); }
  int t = N / div;
  if (t < 64) t = 64;
  if (t > 256) t = 256;
  return t;
}

static int grid_for(vb200_ctx *c, int items, int ctas_per_sm) {
  // grid_div > 1: this launch shares the SMs with the kernels of a concurrent half-batch (encode split)
  int per_sm = ctas_per_sm / c->grid_div;
  if (per_sm < 1) per_sm = 1;
  long g = (long)c->sm_count * per_sm;
  if (g > items) g = items;
  if (g < 1) g = 1;
  return (int)g;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

#define CHECK_CTX(c) do { if (!(c)) return fail(VB200_EINVAL, "null context"); CU(cudaSetDevice((c)->device)); } while (0)
#define CHECK_W(W)   do { if ((W) < 0 || (W) > 1) return fail(VB200_EINVAL, "W must be 0 or 1"); } while (0)

static int post_launch(vb200_ctx *c, int n = 1) {
  c->launches += n;
  CU(cudaGetLastError());
  return 0;
}

// ======================================================================== //
// transforms
extern "C" int vb200_mdct_forward_dev(vb200_ctx *c, int W, int nvec, const float *d_in, float *d_out, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  const XformDev &X = c->dx[W];
  const size_t smem = sizeof(float) * (3 * (size_t)X.N + (X.N / 2 + X.N / 32 + X.N / 512 + 8));   // 2 x in, work, padded hand-over
  const int nt = threads_for(X.N);
  const int grid = grid_for(c, nvec, 8);
  int rc;
#define LAUNCH_MF(NN)                                                                   \
  do {                                                                                  \
    if ((rc = set_smem(k_mdct_forward<NN>, smem))) return rc;                           \
    k_mdct_forward<NN><<<grid, nt, smem, (cudaStream_t)stream>>>(X, nvec, d_in, d_out); \
  } while (0)
  switch (X.N) {
    case 256: LAUNCH_MF(256); break;
    case 512: LAUNCH_MF(512); break;
    case 1024: LAUNCH_MF(1024); break;
    case 2048: LAUNCH_MF(2048); break;
    default: LAUNCH_MF(0); break;
  }
#undef LAUNCH_MF
  return post_launch(c);
}

extern "C" int vb200_mdct_backward_dev(vb200_ctx *c, int W, int nvec, const float *d_in, float *d_out, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  const XformDev &X = c->dx[W];
  const size_t smem = sizeof(float) * (X.N + X.N / 2);
  const int nt = threads_for(X.N), grid = grid_for(c, nvec, 8);
  int rc;
#define LAUNCH_MB(NN)                                                                    \
  do {                                                                                   \
    if ((rc = set_smem(k_mdct_backward<NN>, smem))) return rc;                           \
    k_mdct_backward<NN><<<grid, nt, smem, (cudaStream_t)stream>>>(X, nvec, d_in, d_out); \
  } while (0)
  switch (X.N) {
    case 256: LAUNCH_MB(256); break;
    case 512: LAUNCH_MB(512); break;
    case 1024: LAUNCH_MB(1024); break;
    case 2048: LAUNCH_MB(2048); break;
    default: LAUNCH_MB(0); break;
  }
#undef LAUNCH_MB
  return post_launch(c);
}

// generic "copy in, run, copy out" for the host-buffer variants
struct HostIO {
  vb200_ctx *c;
  int slot = 0;
  int h2d(const void *src, size_t bytes, void **d) {
    int rc = ensure(c, slot++, bytes, d); if (rc) return rc;
    if (src) CU(cudaMemcpyAsync(*d, src, bytes, cudaMemcpyHostToDevice, c->s_main));
    return 0;
  }
  int d2h(void *dst, const void *d, size_t bytes) {
    CU(cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, c->s_main));
    return 0;
  }
  int sync() { CU(cudaStreamSynchronize(c->s_main)); return 0; }
};

extern "C" int vb200_mdct_forward(vb200_ctx *c, int W, int nvec, const float *in, float *out) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int N = c->dx[W].N;
  HostIO io{c};
  void *di, *dout; int rc;
  if ((rc = io.h2d(in, sizeof(float) * (size_t)nvec * N, &di))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(float) * (size_t)nvec * N / 2, &dout))) return rc;
  if ((rc = vb200_mdct_forward_dev(c, W, nvec, (const float *)di, (float *)dout, c->s_main))) return rc;
  if ((rc = io.d2h(out, dout, sizeof(float) * (size_t)nvec * N / 2))) return rc;
  return io.sync();
}

extern "C" int vb200_mdct_backward(vb200_ctx *c, int W, int nvec, const float *in, float *out) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int N = c->dx[W].N;
  HostIO io{c};
  void *di, *dout; int rc;
  if ((rc = io.h2d(in, sizeof(float) * (size_t)nvec * N / 2, &di))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(float) * (size_t)nvec * N, &dout))) return rc;
  if ((rc = vb200_mdct_backward_dev(c, W, nvec, (const float *)di, (float *)dout, c->s_main))) return rc;
  if ((rc = io.d2h(out, dout, sizeof(float) * (size_t)nvec * N))) return rc;
  return io.sync();
}

extern "C" int vb200_apply_window(vb200_ctx *c, int W, int nvec, const int32_t *lW, const int32_t *nW, float *data) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int N = c->dx[W].N;
  HostIO io{c};
  void *dd, *dl = nullptr, *dn = nullptr; int rc;
  if ((rc = io.h2d(data, sizeof(float) * (size_t)nvec * N, &dd))) return rc;
  if (lW && (rc = io.h2d(lW, sizeof(int32_t) * nvec, &dl))) return rc;
  if (nW && (rc = io.h2d(nW, sizeof(int32_t) * nvec, &dn))) return rc;
  k_apply_window<<<grid_for(c, nvec, 8), 256, 0, c->s_main>>>(c->dwin, W, nvec, (const int *)dl, (const int *)dn, (float *)dd);
  if ((rc = post_launch(c))) return rc;
  if ((rc = io.d2h(data, dd, sizeof(float) * (size_t)nvec * N))) return rc;
  return io.sync();
}

extern "C" int vb200_drft_forward(vb200_ctx *c, int W, int nvec, float *data) {
  CHECK_CTX(c); CHECK_W(W);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const XformDev &X = c->dx[W];
  HostIO io{c};
  void *dd; int rc;
  if ((rc = io.h2d(data, sizeof(float) * (size_t)nvec * X.N, &dd))) return rc;
  const size_t smem = sizeof(float) * 2 * fft_buf_floats(X.N);
  if ((rc = set_smem(k_drft_forward, smem))) return rc;
  k_drft_forward<<<grid_for(c, nvec, 8), threads_for(X.N), smem, c->s_main>>>(X, nvec, (float *)dd);
  if ((rc = post_launch(c))) return rc;
  if ((rc = io.d2h(data, dd, sizeof(float) * (size_t)nvec * X.N))) return rc;
  return io.sync();
}

// ======================================================================== //
// stage-isolated psy entry points
static size_t psy2_smem(const PsyDev &a, const PsyDev &b) {
  const int total = a.total > b.total ? a.total : b.total;
  const int nruns = a.nruns > b.nruns ? a.nruns : b.nruns;
  return sizeof(float) * psy_smem_floats(a.n, total, nruns);
}

#define CHECK_LOOK(c, look) do { if ((look) < 0 || (look) >= (c)->n_psy) return fail(VB200_EINVAL, "no such psy look"); } while (0)

extern "C" int vb200_noisemask(vb200_ctx *c, int look, int nvec, const float *logmdct, float *noise) {
  CHECK_CTX(c); CHECK_LOOK(c, look);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const PsyDev &P = c->dpsy[look];
  HostIO io{c};
  void *di, *dout; int rc;
  const size_t bytes = sizeof(float) * (size_t)nvec * P.n;
  if ((rc = io.h2d(logmdct, bytes, &di))) return rc;
  if ((rc = io.h2d(nullptr, bytes, &dout))) return rc;
  const size_t smem = psy2_smem(P, P);
  if ((rc = set_smem(k_noisemask, smem))) return rc;
  k_noisemask<<<grid_for(c, nvec, 4), PSY_THREADS, smem, c->s_main>>>(P, nvec, (const float *)di, (float *)dout);
  if ((rc = post_launch(c))) return rc;
  if ((rc = io.d2h(noise, dout, bytes))) return rc;
  return io.sync();
}

extern "C" int vb200_tonemask(vb200_ctx *c, int look, int nvec, const float *logfft,
                              const float *gmax, const float *lmax, float *tone) {
  CHECK_CTX(c); CHECK_LOOK(c, look);
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const PsyDev &P = c->dpsy[look];
  HostIO io{c};
  void *di, *dg, *dl, *dout; int rc;
  const size_t bytes = sizeof(float) * (size_t)nvec * P.n;
  if ((rc = io.h2d(logfft, bytes, &di))) return rc;
  if ((rc = io.h2d(gmax, sizeof(float) * nvec, &dg))) return rc;
  if ((rc = io.h2d(lmax, sizeof(float) * nvec, &dl))) return rc;
  if ((rc = io.h2d(nullptr, bytes, &dout))) return rc;
  const size_t smem = psy2_smem(P, P);
  if ((rc = set_smem(k_tonemask, smem))) return rc;
  k_tonemask<<<grid_for(c, nvec, 4), PSY_THREADS, smem, c->s_main>>>(P, nvec, (const float *)di, (const float *)dg,
                                                             (const float *)dl, (float *)dout);
  if ((rc = post_launch(c))) return rc;
  if ((rc = io.d2h(tone, dout, bytes))) return rc;
  return io.sync();
}

extern "C" int vb200_offset_and_mix(vb200_ctx *c, int look, int nvec, int sel, const float *noise,
                                    const float *tone, float *mdct, const float *logmdct, float *logmask) {
  CHECK_CTX(c); CHECK_LOOK(c, look);
  if (sel < 0 || sel > 2) return fail(VB200_EINVAL, "offset_select");
  if (nvec <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const PsyDev &P = c->dpsy[look];
  HostIO io{c};
  void *dn, *dt, *dm, *dl, *dk; int rc;
  const size_t bytes = sizeof(float) * (size_t)nvec * P.n;
  if ((rc = io.h2d(noise, bytes, &dn))) return rc;
  if ((rc = io.h2d(tone, bytes, &dt))) return rc;
  if ((rc = io.h2d(mdct, bytes, &dm))) return rc;
  if ((rc = io.h2d(logmdct, bytes, &dl))) return rc;
  if ((rc = io.h2d(nullptr, bytes, &dk))) return rc;
  k_offset_and_mix<<<grid_for(c, (int)(((size_t)nvec * P.n + 255) / 256), 8), 256, 0, c->s_main>>>(
      P, nvec, sel, (const float *)dn, (const float *)dt, (float *)dm, (const float *)dl, (float *)dk);
  if ((rc = post_launch(c))) return rc;
  if ((rc = io.d2h(logmask, dk, bytes))) return rc;
  if ((rc = io.d2h(mdct, dm, bytes))) return rc;
  return io.sync();
}

// ======================================================================== //
// Phase A
// stage 1: window + MDCT + FFT + log spectrum of `nblocks` blocks of size W
static int phaseA_transform_launch(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *io, const PcmSrc &psrc,
                                   cudaStream_t st, float *d_logfft, float *d_lmax) {
  const XformDev &X = c->dx[W];
  const int ch = c->setup.channels, N = X.N;
  const int rows = nblocks * ch;
  {
    const size_t smem = sizeof(float) * (2 * N + 2 * fft_buf_floats(N));   // sx, sw, sf + the cp.async staging row
    int rc;
    float *mdct_raw = io->tap_mdct_raw ? io->tap_mdct_raw : io->mdct;
    const int grid = grid_for(c, rows, 8), nt = threads_for(N);
#define LAUNCH_XF(NN)                                                                        \
    do {                                                                                     \
      if ((rc = set_smem(k_phaseA_transform<NN>, smem))) return rc;                          \
      k_phaseA_transform<NN><<<grid, nt, smem, st>>>(X, c->dwin, W, ch, rows, psrc, io->desc,    \
                                                     mdct_raw, d_logfft, d_lmax);            \
    } while (0)
    switch (N) {
      case 256: LAUNCH_XF(256); break;
      case 512: LAUNCH_XF(512); break;
      case 1024: LAUNCH_XF(1024); break;
      case 2048: LAUNCH_XF(2048); break;
      default: LAUNCH_XF(0); break;
    }
#undef LAUNCH_XF
    rc = post_launch(c); if (rc) return rc;
  }
  return 0;
}

// stage 3: noise / tone masks + mix of `nblocks` blocks of size W, given every block's global ampmax
static int phaseA_psy_launch(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *io, cudaStream_t st,
                             float *d_logfft, float *d_lmax, float *d_gmax) {
  const XformDev &X = c->dx[W];
  const int ch = c->setup.channels, N = X.N;
  const int rows = nblocks * ch;
  int rc;
  {
    const PsyDev &P0 = c->dpsy[2 * W], &P1 = c->dpsy[2 * W + 1];
    const size_t smem = psy2_smem(P0, P1);
    int rc = set_smem(k_phaseA_psy, smem); if (rc) return rc;
    {
      // leave the rest of the 256 KB unified array to L1: the static psy tables (~60 KB per look)
      // are read through it on every row
      int &tuned = c->psy_carveout_ctas;
      const char *e = getenv("VB200_PSY_CTAS");
      const int ctas = e ? atoi(e) : c->psy_ctas_per_sm;
      if (tuned != ctas) {
        int pct = (int)((ctas * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        CU(cudaFuncSetAttribute(k_phaseA_psy, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        tuned = ctas;
      }
      c->psy_ctas_per_sm = ctas;
    }
    PhaseA2Args A;
    A.mdct_in = io->tap_mdct_raw ? io->tap_mdct_raw : io->mdct;
    A.logfft = d_logfft; A.lmax = d_lmax; A.gmax = d_gmax; A.desc = io->desc;
    A.mdct_out = io->mdct; A.logmdct = io->logmdct; A.logmask = io->logmask; A.ampmax_out = io->ampmax_out;
    A.tap_noise = io->tap_noise; A.tap_tone = io->tap_tone;
    { const char *e = getenv("VB200_DEBUG_SKIP"); A.dbg_skip = e ? atoi(e) : 0; }
    A.dbg_cycles = nullptr;
    if (getenv("VB200_PHASE_TIMING")) {
      void *p; if ((rc = ensure(c, 10, 16 * sizeof(unsigned long long), &p))) return rc;
      A.dbg_cycles = (unsigned long long *)p;
    }
    const int n = N / 2;
    const char *ev = getenv("VB200_PSY_V1");
    const bool v2ok = !(ev && atoi(ev)) && (n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048);
    const char *ev2 = getenv("VB200_PSY_V2");
    const bool v3ok = v2ok && !(ev2 && atoi(ev2)) && P0.linesper == P1.linesper &&
                      psy3_supported(n, P0.total > P1.total ? P0.total : P1.total, P0.linesper);
    if (v3ok) {
      const int total = P0.total > P1.total ? P0.total : P1.total;
      const int nruns = P0.nruns > P1.nruns ? P0.nruns : P1.nruns;
      const int ngrp = P0.ngrp > P1.ngrp ? P0.ngrp : P1.ngrp;
      int R = 1;                                     // rows per CTA sharing one scan warp (2: fewer instructions, same time)
      { const char *e = getenv("VB200_PSY_ROWS"); if (e) R = atoi(e) == 1 ? 1 : 2; }
      const size_t row_bytes = sizeof(float) * ((psy3_floats(n, total, nruns, ngrp) + 3) & ~(size_t)3);
      const size_t smem3 = row_bytes * R;
      int ctas = (int)((227 * 1024) / (smem3 + 1024));
      if (ctas > PSY3_MINB / R) ctas = PSY3_MINB / R;
      if (ctas < 1) ctas = 1;
      { const char *e = getenv("VB200_PSY_CTAS"); if (e) ctas = atoi(e); }
      const bool dbg3 = A.dbg_cycles || A.tap_noise || A.tap_tone;   // clock marks / taps: the debug instance
#define LAUNCH_PSY3D(KK, RR, DD)                                                                   \
      do {                                                                                         \
        if ((rc = set_smem(k_phaseA_psy3<KK, RR, DD>, smem3))) return rc;                          \
        k_phaseA_psy3<KK, RR, DD><<<grid_for(c, (rows + RR - 1) / RR, ctas), PSY3_THREADS * RR, smem3, st>>>(P0, P1, ch, rows, A); \
      } while (0)
#define LAUNCH_PSY3R(KK, RR) do { if (dbg3) LAUNCH_PSY3D(KK, RR, true); else LAUNCH_PSY3D(KK, RR, false); } while (0)
#define LAUNCH_PSY4(KK)                                                                            \
      do {                                                                                         \
        if ((rc = set_smem(k_phaseA_psy4<KK>, row_bytes))) return rc;                              \
        k_phaseA_psy4<KK><<<grid_for(c, rows, PSY3_MINB), PSY3_THREADS, row_bytes, st>>>(P0, P1, ch, rows, A); \
      } while (0)
      // k_phaseA_psy4 (regressions run behind the scans) is an experiment that measured slower than psy3: opt-in
      static const bool psy_v4 = []() { const char *e = getenv("VB200_PSY_V4"); return e && atoi(e); }();
#define LAUNCH_PSY3(KK) do { if (psy_v4) LAUNCH_PSY4(KK); else if (R == 1) LAUNCH_PSY3R(KK, 1); else LAUNCH_PSY3R(KK, 2); } while (0)
      switch (n / 128) {
        case 1: LAUNCH_PSY3(1); break;
        case 2: LAUNCH_PSY3(2); break;
        case 4: LAUNCH_PSY3(4); break;
        case 8: LAUNCH_PSY3(8); break;
        default: LAUNCH_PSY3(16); break;
      }
#undef LAUNCH_PSY3R
#undef LAUNCH_PSY3D
#undef LAUNCH_PSY4
#undef LAUNCH_PSY3
    } else if (v2ok) {
      const int total = P0.total > P1.total ? P0.total : P1.total;
      const int nruns = P0.nruns > P1.nruns ? P0.nruns : P1.nruns;
      const int ngrp = P0.ngrp > P1.ngrp ? P0.ngrp : P1.ngrp;
      const size_t smem2 = sizeof(float) * psy2_floats(n, total, nruns, ngrp);
      int ctas = (int)((227 * 1024) / (smem2 + 1024));
      if (ctas > PSY2_MINB) ctas = PSY2_MINB;
      if (ctas < 1) ctas = 1;
      { const char *e = getenv("VB200_PSY_CTAS"); if (e) ctas = atoi(e); }
#define LAUNCH_PSY2(KK)                                                                            \
      do {                                                                                         \
        if ((rc = set_smem(k_phaseA_psy2<KK>, smem2))) return rc;                                  \
        k_phaseA_psy2<KK><<<grid_for(c, rows, ctas), PSY2_THREADS, smem2, st>>>(P0, P1, ch, rows, A); \
      } while (0)
      switch (n / 128) {
        case 1: LAUNCH_PSY2(1); break;
        case 2: LAUNCH_PSY2(2); break;
        case 4: LAUNCH_PSY2(4); break;
        case 8: LAUNCH_PSY2(8); break;
        default: LAUNCH_PSY2(16); break;
      }
#undef LAUNCH_PSY2
    } else {
      k_phaseA_psy<<<grid_for(c, rows, c->psy_ctas_per_sm), PSY_THREADS, smem, st>>>(P0, P1, ch, rows, A);
    }
    rc = post_launch(c); if (rc) return rc;
  }
  return 0;
}

static int phaseA_launch(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *io,
                         int nstreams, int bps, const float *d_amp0, cudaStream_t st,
                         float *d_logfft, float *d_lmax, float *d_gmax, const PcmSrc *pcmsrc = nullptr) {
  PcmSrc psrc;
  if (pcmsrc) psrc = *pcmsrc;
  else { psrc.base = io->pcm; psrc.fmt = 0; psrc.bps = 1; psrc.hop = 0; psrc.stride = 0; psrc.blk_src = nullptr; }
  const XformDev &X = c->dx[W];
  const int ch = c->setup.channels, N = X.N;
  int rc;
  if (c->profiling) CU(cudaEventRecord(c->ev[0], st));
  if ((rc = phaseA_transform_launch(c, W, nblocks, io, psrc, st, d_logfft, d_lmax))) return rc;
  if (c->profiling) CU(cudaEventRecord(c->ev[1], st));
  {
    const int n = N / 2;
    const float secs = (float)n / (float)c->setup.rate;           // lib/psy.c:843
    const float secs_att = secs * c->setup.ampmax_att_per_sec;
    if (nstreams > 0)
      k_ampmax<<<(nstreams + 127) / 128, 128, 0, st>>>(1, nstreams, bps, ch, io->desc, d_lmax, d_amp0, secs_att, d_gmax);
    else
      k_ampmax<<<(nblocks + 127) / 128, 128, 0, st>>>(0, nblocks, 0, ch, io->desc, d_lmax, nullptr, secs_att, d_gmax);
    int rc = post_launch(c); if (rc) return rc;
  }
  if (c->profiling) CU(cudaEventRecord(c->ev[2], st));
  if ((rc = phaseA_psy_launch(c, W, nblocks, io, st, d_logfft, d_lmax, d_gmax))) return rc;
  if (c->profiling) CU(cudaEventRecord(c->ev[3], st));
  return 0;
}

static int ensure_buf(DevBuf &b, size_t bytes, void **out) {
  if (b.cap < bytes) {
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
  }
  *out = b.p;
  return 0;
}

static int phaseA_dev_common(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *io,
                             int nstreams, int bps, const float *d_amp0, void *stream,
                             const PcmSrc *pcmsrc = nullptr) {
  CHECK_CTX(c); CHECK_W(W);
  if (c->n_psy != 4) return fail(VB200_EIMPL, "context has no psy lookups");
  if (!io || (!io->pcm && !pcmsrc) || !io->desc || !io->mdct || !io->logmdct || !io->logmask || !io->ampmax_out)
    return fail(VB200_EINVAL, "phase A io pointers");
  if (nblocks <= 0) return 0;
  const int ch = c->setup.channels, n = c->dx[W].N / 2;
  void *d_logfft = io->tap_logfft, *d_lmax, *d_gmax; int rc;
  if (!d_logfft && (rc = ensure(c, 13, sizeof(float) * (size_t)nblocks * ch * n, &d_logfft))) return rc;
  if ((rc = ensure(c, 14, sizeof(float) * (size_t)nblocks * ch, &d_lmax))) return rc;
  if ((rc = ensure(c, 15, sizeof(float) * (size_t)nblocks, &d_gmax))) return rc;
  if ((rc = scratch_begin(c, (cudaStream_t)stream))) return rc;
  if ((rc = phaseA_launch(c, W, nblocks, io, nstreams, bps, d_amp0, (cudaStream_t)stream,
                          (float *)d_logfft, (float *)d_lmax, (float *)d_gmax, pcmsrc))) return rc;
  return scratch_end(c, (cudaStream_t)stream);
}

extern "C" int vb200_analysis_phaseA_pcmstream_dev(vb200_ctx *c, int W, int nstreams, int bps,
                                                   const void *d_pcm, int fmt, int64_t stream_stride, int hop,
                                                   const vb200_phaseA_io *io, const float *d_amp0, void *stream) {
  if (nstreams <= 0 || bps <= 0) return fail(VB200_EINVAL, "nstreams/blocks_per_stream");
  if (!d_pcm) return fail(VB200_EINVAL, "null pcm");
  if (fmt != VB200_PCM_F32_PLANAR && fmt != VB200_PCM_S16_INTERLEAVED) return fail(VB200_EINVAL, "pcm format");
  if (hop <= 0 || stream_stride <= 0) return fail(VB200_EINVAL, "hop/stream_stride");
  if (fmt == VB200_PCM_F32_PLANAR && ((hop & 3) || (stream_stride & 3)))
    return fail(VB200_EINVAL, "hop and stream_stride must be multiples of 4 for float PCM");
  if (!c || W < 0 || W > 1) return fail(VB200_EINVAL, "ctx/W");
  if ((int64_t)(bps - 1) * hop + c->dx[W].N > stream_stride) return fail(VB200_EINVAL, "blocks exceed the stream buffer");
  PcmSrc ps; ps.base = d_pcm; ps.fmt = fmt; ps.bps = bps; ps.hop = hop; ps.stride = stream_stride; ps.blk_src = nullptr;
  return phaseA_dev_common(c, W, nstreams * bps, io, nstreams, bps, d_amp0, stream, &ps);
}

extern "C" int vb200_analysis_phaseA_dev(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *io, void *stream) {
  return phaseA_dev_common(c, W, nblocks, io, 0, 0, nullptr, stream);
}

extern "C" int vb200_analysis_phaseA_streams_dev(vb200_ctx *c, int W, int nstreams, int bps,
                                                 const vb200_phaseA_io *io, const float *d_amp0, void *stream) {
  if (nstreams <= 0 || bps <= 0) return fail(VB200_EINVAL, "nstreams/blocks_per_stream");
  return phaseA_dev_common(c, W, nstreams * bps, io, nstreams, bps, d_amp0, stream);
}

// Host-buffer Phase A.  Without taps the batch is cut into chunks that alternate between two
// lanes (stream + device buffers each): while one lane computes, the other lane's H2D / D2H
// copies run on the copy engines.  Pinned host memory is needed for the copies to be truly
// asynchronous (pageable memory still works, staged by the driver).
static int phaseA_host_pipelined(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *h) {
  const int ch = c->setup.channels, N = c->dx[W].N, n = N / 2;
  int chunk = 4096;
  { const char *e = getenv("VB200_CHUNK_BLOCKS"); if (e && atoi(e) > 0) chunk = atoi(e); }
  if (chunk > nblocks) chunk = nblocks;
  const size_t crow = (size_t)chunk * ch;
  int rc;
  for (int L = 0; L < 2; L++) {
    void *p;
    const size_t sz[9] = {sizeof(float) * crow * N, sizeof(vb200_block_desc) * (size_t)chunk,
                          sizeof(float) * crow * n, sizeof(float) * crow * n, sizeof(float) * crow * n,
                          sizeof(float) * (size_t)chunk, sizeof(float) * crow * n, sizeof(float) * crow,
                          sizeof(float) * (size_t)chunk};
    for (int k = 0; k < 9; k++) if ((rc = ensure_buf(c->lane_buf[L][k], sz[k], &p))) return rc;
  }
  for (int b0 = 0, it = 0; b0 < nblocks; b0 += chunk, it++) {
    const int L = it & 1, nb = nblocks - b0 < chunk ? nblocks - b0 : chunk;
    const size_t rows = (size_t)nb * ch, r0 = (size_t)b0 * ch;
    cudaStream_t st = c->s_lane[L];
    DevBuf *B = c->lane_buf[L];
    CU(cudaMemcpyAsync(B[0].p, h->pcm + r0 * N, sizeof(float) * rows * N, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[1].p, h->desc + b0, sizeof(vb200_block_desc) * nb, cudaMemcpyHostToDevice, st));
    vb200_phaseA_io d;
    memset(&d, 0, sizeof(d));
    d.pcm = (const float *)B[0].p; d.desc = (const vb200_block_desc *)B[1].p;
    d.mdct = (float *)B[2].p; d.logmdct = (float *)B[3].p; d.logmask = (float *)B[4].p;
    d.ampmax_out = (float *)B[5].p;
    if ((rc = phaseA_launch(c, W, nb, &d, 0, 0, nullptr, st, (float *)B[6].p, (float *)B[7].p, (float *)B[8].p)))
      return rc;
    CU(cudaMemcpyAsync(h->mdct + r0 * n, d.mdct, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->logmdct + r0 * n, d.logmdct, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->logmask + r0 * n, d.logmask, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->ampmax_out + b0, d.ampmax_out, sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(c->s_lane[0]));
  CU(cudaStreamSynchronize(c->s_lane[1]));
  return 0;
}

extern "C" int vb200_analysis_phaseA(vb200_ctx *c, int W, int nblocks, const vb200_phaseA_io *h) {
  CHECK_CTX(c); CHECK_W(W);
  if (!h) return fail(VB200_EINVAL, "null io");
  if (c->n_psy != 4) return fail(VB200_EIMPL, "context has no psy lookups");
  if (!h->pcm || !h->desc || !h->mdct || !h->logmdct || !h->logmask || !h->ampmax_out)
    return fail(VB200_EINVAL, "phase A io pointers");
  if (nblocks <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!h->tap_noise && !h->tap_tone && !h->tap_logfft && !h->tap_mdct_raw && !c->profiling)
    return phaseA_host_pipelined(c, W, nblocks, h);
  const int ch = c->setup.channels, N = c->dx[W].N, n = N / 2;
  const size_t rows = (size_t)nblocks * ch;
  HostIO io{c};
  vb200_phaseA_io d;
  memset(&d, 0, sizeof(d));
  void *p; int rc;
  if ((rc = io.h2d(h->pcm, sizeof(float) * rows * N, &p))) return rc; d.pcm = (const float *)p;
  if ((rc = io.h2d(h->desc, sizeof(vb200_block_desc) * nblocks, &p))) return rc; d.desc = (const vb200_block_desc *)p;
  if ((rc = io.h2d(nullptr, sizeof(float) * rows * n, &p))) return rc; d.mdct = (float *)p;
  if ((rc = io.h2d(nullptr, sizeof(float) * rows * n, &p))) return rc; d.logmdct = (float *)p;
  if ((rc = io.h2d(nullptr, sizeof(float) * rows * n, &p))) return rc; d.logmask = (float *)p;
  if ((rc = io.h2d(nullptr, sizeof(float) * nblocks, &p))) return rc; d.ampmax_out = (float *)p;
  if (h->tap_noise) { if ((rc = io.h2d(nullptr, sizeof(float) * rows * n, &p))) return rc; d.tap_noise = (float *)p; }
  if (h->tap_tone) { if ((rc = io.h2d(nullptr, sizeof(float) * rows * n, &p))) return rc; d.tap_tone = (float *)p; }
  if (h->tap_logfft) { if ((rc = ensure(c, 11, sizeof(float) * rows * n, &p))) return rc; d.tap_logfft = (float *)p; }
  if (h->tap_mdct_raw) { if ((rc = ensure(c, 12, sizeof(float) * rows * n, &p))) return rc; d.tap_mdct_raw = (float *)p; }
  if ((rc = phaseA_dev_common(c, W, nblocks, &d, 0, 0, nullptr, c->s_main))) return rc;
  if ((rc = io.d2h(h->mdct, d.mdct, sizeof(float) * rows * n))) return rc;
  if ((rc = io.d2h(h->logmdct, d.logmdct, sizeof(float) * rows * n))) return rc;
  if ((rc = io.d2h(h->logmask, d.logmask, sizeof(float) * rows * n))) return rc;
  if ((rc = io.d2h(h->ampmax_out, d.ampmax_out, sizeof(float) * nblocks))) return rc;
  if (h->tap_noise && (rc = io.d2h(h->tap_noise, d.tap_noise, sizeof(float) * rows * n))) return rc;
  if (h->tap_tone && (rc = io.d2h(h->tap_tone, d.tap_tone, sizeof(float) * rows * n))) return rc;
  if (h->tap_logfft && (rc = io.d2h(h->tap_logfft, d.tap_logfft, sizeof(float) * rows * n))) return rc;
  if (h->tap_mdct_raw && (rc = io.d2h(h->tap_mdct_raw, d.tap_mdct_raw, sizeof(float) * rows * n))) return rc;
  return io.sync();
}

// ======================================================================== //
// Phase B
static int cqn_setup(vb200_ctx *c, int W, int blocktype, int blobno, CqnDev *Q) {
  if (c->n_psy != 4) return fail(VB200_EIMPL, "context has no psy lookups");
  if (blocktype < 0 || blocktype > 1) return fail(VB200_EINVAL, "blocktype");
  if (blobno < 0 || blobno >= VB200_PACKETBLOBS) return fail(VB200_EINVAL, "blobno");
  const vb200_psy_setup &p = c->setup.psy[blocktype + 2 * W];
  static const double thr[] = {0.0, .5, 1.0, 1.5, 2.5, 4.5, 8.5, 16.5, 9e10};       // lib/psy.c:32
  static const double thr_limited[] = {0.0, .5, 1.0, 1.5, 2.0, 2.5, 4.5, 8.5, 9e10}; // lib/psy.c:33
  const int pre = c->setup.coupling_prepointamp[blobno], post = c->setup.coupling_postpointamp[blobno];
  if (pre < 0 || pre > 8 || post < 0 || post > 8) return fail(VB200_EINVAL, "pointamp index");
  Q->n = p.n; Q->ch = c->setup.channels;
  Q->partition = p.normal_p ? p.normal_partition : 16;
  if (Q->partition != 8 && Q->partition != 16 && Q->partition != 32)
    return fail(VB200_EIMPL, "normal_partition must be 8, 16 or 32");
  Q->limit = c->setup.coupling_pointlimit[p.blockflag][blobno];
  Q->sliding_lowpass = c->setup.sliding_lowpass[W][blobno];
  Q->steps = c->setup.coupling_steps[W];
  Q->normal_p = p.normal_p; Q->normal_start = p.normal_start; Q->normal_thresh = p.normal_thresh;
  Q->prepoint = (float)thr[pre];
  Q->postpoint = (float)(p.n > 1000 ? thr_limited[post] : thr[post]);
  Q->mag = c->d_mag[W]; Q->ang = c->d_ang[W]; Q->fromdB = c->d_fromdB;
  return 0;
}

// one launcher for both entry points: desc == NULL -> every block uses Q0
static int cqn_launch(vb200_ctx *c, const CqnDev &Q0, const CqnDev &Q1, const vb200_block_desc *d_desc, int nblocks,
                      const float *d_mdct, int32_t *d_iwork, int32_t *d_nonzero, cudaStream_t st) {
  int rc;
  const int wpb = 4;
  const long tasks = (long)nblocks * (Q0.n / 32);
  const bool v1 = getenv("VB200_CQN_V1") && atoi(getenv("VB200_CQN_V1"));
  if (!v1 && tasks < (1L << 30) && (Q0.n & (Q0.n - 1)) == 0 && (Q0.ch == 1 || (Q0.ch == 2 && Q0.steps <= 1))) {
    const int grid = grid_for(c, (int)((tasks + wpb - 1) / wpb), 16);
    if (Q0.ch == 1) k_cqn_fast<1><<<grid, wpb * 32, 0, st>>>(Q0, Q1, d_desc, nblocks, d_mdct, d_iwork, d_nonzero);
    else k_cqn_fast<2><<<grid, wpb * 32, 0, st>>>(Q0, Q1, d_desc, nblocks, d_mdct, d_iwork, d_nonzero);
  } else {
    const size_t smem = (size_t)wpb * (CQN_COLS * Q0.ch * 32 * sizeof(float) + Q0.ch * sizeof(int));
    if (smem > 200 * 1024) return fail(VB200_EIMPL, "too many channels for the coupling kernel");
    if ((rc = set_smem(k_cqn, smem))) return rc;
    k_cqn<<<grid_for(c, (int)((tasks + wpb - 1) / wpb), 8), wpb * 32, smem, st>>>(Q0, Q1, d_desc, nblocks, d_mdct,
                                                                              d_iwork, d_nonzero);
  }
  if ((rc = post_launch(c))) return rc;
  if (Q0.steps > 0) {
    k_cqn_nonzero<<<(nblocks + 127) / 128, 128, 0, st>>>(nblocks, Q0.ch, Q0.steps, Q0.mag, Q0.ang, d_nonzero);
    if ((rc = post_launch(c))) return rc;
  }
  return 0;
}

extern "C" int vb200_couple_quantize_normalize_dev(vb200_ctx *c, int W, int blocktype, int blobno, int nblocks,
                                                   const float *d_mdct, int32_t *d_iwork, int32_t *d_nonzero,
                                                   void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0) return 0;
  CqnDev Q; int rc;
  if ((rc = cqn_setup(c, W, blocktype, blobno, &Q))) return rc;
  return cqn_launch(c, Q, Q, nullptr, nblocks, d_mdct, d_iwork, d_nonzero, (cudaStream_t)stream);
}

extern "C" int vb200_couple_quantize_normalize(vb200_ctx *c, int W, int blocktype, int blobno, int nblocks,
                                               const float *mdct, int32_t *iwork, int32_t *nonzero) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels, n = c->dx[W].N / 2;
  HostIO io{c};
  void *dm, *di, *dz; int rc;
  if ((rc = io.h2d(mdct, sizeof(float) * (size_t)nblocks * ch * n, &dm))) return rc;
  if ((rc = io.h2d(iwork, sizeof(int32_t) * (size_t)nblocks * ch * n, &di))) return rc;
  if ((rc = io.h2d(nonzero, sizeof(int32_t) * (size_t)nblocks * ch, &dz))) return rc;
  if ((rc = vb200_couple_quantize_normalize_dev(c, W, blocktype, blobno, nblocks, (const float *)dm,
                                                (int32_t *)di, (int32_t *)dz, c->s_main))) return rc;
  if ((rc = io.d2h(iwork, di, sizeof(int32_t) * (size_t)nblocks * ch * n))) return rc;
  if ((rc = io.d2h(nonzero, dz, sizeof(int32_t) * (size_t)nblocks * ch))) return rc;
  return io.sync();
}

// ======================================================================== //
// decode
extern "C" int vb200_synthesis_dev(vb200_ctx *c, int nstreams, int nblk, const int32_t *d_Wseq,
                                   const int64_t *d_coef_off, const float *d_coef,
                                   const int64_t *d_pcm_off, float *d_pcm, int64_t pcm_stride, void *stream) {
  CHECK_CTX(c);
  if (nstreams <= 0 || nblk <= 0) return 0;
  const int ch = c->setup.channels, N1 = c->dx[1].N;
  const size_t smem = sizeof(float) * ((size_t)N1 / 2 + N1 + N1 / 2);
  int rc = set_smem(k_synthesis<false>, smem); if (rc) return rc;
  k_synthesis<false><<<grid_for(c, nstreams * ch, 8), threads_for(N1), smem, (cudaStream_t)stream>>>(
      c->dx[0], c->dx[1], c->dwin, ch, nstreams, nblk, d_Wseq, (const long long *)d_coef_off, d_coef,
      (const long long *)d_pcm_off, d_pcm, (long long)pcm_stride);
  return post_launch(c);
}

extern "C" int vb200_synthesis_s16_dev(vb200_ctx *c, int nstreams, int nblk, const int32_t *d_Wseq,
                                       const int64_t *d_coef_off, const float *d_coef,
                                       const int64_t *d_pcm_off, int16_t *d_pcm16, int64_t pcm_stride, void *stream) {
  CHECK_CTX(c);
  if (nstreams <= 0 || nblk <= 0) return 0;
  const int ch = c->setup.channels, N1 = c->dx[1].N;
  const size_t smem = sizeof(float) * ((size_t)N1 / 2 + N1 + N1 / 2);
  int rc = set_smem(k_synthesis<true>, smem); if (rc) return rc;
  k_synthesis<true><<<grid_for(c, nstreams * ch, 8), threads_for(N1), smem, (cudaStream_t)stream>>>(
      c->dx[0], c->dx[1], c->dwin, ch, nstreams, nblk, d_Wseq, (const long long *)d_coef_off, d_coef,
      (const long long *)d_pcm_off, d_pcm16, (long long)pcm_stride);
  return post_launch(c);
}

extern "C" int vb200_synthesis(vb200_ctx *c, int nstreams, int nblk, const int32_t *Wseq,
                               const int64_t *coef_off, const float *coef, int64_t coef_len,
                               const int64_t *pcm_off, float *pcm, int64_t pcm_stride) {
  CHECK_CTX(c);
  if (nstreams <= 0 || nblk <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels;
  const size_t nb = (size_t)nstreams * nblk;
  for (size_t i = 0; i < nb; i++) if (Wseq[i] < 0 || Wseq[i] > 1) return fail(VB200_EINVAL, "Wseq values must be 0/1");
  HostIO io{c};
  void *dW, *dco, *dc, *dpo, *dp; int rc;
  if ((rc = io.h2d(Wseq, sizeof(int32_t) * nb, &dW))) return rc;
  if ((rc = io.h2d(coef_off, sizeof(int64_t) * nb, &dco))) return rc;
  if ((rc = io.h2d(coef, sizeof(float) * (size_t)coef_len, &dc))) return rc;
  if ((rc = io.h2d(pcm_off, sizeof(int64_t) * nb, &dpo))) return rc;
  const size_t pbytes = sizeof(float) * (size_t)nstreams * ch * (size_t)pcm_stride;
  if ((rc = io.h2d(nullptr, pbytes, &dp))) return rc;
  CU(cudaMemsetAsync(dp, 0, pbytes, c->s_main));
  if ((rc = vb200_synthesis_dev(c, nstreams, nblk, (const int32_t *)dW, (const int64_t *)dco, (const float *)dc,
                                (const int64_t *)dpo, (float *)dp, pcm_stride, c->s_main))) return rc;
  if ((rc = io.d2h(pcm, dp, pbytes))) return rc;
  return io.sync();
}

extern "C" int vb200_decouple_dev(vb200_ctx *c, int W, int nblocks, float *d_res, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0 || c->setup.coupling_steps[W] <= 0) return 0;
  const int n = c->dx[W].N / 2;
  const long long total = (long long)nblocks * n;
  k_decouple<<<grid_for(c, (int)((total + 255) / 256), 16), 256, 0, (cudaStream_t)stream>>>(
      n, c->setup.channels, c->setup.coupling_steps[W], c->d_mag[W], c->d_ang[W], total, d_res);
  return post_launch(c);
}

extern "C" int vb200_decouple(vb200_ctx *c, int W, int nblocks, float *res) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t bytes = sizeof(float) * (size_t)nblocks * c->setup.channels * (c->dx[W].N / 2);
  HostIO io{c};
  void *d; int rc;
  if ((rc = io.h2d(res, bytes, &d))) return rc;
  if ((rc = vb200_decouple_dev(c, W, nblocks, (float *)d, c->s_main))) return rc;
  if ((rc = io.d2h(res, d, bytes))) return rc;
  return io.sync();
}

// ======================================================================== //
// floor 1
static int floor1_args(vb200_ctx *c, int W, int floor_sel, int nrows, Floor1Args *a) {
  if (floor_sel >= VB200_MAX_SUBMAPS) return fail(VB200_EINVAL, "floor_sel");
  const int ch = c->setup.channels;
  if (floor_sel >= 0) {
    if (c->setup.floor1[W][floor_sel].posts <= 0) return fail(VB200_EINVAL, "no floor1 setup for this floor_sel");
  } else {
    for (int k = 0; k < ch; k++)
      if (c->setup.floor1[W][c->setup.chmux[W][k]].posts <= 0)
        return fail(VB200_EINVAL, "no floor1 setup for a channel's submap");
  }
  a->floors = c->d_floor[W]; a->chmux = c->d_chmux[W];
  a->channels = ch; a->floor_sel = floor_sel; a->nrows = nrows; a->n = c->dx[W].N / 2;
  return 0;
}

extern "C" int vb200_floor1_fit_dev(vb200_ctx *c, int W, int floor_sel, int nrows, const float *d_logmdct,
                                    const float *d_logmask, int32_t *d_posts, int32_t *d_fit_nonzero, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  Floor1Args a; int rc;
  if ((rc = floor1_args(c, W, floor_sel, nrows, &a))) return rc;
  const size_t smem = sizeof(Floor1Dev) * VB200_MAX_SUBMAPS + floor1_fit_smem_per_warp(a.n) * F1_WARPS;
  if ((rc = set_smem(k_floor1_fit, smem))) return rc;
  const int ctas = (nrows + F1_WARPS - 1) / F1_WARPS;
  k_floor1_fit<<<grid_for(c, ctas, 8), 32 * F1_WARPS, smem, (cudaStream_t)stream>>>(a, d_logmdct, d_logmask, d_posts, d_fit_nonzero);
  return post_launch(c);
}

extern "C" int vb200_floor1_render_dev(vb200_ctx *c, int W, int floor_sel, int nrows, int32_t *d_posts,
                                       const int32_t *d_fit_nonzero, int32_t *d_ilogmask, int32_t *d_nonzero,
                                       void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  Floor1Args a; int rc;
  if ((rc = floor1_args(c, W, floor_sel, nrows, &a))) return rc;
  const int ctas = (nrows + F1_WARPS - 1) / F1_WARPS;
  k_floor1_render<<<grid_for(c, ctas, 8), 32 * F1_WARPS, 0, (cudaStream_t)stream>>>(a, d_posts, d_fit_nonzero, d_ilogmask, d_nonzero);
  return post_launch(c);
}

extern "C" int vb200_floor1_fit(vb200_ctx *c, int W, int floor_sel, int nrows, const float *logmdct,
                                const float *logmask, int32_t *posts, int32_t *fit_nonzero) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t n = c->dx[W].N / 2;
  HostIO io{c};
  void *da, *db, *dp, *dz; int rc;
  if ((rc = io.h2d(logmdct, sizeof(float) * nrows * n, &da))) return rc;
  if ((rc = io.h2d(logmask, sizeof(float) * nrows * n, &db))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(int32_t) * (size_t)nrows * VB200_FLOOR1_STRIDE, &dp))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(int32_t) * (size_t)nrows, &dz))) return rc;
  if ((rc = vb200_floor1_fit_dev(c, W, floor_sel, nrows, (const float *)da, (const float *)db, (int32_t *)dp,
                                 (int32_t *)dz, c->s_main))) return rc;
  if ((rc = io.d2h(posts, dp, sizeof(int32_t) * (size_t)nrows * VB200_FLOOR1_STRIDE))) return rc;
  if ((rc = io.d2h(fit_nonzero, dz, sizeof(int32_t) * (size_t)nrows))) return rc;
  return io.sync();
}

extern "C" int vb200_floor1_render(vb200_ctx *c, int W, int floor_sel, int nrows, int32_t *posts,
                                   const int32_t *fit_nonzero, int32_t *ilogmask, int32_t *nonzero) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t n = c->dx[W].N / 2;
  HostIO io{c};
  void *dp, *dz, *di, *dn; int rc;
  if ((rc = io.h2d(posts, sizeof(int32_t) * (size_t)nrows * VB200_FLOOR1_STRIDE, &dp))) return rc;
  if ((rc = io.h2d(fit_nonzero, sizeof(int32_t) * (size_t)nrows, &dz))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(int32_t) * nrows * n, &di))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(int32_t) * (size_t)nrows, &dn))) return rc;
  if ((rc = vb200_floor1_render_dev(c, W, floor_sel, nrows, (int32_t *)dp, (const int32_t *)dz, (int32_t *)di,
                                    (int32_t *)dn, c->s_main))) return rc;
  if ((rc = io.d2h(posts, dp, sizeof(int32_t) * (size_t)nrows * VB200_FLOOR1_STRIDE))) return rc;
  if ((rc = io.d2h(ilogmask, di, sizeof(int32_t) * nrows * n))) return rc;
  if ((rc = io.d2h(nonzero, dn, sizeof(int32_t) * (size_t)nrows))) return rc;
  return io.sync();
}

// ======================================================================== //
// the whole per-block encode DSP (Phase A -> floor1 fit -> floor render -> Phase B)
static_assert(sizeof(vb200_encode_io) == 128, "vb200_encode_io layout (mirrored by vorbis_b200/abi.py)");
struct EncScratch {
  float *mdct, *logmdct, *logmask, *logfft, *lmax, *gmax;
  int32_t *fitnz;
  int32_t *iw32;                     // residue as int32 when the caller wants int16 out
};

// int32 residue -> saturated int16, counting per block what did not fit (VB200_IWORK_S16)
__global__ void __launch_bounds__(256)
k_pack_s16(const int4 *__restrict__ src, short4 *__restrict__ dst, long nvec, int vec_per_block_log2,
           int vec_per_block, int32_t *__restrict__ overflow) {
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    const int4 a = __ldcs(src + v);
    int bad = 0;
    auto sat = [&](int x) { if (x > 32767) { bad++; return (short)32767; } if (x < -32768) { bad++; return (short)-32768; } return (short)x; };
    short4 o;
    o.x = sat(a.x); o.y = sat(a.y); o.z = sat(a.z); o.w = sat(a.w);
    __stcs(dst + v, o);
    // clipping is the rare case: the division for channel counts that are not a power of two only runs there
    if (bad) atomicAdd(overflow + (vec_per_block_log2 >= 0 ? (v >> vec_per_block_log2) : (v / vec_per_block)), bad);
  }
}

static size_t enc_pcm_bytes(const vb200_encode_io *io, int ch, int N, int nstreams, int bps) {
  switch (io->pcm_fmt) {
    case VB200_PCM_F32_BLOCKS: return sizeof(float) * (size_t)nstreams * bps * ch * N;
    case VB200_PCM_F32_PLANAR: return sizeof(float) * (size_t)nstreams * ch * (size_t)io->stream_stride;
    default: return sizeof(int16_t) * (size_t)nstreams * ch * (size_t)io->stream_stride;
  }
}

static int enc_check(vb200_ctx *c, int W, int nstreams, int bps, int blobno, const vb200_encode_io *io) {
  if (!io || !io->pcm || !io->desc || !io->posts || !io->nonzero || !io->iwork || !io->ampmax_out)
    return fail(VB200_EINVAL, "encode io pointers");
  if (io->iwork_fmt != VB200_IWORK_S32 && io->iwork_fmt != VB200_IWORK_S16) return fail(VB200_EINVAL, "iwork format");
  if (io->iwork_fmt == VB200_IWORK_S16 && !io->overflow) return fail(VB200_EINVAL, "int16 residue needs the overflow array");
  if (nstreams <= 0 || bps <= 0) return fail(VB200_EINVAL, "nstreams/blocks_per_stream");
  if (c->n_psy != 4) return fail(VB200_EIMPL, "context has no psy lookups");
  if (blobno < 0 || blobno >= VB200_PACKETBLOBS) return fail(VB200_EINVAL, "blobno");
  const int fmt = io->pcm_fmt;
  if (fmt != VB200_PCM_F32_BLOCKS && fmt != VB200_PCM_F32_PLANAR && fmt != VB200_PCM_S16_INTERLEAVED)
    return fail(VB200_EINVAL, "pcm format");
  if (fmt != VB200_PCM_F32_BLOCKS) {
    if (io->hop <= 0 || io->stream_stride <= 0) return fail(VB200_EINVAL, "hop/stream_stride");
    if (fmt == VB200_PCM_F32_PLANAR && ((io->hop & 3) || (io->stream_stride & 3)))
      return fail(VB200_EINVAL, "hop and stream_stride must be multiples of 4 for float PCM");
    if ((int64_t)(bps - 1) * io->hop + c->dx[W].N > io->stream_stride)
      return fail(VB200_EINVAL, "blocks exceed the stream buffer");
  }
  return 0;
}

// all pointers device; S holds the float intermediates
static int encode_launch(vb200_ctx *c, int W, int nstreams, int bps, int blobno, const vb200_encode_io *d,
                         const EncScratch &S, cudaStream_t st) {
  const int ch = c->setup.channels, nblocks = nstreams * bps, rows = nblocks * ch;
  vb200_phaseA_io a;
  memset(&a, 0, sizeof(a));
  a.desc = d->desc; a.mdct = S.mdct; a.logmdct = S.logmdct; a.logmask = S.logmask; a.ampmax_out = d->ampmax_out;
  PcmSrc ps; const PcmSrc *pp = nullptr;
  if (d->pcm_fmt == VB200_PCM_F32_BLOCKS) a.pcm = (const float *)d->pcm;
  else { ps.base = d->pcm; ps.fmt = d->pcm_fmt; ps.bps = bps; ps.hop = d->hop; ps.stride = d->stream_stride; ps.blk_src = nullptr; pp = &ps; }
  int rc;
  if ((rc = phaseA_launch(c, W, nblocks, &a, d->independent ? 0 : nstreams, bps, d->ampmax0, st,
                          S.logfft, S.lmax, S.gmax, pp))) return rc;
  if ((rc = vb200_floor1_fit_dev(c, W, -1, rows, S.logmdct, S.logmask, d->posts, S.fitnz, st))) return rc;
  if (c->profiling) CU(cudaEventRecord(c->ev[4], st));
  const bool s16 = d->iwork_fmt == VB200_IWORK_S16;
  int32_t *iw = s16 ? S.iw32 : (int32_t *)d->iwork;
  if ((rc = vb200_floor1_render_dev(c, W, -1, rows, d->posts, S.fitnz, iw, d->nonzero, st))) return rc;
  if (c->profiling) CU(cudaEventRecord(c->ev[5], st));
  CqnDev Q0, Q1;
  if ((rc = cqn_setup(c, W, 0, blobno, &Q0))) return rc;
  if ((rc = cqn_setup(c, W, 1, blobno, &Q1))) return rc;
  if ((rc = cqn_launch(c, Q0, Q1, d->desc, nblocks, S.mdct, iw, d->nonzero, st))) return rc;
  if (d->classes &&
      (rc = vb200_residue_classify_dev(c, W, nblocks, iw, d->nonzero, d->classes, (int)d->class_stride, st))) return rc;
  if (s16) {
    const int n = c->dx[W].N / 2;
    const long nvec = (long)rows * n / 4;
    int lg = 0;
    while ((1L << lg) < (long)ch * n / 4) lg++;
    if ((1L << lg) != (long)ch * n / 4) lg = -1;      // channels not a power of two (5.1): per-block counts by division
    CU(cudaMemsetAsync(d->overflow, 0, sizeof(int32_t) * (size_t)nblocks, st));
    k_pack_s16<<<grid_for(c, (int)((nvec + 255) / 256), 8), 256, 0, st>>>((const int4 *)iw, (short4 *)d->iwork, nvec, lg, (int)((long)ch * n / 4), d->overflow);
    if ((rc = post_launch(c))) return rc;
  }
  if (c->profiling) CU(cudaEventRecord(c->ev[6], st));
  return 0;
}

static int enc_scratch(DevBuf *B, size_t rows, size_t nblocks, size_t n, const vb200_encode_io *d, bool s16, EncScratch *S) {
  void *p; int rc;
  const size_t big = sizeof(float) * rows * n;
  if (d && d->mdct) S->mdct = d->mdct; else { if ((rc = ensure_buf(B[0], big, &p))) return rc; S->mdct = (float *)p; }
  if (d && d->logmdct) S->logmdct = d->logmdct; else { if ((rc = ensure_buf(B[1], big, &p))) return rc; S->logmdct = (float *)p; }
  if (d && d->logmask) S->logmask = d->logmask; else { if ((rc = ensure_buf(B[2], big, &p))) return rc; S->logmask = (float *)p; }
  if ((rc = ensure_buf(B[3], big, &p))) return rc; S->logfft = (float *)p;
  if ((rc = ensure_buf(B[4], sizeof(float) * rows, &p))) return rc; S->lmax = (float *)p;
  if ((rc = ensure_buf(B[5], sizeof(float) * nblocks, &p))) return rc; S->gmax = (float *)p;
  if ((rc = ensure_buf(B[6], sizeof(int32_t) * rows, &p))) return rc; S->fitnz = (int32_t *)p;
  S->iw32 = nullptr;
  if (s16) { if ((rc = ensure_buf(B[7], sizeof(int32_t) * rows * n, &p))) return rc; S->iw32 = (int32_t *)p; }
  return 0;
}

extern "C" int vb200_encode_dsp_dev(vb200_ctx *c, int W, int nstreams, int bps, int blobno,
                                    const vb200_encode_io *d, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  int rc;
  if ((rc = enc_check(c, W, nstreams, bps, blobno, d))) return rc;
  const size_t ch = c->setup.channels, n = c->dx[W].N / 2, nblocks = (size_t)nstreams * bps;
  EncScratch S;
  if ((rc = enc_scratch(c->enc_buf, nblocks * ch, nblocks, n, d, d->iwork_fmt == VB200_IWORK_S16, &S))) return rc;
  // Two half-batches (whole streams each) on two internal streams, every kernel launched with half its
  // grid: the halves drift apart, so CTAs of different kernels (shared-memory bound transform, issue bound
  // psy, latency bound floor fit) share the SMs instead of one kernel type owning the machine at a time.
  int split = 2;                      // measured: 19.24 -> 18.88 ms per 100 000 blocks; 4 staggered pieces: no gain
  { const char *e = getenv("VB200_SPLIT"); if (e) split = atoi(e); }
  size_t split_min = 2048;
  { const char *e = getenv("VB200_SPLIT_MIN"); if (e && atoi(e) > 0) split_min = (size_t)atoi(e); }
  if ((rc = scratch_begin(c, (cudaStream_t)stream))) return rc;
  if (split < 2 || nstreams < 2 || c->profiling || nblocks < split_min) {
    if ((rc = encode_launch(c, W, nstreams, bps, blobno, d, S, (cudaStream_t)stream))) return rc;
    return scratch_end(c, (cudaStream_t)stream);
  }
  cudaStream_t user = (cudaStream_t)stream;
  CU(cudaEventRecord(c->ev_fork, user));
  // pieces: (internal stream, share of the streams).  split 2: two halves.  split 4: four pieces on the
  // two streams, 20/30 % then 30/20 %, so that the two streams are always at different kernels.
  int npieces = 2, pst[4] = {0, 1, 0, 1}, share[4] = {50, 50, 0, 0};
  if (split >= 4 && nstreams >= 8) { npieces = 4; share[0] = 20; share[1] = 30; share[2] = 30; share[3] = 20; }
  { const char *e = getenv("VB200_SPLIT_SKEW"); if (e && atoi(e) > 0 && atoi(e) < 50 && npieces == 4) {
      share[0] = atoi(e); share[1] = 50 - share[0]; share[2] = share[1]; share[3] = share[0]; } }
  c->grid_div = 2;
  int s0 = 0;
  for (int pc = 0; pc < npieces; pc++) {
    int ns = pc == npieces - 1 ? nstreams - s0 : (int)((long)nstreams * share[pc] / 100);
    if (ns < 1) ns = 1;
    if (s0 + ns > nstreams - (npieces - 1 - pc)) ns = nstreams - (npieces - 1 - pc) - s0;
    const size_t b0 = (size_t)s0 * bps, r0 = b0 * ch;
    vb200_encode_io h = *d;
    EncScratch T = S;
    h.pcm = (const char *)d->pcm + enc_pcm_bytes(d, (int)ch, c->dx[W].N, 1, bps) * s0;
    h.desc = d->desc + b0;
    if (d->ampmax0) h.ampmax0 = d->ampmax0 + s0;
    h.posts = d->posts + r0 * VB200_FLOOR1_STRIDE; h.nonzero = d->nonzero + r0; h.ampmax_out = d->ampmax_out + b0;
    h.iwork = (char *)d->iwork + (d->iwork_fmt == VB200_IWORK_S16 ? sizeof(int16_t) : sizeof(int32_t)) * r0 * n;
    if (d->overflow) h.overflow = d->overflow + b0;
    if (d->classes) h.classes = d->classes + r0 * (size_t)d->class_stride;
    T.mdct += r0 * n; T.logmdct += r0 * n; T.logmask += r0 * n; T.logfft += r0 * n;
    T.lmax += r0; T.gmax += b0; T.fitnz += r0;
    if (T.iw32) T.iw32 += r0 * n;
    cudaStream_t st = c->s_split[pst[pc]];
    cudaError_t e = pc < 2 ? cudaStreamWaitEvent(st, c->ev_fork, 0) : cudaSuccess;
    if (e == cudaSuccess) rc = encode_launch(c, W, ns, bps, blobno, &h, T, st);
    if (rc || e != cudaSuccess) { c->grid_div = 1; return rc ? rc : fail(VB200_EFAULT, "encode split", e); }
    s0 += ns;
  }
  for (int k = 0; k < 2; k++) {
    cudaError_t e = cudaEventRecord(c->ev_join[k], c->s_split[k]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(user, c->ev_join[k], 0);
    if (e != cudaSuccess) { c->grid_div = 1; return fail(VB200_EFAULT, "encode split join", e); }
  }
  c->grid_div = 1;
  return scratch_end(c, user);
}

// ---- bitrate-managed mode (lib/mapping0.c:507-573, 596-646): the chain above with three masks, three fits, the
// twelve interpolated curves, and the floor render + couple/quantise/normalise of every one of the 15 curves
static int managed_check(vb200_ctx *c, int W, int nstreams, int bps, const vb200_encode_io *io) {
  int rc;
  if ((rc = enc_check(c, W, nstreams, bps, VB200_PACKETBLOBS / 2, io))) return rc;
  if (io->iwork_fmt != VB200_IWORK_S32) return fail(VB200_EINVAL, "managed mode writes int32 residue");
  if (io->classes) return fail(VB200_EINVAL, "managed mode does not classify");
  return 0;
}

extern "C" int vb200_encode_dsp_managed_dev(vb200_ctx *c, int W, int nstreams, int bps, const vb200_encode_io *d,
                                            void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  int rc;
  if ((rc = managed_check(c, W, nstreams, bps, d))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ch = c->setup.channels, n = c->dx[W].N / 2, nblocks = nstreams * bps;
  const size_t rows = (size_t)nblocks * ch, big = sizeof(float) * rows * n;
  constexpr int NB = VB200_PACKETBLOBS, MID = VB200_PACKETBLOBS / 2;
  DevBuf *B = c->mgd_buf;
  void *p;
  float *mdct, *logmdct, *logmask, *logfft, *lmax, *gmax, *noise, *tone, *alt;
  int32_t *fz3, *present;
  if ((rc = ensure_buf(B[0], big, &p))) return rc; mdct = (float *)p;
  if ((rc = ensure_buf(B[1], big, &p))) return rc; logmdct = (float *)p;
  if ((rc = ensure_buf(B[2], big, &p))) return rc; logmask = (float *)p;
  if ((rc = ensure_buf(B[3], big, &p))) return rc; logfft = (float *)p;
  if ((rc = ensure_buf(B[4], sizeof(float) * rows, &p))) return rc; lmax = (float *)p;
  if ((rc = ensure_buf(B[5], sizeof(float) * nblocks, &p))) return rc; gmax = (float *)p;
  if ((rc = ensure_buf(B[6], big, &p))) return rc; noise = (float *)p;
  if ((rc = ensure_buf(B[7], big, &p))) return rc; tone = (float *)p;
  if ((rc = ensure_buf(B[8], big, &p))) return rc; alt = (float *)p;
  if ((rc = ensure_buf(B[9], sizeof(int32_t) * rows * 3, &p))) return rc; fz3 = (int32_t *)p;
  if ((rc = ensure_buf(B[10], sizeof(int32_t) * rows * NB, &p))) return rc; present = (int32_t *)p;
  if ((rc = scratch_begin(c, st))) return rc;
  vb200_phaseA_io a;
  memset(&a, 0, sizeof(a));
  a.desc = d->desc; a.mdct = mdct; a.logmdct = logmdct; a.logmask = logmask; a.ampmax_out = d->ampmax_out;
  a.tap_noise = noise; a.tap_tone = tone;
  PcmSrc ps; const PcmSrc *pp = nullptr;
  if (d->pcm_fmt == VB200_PCM_F32_BLOCKS) a.pcm = (const float *)d->pcm;
  else { ps.base = d->pcm; ps.fmt = d->pcm_fmt; ps.bps = bps; ps.hop = d->hop; ps.stride = d->stream_stride; ps.blk_src = nullptr; pp = &ps; }
  if ((rc = phaseA_launch(c, W, nblocks, &a, d->independent ? 0 : nstreams, bps, d->ampmax0, st, logfft, lmax, gmax, pp)))
    return rc;
  const size_t pblob = rows * VB200_FLOOR1_STRIDE;
  int32_t *fz_lo = fz3, *fz_mid = fz3 + rows, *fz_hi = fz3 + 2 * rows;
  // the middle curve (select 1, what un-managed mode codes), then the low-noise (2) and high-noise (0) masks
  if ((rc = vb200_floor1_fit_dev(c, W, -1, (int)rows, logmdct, logmask, d->posts + MID * pblob, fz_mid, st))) return rc;
  const PsyDev &P0 = c->dpsy[(W ? 2 : 0)], &P1 = c->dpsy[(W ? 2 : 0) + 1];
  const long long total = (long long)rows * n;
  for (int pass = 0; pass < 2; pass++) {
    const int sel = pass ? 0 : 2;
    k_mix_select<<<grid_for(c, (int)((total + 255) / 256), 8), 256, 0, st>>>(P0, P1, d->desc, ch, total, sel, noise, tone,
                                                                            logmdct, alt);
    if ((rc = post_launch(c))) return rc;
    if ((rc = vb200_floor1_fit_dev(c, W, -1, (int)rows, logmdct, alt, d->posts + (sel ? NB - 1 : 0) * pblob,
                                   sel ? fz_hi : fz_lo, st))) return rc;
  }
  k_floor1_interpolate<<<grid_for(c, (int)((rows * VB200_FLOOR1_STRIDE + 255) / 256), 8), 256, 0, st>>>(
      (long long)rows, d->posts, fz_lo, fz_mid, fz_hi, present);
  if ((rc = post_launch(c))) return rc;
  for (int k = 0; k < NB; k++) {
    int32_t *iw = (int32_t *)d->iwork + (size_t)k * rows * n, *nz = d->nonzero + (size_t)k * rows;
    if ((rc = vb200_floor1_render_dev(c, W, -1, (int)rows, d->posts + k * pblob, present + (size_t)k * rows, iw, nz, st)))
      return rc;
    CqnDev Q0, Q1;
    if ((rc = cqn_setup(c, W, 0, k, &Q0))) return rc;
    if ((rc = cqn_setup(c, W, 1, k, &Q1))) return rc;
    if ((rc = cqn_launch(c, Q0, Q1, d->desc, nblocks, mdct, iw, nz, st))) return rc;
  }
  return scratch_end(c, st);
}

// host buffers: one synchronous H2D - compute - D2H round trip (managed mode is not the throughput path)
extern "C" int vb200_encode_dsp_managed(vb200_ctx *c, int W, int nstreams, int bps, const vb200_encode_io *h) {
  CHECK_CTX(c); CHECK_W(W);
  int rc;
  if ((rc = managed_check(c, W, nstreams, bps, h))) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels, N = c->dx[W].N, n = N / 2;
  const size_t nb = (size_t)nstreams * bps, rows = nb * ch;
  constexpr int NB = VB200_PACKETBLOBS;
  cudaStream_t st = c->s_main;
  DevBuf *B = c->mgd_buf;
  void *p;
  vb200_encode_io d = *h;
  d.mdct = d.logmdct = d.logmask = nullptr;
  const size_t pcm_bytes = enc_pcm_bytes(h, ch, N, nstreams, bps);
  if ((rc = ensure_buf(B[12], pcm_bytes, &p))) return rc; d.pcm = p;
  if ((rc = ensure_buf(B[13], sizeof(vb200_block_desc) * nb, &p))) return rc; d.desc = (const vb200_block_desc *)p;
  if ((rc = ensure_buf(B[14], sizeof(float) * nstreams, &p))) return rc; d.ampmax0 = h->ampmax0 ? (const float *)p : nullptr;
  if ((rc = ensure_buf(B[15], sizeof(int32_t) * NB * rows * VB200_FLOOR1_STRIDE, &p))) return rc; d.posts = (int32_t *)p;
  if ((rc = ensure_buf(B[16], sizeof(int32_t) * NB * rows, &p))) return rc; d.nonzero = (int32_t *)p;
  if ((rc = ensure_buf(B[17], sizeof(int32_t) * NB * rows * n, &p))) return rc; d.iwork = p;
  if ((rc = ensure_buf(B[18], sizeof(float) * nb, &p))) return rc; d.ampmax_out = (float *)p;
  CU(cudaMemcpyAsync((void *)d.pcm, h->pcm, pcm_bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync((void *)d.desc, h->desc, sizeof(vb200_block_desc) * nb, cudaMemcpyHostToDevice, st));
  if (h->ampmax0) CU(cudaMemcpyAsync((void *)d.ampmax0, h->ampmax0, sizeof(float) * nstreams, cudaMemcpyHostToDevice, st));
  if ((rc = vb200_encode_dsp_managed_dev(c, W, nstreams, bps, &d, st))) return rc;
  CU(cudaMemcpyAsync(h->posts, d.posts, sizeof(int32_t) * NB * rows * VB200_FLOOR1_STRIDE, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->nonzero, d.nonzero, sizeof(int32_t) * NB * rows, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->iwork, d.iwork, sizeof(int32_t) * NB * rows * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->ampmax_out, d.ampmax_out, sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int vb200_encode_dsp(vb200_ctx *c, int W, int nstreams, int bps, int blobno, const vb200_encode_io *h) {
  CHECK_CTX(c); CHECK_W(W);
  int rc;
  if ((rc = enc_check(c, W, nstreams, bps, blobno, h))) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels, N = c->dx[W].N, n = N / 2;
  int chunk_blocks = 8192;                           // measured on B200 (ramped schedule): 2048 -> 5.56, 4096 -> 6.11, 8192 -> 6.16 M blocks/s end to end
  { const char *e = getenv("VB200_CHUNK_BLOCKS"); if (e && atoi(e) > 0) chunk_blocks = atoi(e); }
  int cs = chunk_blocks / bps;                       // whole streams per chunk
  if (cs < 1) cs = 1;
  if (cs > nstreams) cs = nstreams;
  const size_t pcm_per_stream = enc_pcm_bytes(h, ch, N, 1, bps);
  // Chunk schedule: the pipeline's fill (first H2D + first kernels before anything overlaps) and drain (last kernels
  // + last D2H) are exposed, so the first and the last chunks are small (cs/4, cs/2) and the middle ones full size.
  int ramp = 1;
  { const char *e = getenv("VB200_CHUNK_RAMP"); if (e) ramp = atoi(e); }
  std::vector<int> sched;
  {
    int left = nstreams;
    const int q = cs / 4 > 0 ? cs / 4 : 1, hlf = cs / 2 > 0 ? cs / 2 : 1;
    if (ramp && nstreams >= 6 * cs) {
      const int head[2] = {q, hlf};
      for (int k = 0; k < 2; k++) { sched.push_back(head[k]); left -= head[k]; }
      const int tail = q + hlf;
      while (left - tail >= cs) { sched.push_back(cs); left -= cs; }
      if (left - tail > 0) { sched.push_back(left - tail); left = tail; }
      sched.push_back(hlf); sched.push_back(left - hlf);
    } else {
      while (left > 0) { const int t = left < cs ? left : cs; sched.push_back(t); left -= t; }
    }
  }
  // Three-stage pipeline over ENC_SETS buffer sets: all host->device copies on one stream, all device->host
  // copies on another (one DMA engine per direction anyway), the kernels of consecutive chunks alternately on two
  // compute streams (a chunk's kernel tails overlap the next chunk's kernels).  Events carry the order
  // H2D(k) -> kernels(k) -> D2H(k) -> H2D(k + ENC_SETS); nothing else is ordered, so the copy engines run ahead
  // of / behind the kernels instead of every lane alternating copy and compute on its own stream.
  constexpr int ENC_SETS = 4;
  cudaStream_t s_h2d = c->s_enc[2], s_d2h = c->s_d2h;
  for (int s0 = 0, it = 0; it < (int)sched.size(); s0 += sched[it], it++) {
    const int L = it % ENC_SETS, ns = sched[it];
    const size_t nb = (size_t)ns * bps, b0 = (size_t)s0 * bps, rows = nb * ch, r0 = b0 * ch;
    cudaStream_t st = c->s_enc[it & 1];
    DevBuf *B = c->enc_lane[L];
    void *p;
    vb200_encode_io d = *h;
    d.mdct = d.logmdct = d.logmask = nullptr;
    if ((rc = ensure_buf(B[8], pcm_per_stream * cs, &p))) return rc; d.pcm = p;
    if ((rc = ensure_buf(B[9], sizeof(vb200_block_desc) * (size_t)cs * bps, &p))) return rc; d.desc = (const vb200_block_desc *)p;
    if ((rc = ensure_buf(B[10], sizeof(float) * cs, &p))) return rc; d.ampmax0 = h->ampmax0 ? (const float *)p : nullptr;
    if ((rc = ensure_buf(B[11], sizeof(int32_t) * (size_t)cs * bps * ch * VB200_FLOOR1_STRIDE, &p))) return rc; d.posts = (int32_t *)p;
    if ((rc = ensure_buf(B[12], sizeof(int32_t) * (size_t)cs * bps * ch, &p))) return rc; d.nonzero = (int32_t *)p;
    const bool s16 = h->iwork_fmt == VB200_IWORK_S16;
    const size_t isz = s16 ? sizeof(int16_t) : sizeof(int32_t);
    if ((rc = ensure_buf(B[13], isz * (size_t)cs * bps * ch * n, &p))) return rc; d.iwork = p;
    if (s16) { if ((rc = ensure_buf(B[15], sizeof(int32_t) * (size_t)cs * bps, &p))) return rc; d.overflow = (int32_t *)p; }
    if (h->classes) {
      if ((rc = ensure_buf(B[16], sizeof(int32_t) * (size_t)cs * bps * ch * (size_t)h->class_stride, &p))) return rc;
      d.classes = (int32_t *)p;
    }
    if ((rc = ensure_buf(B[14], sizeof(float) * (size_t)cs * bps, &p))) return rc; d.ampmax_out = (float *)p;
    EncScratch S;
    if ((rc = enc_scratch(B, (size_t)cs * bps * ch, (size_t)cs * bps, n, nullptr, s16, &S))) return rc;
    if (it >= ENC_SETS) CU(cudaStreamWaitEvent(s_h2d, c->ev_d2h[L], 0));   // this set's previous chunk is back on the host
    CU(cudaMemcpyAsync((void *)d.pcm, (const char *)h->pcm + pcm_per_stream * s0, pcm_per_stream * ns, cudaMemcpyHostToDevice, s_h2d));
    CU(cudaMemcpyAsync((void *)d.desc, h->desc + b0, sizeof(vb200_block_desc) * nb, cudaMemcpyHostToDevice, s_h2d));
    if (h->ampmax0) CU(cudaMemcpyAsync((void *)d.ampmax0, h->ampmax0 + s0, sizeof(float) * ns, cudaMemcpyHostToDevice, s_h2d));
    CU(cudaEventRecord(c->ev_h2d[L], s_h2d));
    CU(cudaStreamWaitEvent(st, c->ev_h2d[L], 0));
    if ((rc = encode_launch(c, W, ns, bps, blobno, &d, S, st))) return rc;
    CU(cudaEventRecord(c->ev_cmp[L], st));
    CU(cudaStreamWaitEvent(s_d2h, c->ev_cmp[L], 0));
    st = s_d2h;
    CU(cudaMemcpyAsync(h->posts + r0 * VB200_FLOOR1_STRIDE, d.posts, sizeof(int32_t) * rows * VB200_FLOOR1_STRIDE, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->nonzero + r0, d.nonzero, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync((char *)h->iwork + isz * r0 * n, d.iwork, isz * rows * n, cudaMemcpyDeviceToHost, st));
    if (s16) CU(cudaMemcpyAsync(h->overflow + b0, d.overflow, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
    if (h->classes) CU(cudaMemcpyAsync(h->classes + r0 * (size_t)h->class_stride, d.classes,
                                       sizeof(int32_t) * rows * (size_t)h->class_stride, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->ampmax_out + b0, d.ampmax_out, sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
    if (h->mdct) CU(cudaMemcpyAsync(h->mdct + r0 * n, S.mdct, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    if (h->logmdct) CU(cudaMemcpyAsync(h->logmdct + r0 * n, S.logmdct, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    if (h->logmask) CU(cudaMemcpyAsync(h->logmask + r0 * n, S.logmask, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->ev_d2h[L], s_d2h));
  }
  for (auto &st : c->s_enc) CU(cudaStreamSynchronize(st));
  CU(cudaStreamSynchronize(c->s_d2h));
  return 0;
}

// ======================================================================== //
// whole streams: block planning and the two size batches
static int plan_check(vb200_ctx *c) {
  if (c->n_psy != 4) return fail(VB200_EINVAL, "context has no psy setup");
  if ((c->setup.blocksizes[0] / 4) % PLAN_STEP) return fail(VB200_EIMPL, "blocksizes[0]/4 must be a multiple of the 64-sample envelope step");
  return 0;
}

// device: marks -> plan (slots final), gather tables and descriptors of both sizes; d_totals[2] on the device
static int plan_launch(vb200_ctx *c, int nstreams, const int32_t *d_mark, int64_t mark_stride, int nsteps,
                       const int64_t *d_len, const int64_t *d_eof, int max_blocks, vb200_stream_block *d_plan,
                       int32_t *d_nblocks, const int cap[2], int2 *d_src[2], vb200_block_desc *d_desc[2],
                       int32_t *d_totals, cudaStream_t st) {
  void *p; int rc;
  if ((rc = ensure_buf(c->str_buf[0], sizeof(int32_t) * 2 * (size_t)nstreams, &p))) return rc;
  int32_t *d_counts = (int32_t *)p;
  if ((rc = ensure_buf(c->str_buf[1], sizeof(int32_t) * 2 * (size_t)nstreams, &p))) return rc;
  int32_t *d_offs = (int32_t *)p;
  const int g = (nstreams + 127) / 128;
  k_plan_blocks<<<g, 128, 0, st>>>(nstreams, c->setup.blocksizes[0], c->setup.blocksizes[1], d_mark, mark_stride, nsteps,
                                   d_len, d_eof, max_blocks, d_plan, d_nblocks, d_counts);
  k_plan_offsets<<<1, 32, 0, st>>>(nstreams, d_counts, d_offs, d_totals);
  k_plan_fill<<<g, 128, 0, st>>>(nstreams, max_blocks, d_plan, d_nblocks, d_offs, cap[0], cap[1],
                                 d_src[0], d_src[1], d_desc[0], d_desc[1]);
  return post_launch(c, 3);
}

extern "C" int vb200_plan_blocks(vb200_ctx *c, int nstreams, const int32_t *mark, int64_t mark_stride, int nsteps,
                                 const int64_t *pcm_len, const int64_t *eof, int max_blocks,
                                 vb200_stream_block *plan, int32_t *nblocks) {
  CHECK_CTX(c);
  int rc;
  if ((rc = plan_check(c))) return rc;
  if (nstreams <= 0) return 0;
  if (!mark || !pcm_len || !plan || !nblocks || max_blocks < 1 || nsteps < 0 || mark_stride < nsteps + 3)
    return fail(VB200_EINVAL, "plan_blocks arguments");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t st = c->s_main;
  void *p;
  const size_t cap = (size_t)nstreams * max_blocks;
  if ((rc = ensure_buf(c->str_buf[2], sizeof(int32_t) * (size_t)nstreams * mark_stride, &p))) return rc; int32_t *d_mark = (int32_t *)p;
  if ((rc = ensure_buf(c->str_buf[3], sizeof(int64_t) * 2 * (size_t)nstreams, &p))) return rc; int64_t *d_len = (int64_t *)p, *d_eof = d_len + nstreams;
  if ((rc = ensure_buf(c->str_buf[4], sizeof(vb200_stream_block) * cap, &p))) return rc; vb200_stream_block *d_plan = (vb200_stream_block *)p;
  if ((rc = ensure_buf(c->str_buf[5], sizeof(int32_t) * ((size_t)nstreams + 2), &p))) return rc; int32_t *d_nb = (int32_t *)p, *d_tot = d_nb + nstreams;
  int2 *d_src[2]; vb200_block_desc *d_desc[2];
  for (int w = 0; w < 2; w++) {
    if ((rc = ensure_buf(c->str_buf[6 + w], sizeof(int2) * cap, &p))) return rc; d_src[w] = (int2 *)p;
    if ((rc = ensure_buf(c->str_buf[8 + w], sizeof(vb200_block_desc) * cap, &p))) return rc; d_desc[w] = (vb200_block_desc *)p;
  }
  CU(cudaMemcpyAsync(d_mark, mark, sizeof(int32_t) * (size_t)nstreams * mark_stride, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_len, pcm_len, sizeof(int64_t) * nstreams, cudaMemcpyHostToDevice, st));
  if (eof) CU(cudaMemcpyAsync(d_eof, eof, sizeof(int64_t) * nstreams, cudaMemcpyHostToDevice, st));
  const int caps[2] = {(int)cap, (int)cap};
  if ((rc = plan_launch(c, nstreams, d_mark, mark_stride, nsteps, d_len, eof ? d_eof : nullptr, max_blocks, d_plan, d_nb,
                        caps, d_src, d_desc, d_tot, st))) return rc;
  CU(cudaMemcpyAsync(plan, d_plan, sizeof(vb200_stream_block) * cap, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(nblocks, d_nb, sizeof(int32_t) * nstreams, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int vb200_encode_streams_dev(vb200_ctx *c, int nstreams, int blobno, vb200_streams_io *d, void *stream) {
  CHECK_CTX(c);
  int rc;
  if ((rc = plan_check(c))) return rc;
  if (!d) return fail(VB200_EINVAL, "null io");
  d->count[0] = d->count[1] = 0;
  if (nstreams <= 0) return 0;
  if (!d->pcm || !d->pcm_len || !d->plan || !d->nblocks || d->max_blocks < 1) return fail(VB200_EINVAL, "encode_streams: pcm, pcm_len, plan, nblocks");
  if (d->pcm_fmt != VB200_PCM_F32_PLANAR && d->pcm_fmt != VB200_PCM_S16_INTERLEAVED) return fail(VB200_EINVAL, "pcm_fmt");
  if (blobno < 0 || blobno >= VB200_PACKETBLOBS) return fail(VB200_EINVAL, "blobno");
  for (int w = 0; w < 2; w++)
    if (d->cap[w] < 0 || (d->cap[w] > 0 && (!d->posts[w] || !d->nonzero[w] || !d->iwork[w] || !d->ampmax_out[w])))
      return fail(VB200_EINVAL, "encode_streams: outputs of a size with capacity > 0");
  if (d->stream_stride < c->setup.blocksizes[1]) return fail(VB200_EINVAL, "stream_stride");
  cudaStream_t st = (cudaStream_t)stream;
  const int ch = c->setup.channels;
  // 1. envelope search over the whole timeline (fresh detector state), 2. marks, 3. plan
  const int nsteps = (int)(d->stream_stride / PLAN_STEP) - PLAN_VE_WIN;
  if (nsteps < 1) return fail(VB200_EINVAL, "stream too short");
  const int64_t mark_stride = nsteps + 4;
  void *p;
  const size_t sw = VB200_VE_STATE_WORDS(ch);
  if ((rc = ensure_buf(c->str_buf[10], sizeof(int32_t) * sw * nstreams, &p))) return rc; int32_t *d_state = (int32_t *)p;
  if ((rc = ensure_buf(c->str_buf[11], (size_t)nstreams * nsteps, &p))) return rc; uint8_t *d_ret = (uint8_t *)p;
  if ((rc = ensure_buf(c->str_buf[2], sizeof(int32_t) * (size_t)nstreams * mark_stride, &p))) return rc; int32_t *d_mark = (int32_t *)p;
  if ((rc = ensure_buf(c->str_buf[5], sizeof(int32_t) * 2, &p))) return rc; int32_t *d_tot = (int32_t *)p;
  int2 *d_src[2]; vb200_block_desc *d_desc[2];
  for (int w = 0; w < 2; w++) {
    const size_t cw = d->cap[w] > 0 ? d->cap[w] : 1;
    if ((rc = ensure_buf(c->str_buf[6 + w], sizeof(int2) * cw, &p))) return rc; d_src[w] = (int2 *)p;
    if ((rc = ensure_buf(c->str_buf[8 + w], sizeof(vb200_block_desc) * cw, &p))) return rc; d_desc[w] = (vb200_block_desc *)p;
  }
  CU(cudaMemsetAsync(d_state, 0, sizeof(int32_t) * sw * nstreams, st));
  if ((rc = vb200_envelope_search_dev(c, nstreams, d->pcm, d->pcm_fmt, d->stream_stride, 0, nsteps, d_state, d_ret, st))) return rc;
  {
    const long long total = (long long)nstreams * mark_stride;
    k_env_marks<<<grid_for(c, (int)((total + 255) / 256), 8), 256, 0, st>>>(nstreams, nsteps, d_ret, d->pcm_len, d_mark, mark_stride);
    if ((rc = post_launch(c))) return rc;
  }
  if ((rc = plan_launch(c, nstreams, d_mark, mark_stride, nsteps, d->pcm_len, d->eof, d->max_blocks, d->plan, d->nblocks,
                        d->cap, d_src, d_desc, d_tot, st))) return rc;
  int tot[2];
  CU(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));                       // the launches below are sized by the plan
  d->count[0] = tot[0]; d->count[1] = tot[1];
  if (tot[0] > d->cap[0] || tot[1] > d->cap[1]) return fail(VB200_EINVAL, "encode_streams: more blocks than cap[] (count[] holds the need)");
  // 4. transforms of both sizes, 5. the ampmax chain along every stream, 6. the rest of the chain per size
  EncScratch S[2];
  vb200_phaseA_io a[2];
  for (int w = 0; w < 2; w++) {
    memset(&a[w], 0, sizeof(a[w]));
    if (!tot[w]) continue;
    const size_t n = c->dx[w].N / 2, nb = tot[w];
    if ((rc = enc_scratch(c->str_buf + 12 + 8 * w, nb * ch, nb, n, nullptr, false, &S[w]))) return rc;
    a[w].desc = d_desc[w]; a[w].mdct = S[w].mdct; a[w].logmdct = S[w].logmdct; a[w].logmask = S[w].logmask;
    a[w].ampmax_out = d->ampmax_out[w];
    PcmSrc ps; ps.base = d->pcm; ps.fmt = d->pcm_fmt; ps.bps = 1; ps.hop = 0; ps.stride = d->stream_stride; ps.blk_src = d_src[w];
    if ((rc = phaseA_transform_launch(c, w, (int)nb, &a[w], ps, st, S[w].logfft, S[w].lmax))) return rc;
  }
  {
    float sa[2];
    for (int w = 0; w < 2; w++) sa[w] = ((float)(c->dx[w].N / 2) / (float)c->setup.rate) * c->setup.ampmax_att_per_sec;   // lib/psy.c:843
    k_ampmax_plan<<<(nstreams + 127) / 128, 128, 0, st>>>(nstreams, d->max_blocks, ch, d->plan, d->nblocks,
                                                          tot[0] ? S[0].lmax : nullptr, tot[1] ? S[1].lmax : nullptr,
                                                          sa[0], sa[1], tot[0] ? S[0].gmax : nullptr, tot[1] ? S[1].gmax : nullptr);
    if ((rc = post_launch(c))) return rc;
  }
  for (int w = 0; w < 2; w++) {
    if (!tot[w]) continue;
    const int nb = tot[w], rows = nb * ch;
    if ((rc = phaseA_psy_launch(c, w, nb, &a[w], st, S[w].logfft, S[w].lmax, S[w].gmax))) return rc;
    if ((rc = vb200_floor1_fit_dev(c, w, -1, rows, S[w].logmdct, S[w].logmask, d->posts[w], S[w].fitnz, st))) return rc;
    if ((rc = vb200_floor1_render_dev(c, w, -1, rows, d->posts[w], S[w].fitnz, d->iwork[w], d->nonzero[w], st))) return rc;
    CqnDev Q0, Q1;
    if ((rc = cqn_setup(c, w, 0, blobno, &Q0))) return rc;
    if ((rc = cqn_setup(c, w, 1, blobno, &Q1))) return rc;
    if ((rc = cqn_launch(c, Q0, Q1, d_desc[w], nb, S[w].mdct, d->iwork[w], d->nonzero[w], st))) return rc;
  }
  return scratch_end(c, st);
}

extern "C" int vb200_encode_streams(vb200_ctx *c, int nstreams, int blobno, vb200_streams_io *h) {
  CHECK_CTX(c);
  int rc;
  if ((rc = plan_check(c))) return rc;
  if (!h) return fail(VB200_EINVAL, "null io");
  h->count[0] = h->count[1] = 0;
  if (nstreams <= 0) return 0;
  if (!h->pcm || !h->pcm_len || !h->plan || !h->nblocks || h->max_blocks < 1) return fail(VB200_EINVAL, "encode_streams: pcm, pcm_len, plan, nblocks");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t st = c->s_main;
  const size_t ch = c->setup.channels;
  const size_t pcm_bytes = (size_t)nstreams * ch * (size_t)h->stream_stride * (h->pcm_fmt == VB200_PCM_S16_INTERLEAVED ? 2 : 4);
  const size_t pcap = (size_t)nstreams * h->max_blocks;
  void *p;
  vb200_streams_io d = *h;
  if ((rc = ensure_buf(c->str_buf[28], pcm_bytes, &p))) return rc; d.pcm = p;
  if ((rc = ensure_buf(c->str_buf[3], sizeof(int64_t) * 2 * (size_t)nstreams, &p))) return rc;
  d.pcm_len = (int64_t *)p; d.eof = h->eof ? (int64_t *)p + nstreams : nullptr;
  if ((rc = ensure_buf(c->str_buf[4], sizeof(vb200_stream_block) * pcap, &p))) return rc; d.plan = (vb200_stream_block *)p;
  if ((rc = ensure_buf(c->str_buf[29], sizeof(int32_t) * nstreams, &p))) return rc; d.nblocks = (int32_t *)p;
  for (int w = 0; w < 2; w++) {
    const size_t cw = h->cap[w] > 0 ? h->cap[w] : 0, n = c->dx[w].N / 2;
    if (!cw) continue;
    if ((rc = ensure_buf(c->str_buf[30 + 4 * w], sizeof(int32_t) * cw * ch * VB200_FLOOR1_STRIDE, &p))) return rc; d.posts[w] = (int32_t *)p;
    if ((rc = ensure_buf(c->str_buf[31 + 4 * w], sizeof(int32_t) * cw * ch, &p))) return rc; d.nonzero[w] = (int32_t *)p;
    if ((rc = ensure_buf(c->str_buf[32 + 4 * w], sizeof(int32_t) * cw * ch * n, &p))) return rc; d.iwork[w] = (int32_t *)p;
    if ((rc = ensure_buf(c->str_buf[33 + 4 * w], sizeof(float) * cw, &p))) return rc; d.ampmax_out[w] = (float *)p;
  }
  CU(cudaMemcpyAsync((void *)d.pcm, h->pcm, pcm_bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync((void *)d.pcm_len, h->pcm_len, sizeof(int64_t) * nstreams, cudaMemcpyHostToDevice, st));
  if (h->eof) CU(cudaMemcpyAsync((void *)d.eof, h->eof, sizeof(int64_t) * nstreams, cudaMemcpyHostToDevice, st));
  rc = vb200_encode_streams_dev(c, nstreams, blobno, &d, st);
  h->count[0] = d.count[0]; h->count[1] = d.count[1];
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->plan, d.plan, sizeof(vb200_stream_block) * pcap, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->nblocks, d.nblocks, sizeof(int32_t) * nstreams, cudaMemcpyDeviceToHost, st));
  for (int w = 0; w < 2; w++) {
    const size_t nb = d.count[w], n = c->dx[w].N / 2;
    if (!nb) continue;
    CU(cudaMemcpyAsync(h->posts[w], d.posts[w], sizeof(int32_t) * nb * ch * VB200_FLOOR1_STRIDE, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->nonzero[w], d.nonzero[w], sizeof(int32_t) * nb * ch, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->iwork[w], d.iwork[w], sizeof(int32_t) * nb * ch * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->ampmax_out[w], d.ampmax_out[w], sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return 0;
}

// ======================================================================== //
// residue partition classification
extern "C" int vb200_residue_partvals(vb200_ctx *c, int W) {
  if (!c || W < 0 || W > 1) return VB200_EINVAL;
  return c->res_partvals[W];
}

extern "C" int vb200_residue_classify_dev(vb200_ctx *c, int W, int nblocks, const int32_t *d_iwork,
                                          const int32_t *d_nonzero, int32_t *d_classes, int class_stride, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0) return 0;
  if (!d_iwork || !d_nonzero || !d_classes) return fail(VB200_EINVAL, "residue_classify pointers");
  if (c->res_partvals[W] <= 0) return fail(VB200_EINVAL, "no residue setup for this block size");
  if (class_stride < c->res_partvals[W]) return fail(VB200_EINVAL, "class_stride < partvals");
  const int ch = c->setup.channels, n = c->dx[W].N / 2, submaps = c->setup.submaps[W] > 0 ? c->setup.submaps[W] : 1;
  for (int sm = 0; sm < submaps; sm++) {                   // the reads must stay inside the rows
    const vb200_residue_setup &r = c->setup.residue[W][sm];
    if (r.type < 0 || r.grouping <= 0) continue;
    int cib = 0;
    for (int k = 0; k < ch; k++) if (c->setup.chmux[W][k] == sm) cib++;
    const long reach = r.type == 2 ? (cib ? r.begin / cib + (long)((r.end - r.begin) / r.grouping) * ((r.grouping + cib - 1) / cib) : 0)
                                   : (long)r.begin + (long)((r.end - r.begin) / r.grouping) * r.grouping;
    if (reach > n) return fail(VB200_EINVAL, "residue range exceeds the block");
  }
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemsetAsync(d_classes, 0, sizeof(int32_t) * (size_t)nblocks * ch * class_stride, st));
  ResArgs A;
  A.res = c->d_res[W]; A.chmux = c->d_chmux[W]; A.ch = ch; A.n = n; A.submaps = submaps; A.nblocks = nblocks;
  A.stride = class_stride;
  const long tasks = (long)nblocks * submaps;
  k_residue_classify<<<grid_for(c, (int)((tasks + RES_WARPS - 1) / RES_WARPS), 16), 32 * RES_WARPS, 0, st>>>(
      A, d_iwork, d_nonzero, d_classes);
  return post_launch(c);
}

extern "C" int vb200_residue_classify(vb200_ctx *c, int W, int nblocks, const int32_t *iwork, const int32_t *nonzero,
                                      int32_t *classes, int class_stride) {
  CHECK_CTX(c); CHECK_W(W);
  if (nblocks <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t ch = c->setup.channels, n = c->dx[W].N / 2;
  HostIO io{c};
  void *di, *dz, *dc; int rc;
  if ((rc = io.h2d(iwork, sizeof(int32_t) * nblocks * ch * n, &di))) return rc;
  if ((rc = io.h2d(nonzero, sizeof(int32_t) * nblocks * ch, &dz))) return rc;
  if ((rc = io.h2d(nullptr, sizeof(int32_t) * nblocks * ch * (size_t)(class_stride > 0 ? class_stride : 1), &dc))) return rc;
  if ((rc = vb200_residue_classify_dev(c, W, nblocks, (const int32_t *)di, (const int32_t *)dz, (int32_t *)dc,
                                       class_stride, c->s_main))) return rc;
  if ((rc = io.d2h(classes, dc, sizeof(int32_t) * nblocks * ch * (size_t)class_stride))) return rc;
  return io.sync();
}

// ======================================================================== //
// decode: floor multiply and the fused decode chain
extern "C" int vb200_floor1_inverse2_dev(vb200_ctx *c, int W, int floor_sel, int nrows, const int32_t *d_posts,
                                         const int32_t *d_present, float *d_data, void *stream) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  if (!d_posts || !d_present || !d_data) return fail(VB200_EINVAL, "floor1_inverse2 pointers");
  Floor1Args a; int rc;
  if ((rc = floor1_args(c, W, floor_sel, nrows, &a))) return rc;
  const int ctas = (nrows + F1_WARPS - 1) / F1_WARPS;
  k_floor1_inverse2<<<grid_for(c, ctas, 8), 32 * F1_WARPS, 0, (cudaStream_t)stream>>>(a, d_posts, d_present, d_data, c->d_fromdB);
  return post_launch(c);
}

extern "C" int vb200_floor1_inverse2(vb200_ctx *c, int W, int floor_sel, int nrows, const int32_t *posts,
                                     const int32_t *present, float *data) {
  CHECK_CTX(c); CHECK_W(W);
  if (nrows <= 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t n = c->dx[W].N / 2;
  HostIO io{c};
  void *dp, *dz, *dd; int rc;
  if ((rc = io.h2d(posts, sizeof(int32_t) * (size_t)nrows * VB200_FLOOR1_STRIDE, &dp))) return rc;
  if ((rc = io.h2d(present, sizeof(int32_t) * (size_t)nrows, &dz))) return rc;
  if ((rc = io.h2d(data, sizeof(float) * nrows * n, &dd))) return rc;
  if ((rc = vb200_floor1_inverse2_dev(c, W, floor_sel, nrows, (const int32_t *)dp, (const int32_t *)dz, (float *)dd, c->s_main))) return rc;
  if ((rc = io.d2h(data, dd, sizeof(float) * nrows * n))) return rc;
  return io.sync();
}

extern "C" int vb200_decode_dsp_dev(vb200_ctx *c, int nstreams, int nblk, const int32_t *d_Wseq, const int64_t *d_coef_off,
                                    float *d_res, const int32_t *d_posts, const int32_t *d_present,
                                    const int64_t *d_pcm_off, void *d_pcm, int pcm_s16, int64_t pcm_stride, void *stream) {
  CHECK_CTX(c);
  if (nstreams <= 0 || nblk <= 0) return 0;
  if (!d_Wseq || !d_coef_off || !d_res || !d_posts || !d_present || !d_pcm_off || !d_pcm)
    return fail(VB200_EINVAL, "decode pointers");
  const int ch = c->setup.channels;
  DecodePrepArgs A;
  for (int w = 0; w < 2; w++) {
    for (int k = 0; k < ch; k++)
      if (c->setup.floor1[w][c->setup.chmux[w][k]].posts <= 0)
        return fail(VB200_EINVAL, "no floor1 setup for a channel's submap");
    A.floors[w] = c->d_floor[w]; A.chmux[w] = c->d_chmux[w];
    A.mag[w] = c->d_mag[w]; A.ang[w] = c->d_ang[w];
    A.steps[w] = c->setup.coupling_steps[w]; A.n[w] = c->dx[w].N / 2;
  }
  A.ch = ch; A.nblk = nblk; A.nitems = (long)nstreams * nblk;
  int rc;
  k_decode_prepare<<<grid_for(c, (int)A.nitems, 8), 128, 0, (cudaStream_t)stream>>>(
      A, d_Wseq, (const long long *)d_coef_off, d_res, d_posts, d_present, c->d_fromdB);
  if ((rc = post_launch(c))) return rc;
  if (pcm_s16) return vb200_synthesis_s16_dev(c, nstreams, nblk, d_Wseq, d_coef_off, d_res, d_pcm_off, (int16_t *)d_pcm, pcm_stride, stream);
  return vb200_synthesis_dev(c, nstreams, nblk, d_Wseq, d_coef_off, d_res, d_pcm_off, (float *)d_pcm, pcm_stride, stream);
}

extern "C" int vb200_decode_dsp(vb200_ctx *c, int nstreams, int nblk, const int32_t *Wseq, const int64_t *coef_off,
                                float *res, int64_t res_len, const int32_t *posts, const int32_t *present,
                                const int64_t *pcm_off, void *pcm, int pcm_s16, int64_t pcm_stride) {
  CHECK_CTX(c);
  if (nstreams <= 0 || nblk <= 0) return 0;
  if (!Wseq || !coef_off || !res || !posts || !present || !pcm_off || !pcm) return fail(VB200_EINVAL, "decode pointers");
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels;
  const size_t nb = (size_t)nstreams * nblk;
  for (size_t i = 0; i < nb; i++) if (Wseq[i] < 0 || Wseq[i] > 1) return fail(VB200_EINVAL, "Wseq values must be 0/1");
  HostIO io{c};
  void *dW, *dco, *dc, *dpo, *dp, *dps, *dpr; int rc;
  if ((rc = io.h2d(Wseq, sizeof(int32_t) * nb, &dW))) return rc;
  if ((rc = io.h2d(coef_off, sizeof(int64_t) * nb, &dco))) return rc;
  if ((rc = io.h2d(res, sizeof(float) * (size_t)res_len, &dc))) return rc;
  if ((rc = io.h2d(pcm_off, sizeof(int64_t) * nb, &dpo))) return rc;
  const size_t pbytes = (pcm_s16 ? sizeof(int16_t) : sizeof(float)) * (size_t)nstreams * ch * (size_t)pcm_stride;
  if ((rc = io.h2d(nullptr, pbytes, &dp))) return rc;
  if ((rc = io.h2d(posts, sizeof(int32_t) * nb * ch * VB200_FLOOR1_STRIDE, &dps))) return rc;
  if ((rc = io.h2d(present, sizeof(int32_t) * nb * ch, &dpr))) return rc;
  CU(cudaMemsetAsync(dp, 0, pbytes, c->s_main));
  if ((rc = vb200_decode_dsp_dev(c, nstreams, nblk, (const int32_t *)dW, (const int64_t *)dco, (float *)dc,
                                 (const int32_t *)dps, (const int32_t *)dpr, (const int64_t *)dpo, dp, pcm_s16,
                                 pcm_stride, c->s_main))) return rc;
  if ((rc = io.d2h(pcm, dp, pbytes))) return rc;
  return io.sync();
}

// ======================================================================== //
// envelope / block-switch detector
static int env_check(vb200_ctx *c, int nstreams, const void *pcm, int fmt, int64_t stride, int first, int nsteps,
                     const void *state, const void *ret) {
  if (!pcm || !state || !ret) return fail(VB200_EINVAL, "envelope pointers");
  if (nstreams <= 0 || nsteps < 0 || first < 0) return fail(VB200_EINVAL, "nstreams/steps");
  if (fmt != VB200_PCM_F32_PLANAR && fmt != VB200_PCM_S16_INTERLEAVED) return fail(VB200_EINVAL, "pcm format");
  if ((int64_t)ENV_STEP * ((int64_t)first + nsteps - 1) + ENV_N > stride && nsteps > 0)
    return fail(VB200_EINVAL, "steps exceed the stream buffer");
  (void)c;
  return 0;
}

static int envelope_search_launch(vb200_ctx *c, int nstreams, const void *d_pcm, int fmt, int64_t stride,
                                  int first_step, int nsteps, int32_t *d_state, uint8_t *d_ret, void *stream,
                                  const int32_t *d_steps_per_stream) {
  CHECK_CTX(c);
  int rc;
  if ((rc = env_check(c, nstreams, d_pcm, fmt, stride, first_step, nsteps, d_state, d_ret))) return rc;
  if (nsteps == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int ch = c->setup.channels;
  // bounded scratch: the spectra of at most ~4M (stream, channel, step) items at a time
  long per_step = (long)nstreams * ch;
  int chunk = (int)((1L << 22) / per_step);
  if (chunk < 1) chunk = 1;
  if (chunk > nsteps) chunk = nsteps;
  void *p_t, *p_v;
  if ((rc = ensure_buf(c->env_buf[0], sizeof(float) * (size_t)per_step * chunk, &p_t))) return rc;
  if ((rc = ensure_buf(c->env_buf[1], sizeof(float) * (size_t)per_step * chunk * 32, &p_v))) return rc;
  EnvSrc src; src.base = d_pcm; src.fmt = fmt; src.stride = stride; src.ch = ch;
  if ((rc = scratch_begin(c, st))) return rc;
  for (int j0 = 0; j0 < nsteps; j0 += chunk) {
    const int ns = nsteps - j0 < chunk ? nsteps - j0 : chunk;
    const long items = per_step * ns;
    const int grid = grid_for(c, (int)((items + ENV_WARPS - 1) / ENV_WARPS), 16);
    k_env_spectrum<<<grid, 32 * ENV_WARPS, 0, st>>>(c->env, src, nstreams, first_step + j0, ns, (float *)p_t, (float *)p_v);
    if ((rc = post_launch(c))) return rc;
    k_env_filter<<<(nstreams + ENV_WARPS - 1) / ENV_WARPS, 32 * ENV_WARPS, 0, st>>>(
        c->env, nstreams, ch, ns, nsteps, j0, (const float *)p_t, (const float *)p_v, d_state, d_ret, d_steps_per_stream);
    if ((rc = post_launch(c))) return rc;
  }
  return scratch_end(c, st);
}

extern "C" int vb200_envelope_search_dev(vb200_ctx *c, int nstreams, const void *d_pcm, int fmt, int64_t stride,
                                         int first_step, int nsteps, int32_t *d_state, uint8_t *d_ret, void *stream) {
  return envelope_search_launch(c, nstreams, d_pcm, fmt, stride, first_step, nsteps, d_state, d_ret, stream, nullptr);
}

static int envelope_search_host(vb200_ctx *c, int nstreams, const void *pcm, int fmt, int64_t stride,
                                int first_step, int nsteps, int32_t *state, uint8_t *ret, const int32_t *steps_per_stream) {
  CHECK_CTX(c);
  int rc;
  if ((rc = env_check(c, nstreams, pcm, fmt, stride, first_step, nsteps, state, ret))) return rc;
  if (nsteps == 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const int ch = c->setup.channels;
  const size_t pcm_bytes = (fmt == VB200_PCM_S16_INTERLEAVED ? sizeof(int16_t) : sizeof(float)) * (size_t)nstreams * ch * (size_t)stride;
  const size_t st_bytes = sizeof(int32_t) * (size_t)nstreams * VB200_VE_STATE_WORDS(ch);
  void *dp, *ds, *dr;
  if ((rc = ensure_buf(c->env_buf[2], pcm_bytes, &dp))) return rc;
  if ((rc = ensure_buf(c->env_buf[3], st_bytes + (size_t)nstreams * nsteps + sizeof(int32_t) * (size_t)nstreams + 16, &ds))) return rc;
  dr = (char *)ds + st_bytes;
  int32_t *dn = nullptr;
  if (steps_per_stream) {
    dn = (int32_t *)((char *)dr + (((size_t)nstreams * nsteps + 15) & ~(size_t)15));
    CU(cudaMemcpyAsync(dn, steps_per_stream, sizeof(int32_t) * (size_t)nstreams, cudaMemcpyHostToDevice, c->s_main));
  }
  CU(cudaMemcpyAsync(dp, pcm, pcm_bytes, cudaMemcpyHostToDevice, c->s_main));
  CU(cudaMemcpyAsync(ds, state, st_bytes, cudaMemcpyHostToDevice, c->s_main));
  if ((rc = envelope_search_launch(c, nstreams, dp, fmt, stride, first_step, nsteps, (int32_t *)ds, (uint8_t *)dr, c->s_main, dn))) return rc;
  CU(cudaMemcpyAsync(state, ds, st_bytes, cudaMemcpyDeviceToHost, c->s_main));
  CU(cudaMemcpyAsync(ret, dr, (size_t)nstreams * nsteps, cudaMemcpyDeviceToHost, c->s_main));
  CU(cudaStreamSynchronize(c->s_main));
  return 0;
}

extern "C" int vb200_envelope_search(vb200_ctx *c, int nstreams, const void *pcm, int fmt, int64_t stride,
                                     int first_step, int nsteps, int32_t *state, uint8_t *ret) {
  return envelope_search_host(c, nstreams, pcm, fmt, stride, first_step, nsteps, state, ret, nullptr);
}

// streams with different amounts of new data in one call: stream s analyses its first steps_per_stream[s]
// (<= nsteps) steps, its later ret entries are left untouched
extern "C" int vb200_envelope_search_var(vb200_ctx *c, int nstreams, const void *pcm, int fmt, int64_t stride,
                                         int nsteps, const int32_t *steps_per_stream, int32_t *state, uint8_t *ret) {
  if (!steps_per_stream) return fail(VB200_EINVAL, "steps_per_stream");
  return envelope_search_host(c, nstreams, pcm, fmt, stride, 0, nsteps, state, ret, steps_per_stream);
}

// lib/envelope.c:254-264, replayed on the host from the per-step trigger bits (plain C, no CUDA)
extern "C" void vb200_envelope_apply_marks(const uint8_t *ret, int first_step, int nsteps, int32_t *mark) {
  for (int k = 0; k < nsteps; k++) {
    const int j = first_step + k;
    mark[j + 2] = 0;                                    // VE_POST
    if (ret[k] & 1) { mark[j] = 1; mark[j + 1] = 1; }
    if (ret[k] & 2) { mark[j] = 1; if (j > 0) mark[j - 1] = 1; }
  }
}

// ======================================================================== //
// device memory helpers
extern "C" int vb200_malloc_device(vb200_ctx *c, size_t bytes, void **dptr) {
  CHECK_CTX(c);
  CU(cudaMalloc(dptr, bytes ? bytes : 1));
  return 0;
}
extern "C" int vb200_free_device(vb200_ctx *c, void *dptr) { CHECK_CTX(c); CU(cudaFree(dptr)); return 0; }
extern "C" int vb200_memcpy_h2d(vb200_ctx *c, void *dptr, const void *src, size_t bytes) {
  CHECK_CTX(c); CU(cudaMemcpy(dptr, src, bytes, cudaMemcpyHostToDevice)); return 0;
}
extern "C" int vb200_memcpy_d2h(vb200_ctx *c, void *dst, const void *dptr, size_t bytes) {
  CHECK_CTX(c); CU(cudaMemcpy(dst, dptr, bytes, cudaMemcpyDeviceToHost)); return 0;
}
extern "C" int vb200_synchronize(vb200_ctx *c) { CHECK_CTX(c); CU(cudaDeviceSynchronize()); return 0; }
