// vb200_tables.h — host-side construction of the lookup tables the kernels read.
//
// Product code (no oracle/ dependency).  The transform tables restate what the
// reference derives at init time; citations are to the xiph/vorbis tree
// (libvorbis 1.3.7).  The psychoacoustic lookups are NOT rebuilt here: they are
// uploaded as given in vb200_setup (built by the reference's _vp_psy_init,
// lib/psy.c:266); only data-independent control-flow tables are derived.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "vorbis_b200.h"

namespace vb200 {

struct HostXform {
  int N = 0, log2n = 0;
  std::vector<float> trig;      // N + N/4   (mdct_init, lib/mdct.c:51-90)
  std::vector<int>   bitrev;    // N/4
  std::vector<float> stage_tw;  // per radix-2 stage s: (N/8 >> s) float2, contiguous
  std::vector<int>   stage_off; // float2 offset of stage s in stage_tw
  std::vector<float> win;       // N/2 rising half window (lib/window.c tables)
  std::vector<float> wa;        // N   real-FFT twiddles (drfti1, lib/smallft.c:37-108)
  int nf = 0;
  int fac[16] = {0};            // factors in the order drfti1 stores them
  float scale = 0.f;
};

inline void build_xform(HostXform &x, int N, const float *given_window) {
  x.N = N;
  x.log2n = (int)std::rint(std::log((double)(float)N) / std::log((double)2.f));
  const int n2 = N >> 1;
  x.trig.assign(N + N / 4, 0.f);
  x.bitrev.assign(N / 4, 0);
  for (int i = 0; i < N / 4; i++) {
    x.trig[i * 2]          = (float)std::cos((M_PI / N) * (4 * i));
    x.trig[i * 2 + 1]      = (float)-std::sin((M_PI / N) * (4 * i));
    x.trig[n2 + i * 2]     = (float)std::cos((M_PI / (2 * N)) * (2 * i + 1));
    x.trig[n2 + i * 2 + 1] = (float)std::sin((M_PI / (2 * N)) * (2 * i + 1));
  }
  for (int i = 0; i < N / 8; i++) {
    x.trig[N + i * 2]     = (float)(std::cos((M_PI / N) * (4 * i + 2)) * .5);
    x.trig[N + i * 2 + 1] = (float)(-std::sin((M_PI / N) * (4 * i + 2)) * .5);
  }
  {
    const int mask = (1 << (x.log2n - 1)) - 1;
    const int msb = 1 << (x.log2n - 2);
    for (int i = 0; i < N / 8; i++) {
      int acc = 0;
      for (int j = 0; msb >> j; j++)
        if ((msb >> j) & i) acc |= 1 << j;
      x.bitrev[i * 2] = ((~acc) & mask) - 1;
      x.bitrev[i * 2 + 1] = acc;
    }
  }
  x.scale = 4.f / N;

  // Radix-2 stage s of mdct_butterflies (lib/mdct.c:316-336) reads trig[q*(4<<s)]
  // and trig[q*(4<<s)+1] for q < (N/2 >> s)/4.  Repack them densely per stage so a
  // warp reads consecutive float2 instead of a 16B<<s stride.
  const int nst = x.log2n - 6;
  x.stage_off.clear();
  x.stage_tw.clear();
  for (int s = 0; s < nst; s++) {
    const int P = n2 >> s, stride = 4 << s;
    x.stage_off.push_back((int)(x.stage_tw.size() / 2));
    for (int q = 0; q < P / 4; q++) {
      x.stage_tw.push_back(x.trig[q * stride]);
      x.stage_tw.push_back(x.trig[q * stride + 1]);
    }
  }
  if (x.stage_tw.empty()) x.stage_tw.assign(2, 0.f);

  // window: tabulated by the reference (lib/window.c:23-2096); closed form of
  // doc/04-codec.tex:320 as fallback (differs from the table by <=1 ulp in a few
  // entries, so drop-in callers pass the table).
  x.win.assign(N / 2, 0.f);
  for (int i = 0; i < N / 2; i++) {
    if (given_window) x.win[i] = given_window[i];
    else {
      const double s = std::sin((i + .5) / N * M_PI);
      x.win[i] = (float)std::sin(M_PI * .5 * s * s);
    }
  }

  // real FFT: factors 4,..,4 with one 2 (if any) moved to the front; twiddles
  // formed in float exactly as drfti1 does (argh, argld, arg are floats).
  int nl = N, nf = 0;
  while (nl % 4 == 0) { x.fac[nf++] = 4; nl /= 4; }
  if (nl == 2) {
    for (int j = nf; j > 0; j--) x.fac[j] = x.fac[j - 1];
    x.fac[0] = 2; nf++; nl = 1;
  }
  x.nf = nf;
  x.wa.assign(N, 0.f);
  {
    const float tpi = 6.28318530717958648f;
    const float argh = tpi / N;
    int l1 = 1, is = 0;
    for (int k1 = 0; k1 < nf - 1; k1++) {
      const int ip = x.fac[k1];
      int ld = 0;
      const int l2 = l1 * ip, ido = N / l2;
      for (int j = 0; j < ip - 1; j++) {
        int i = is;
        float fi = 0.f;
        ld += l1;
        const float argld = (float)ld * argh;
        for (int ii = 2; ii < ido; ii += 2) {
          fi += 1.f;
          const float arg = fi * argld;
          x.wa[i++] = (float)std::cos((double)arg);   // double libm call, as C's cos(float) promotes
          x.wa[i++] = (float)std::sin((double)arg);
        }
        is += ido;
      }
      l1 = l2;
    }
  }
}

// Data-independent control flow of the psy stages, derived once per look.
struct HostPsyFlow {
  std::vector<int> run_lo, run_hi;             // seed_loop runs (lib/psy.c:430-436)
  std::vector<int> runinfo;                    // per run: lo, hi, octave[hi]-firstoc, band
  std::vector<int> cls_run;                    // (run id, oc) pairs grouped by residue class of the seed slots
  std::vector<int> slot_rng;                   // per seed slot: [k0,k1) into cls_run (seed_curve, lib/psy.c:390-415)
  int linesper_log2 = 0;
  std::vector<int> cls_off;                    // [L+1]
  std::vector<int> grp;                        // max_seeds groups: pos0,pos1,lin0,lin1 (lib/psy.c:522-538)
  int tail_lin0 = 0;
  std::vector<int> runrec;                     // per class-ordered run: lo|hi<<16, oc-firstoc, band, bits(ath[hi])
  std::vector<int> long_grp;                   // groups whose seed range is longer than 16 (folded by a whole warp)
  std::vector<short> bin_grp;                  // per bin: its max_seeds group, ngrp for the tail bins
  int bark_first_extra = 0;                    // first bin that reuses the last A,B,D (lib/psy.c:604-658)
  int fixed_first_extra = 0;                   // same for the fixed window (lib/psy.c:660-703)
};

inline void build_psy_flow(HostPsyFlow &f, const vb200_psy_setup &s) {
  const int n = s.n;
  f.run_lo.clear(); f.run_hi.clear(); f.grp.clear();
  for (int i = 0; i < n;) {
    int j = i;
    while (j + 1 < n && s.octave[j + 1] == s.octave[i]) j++;
    f.run_lo.push_back(i); f.run_hi.push_back(j);
    i = j + 1;
  }
  {
    // owner-computes form of seed_curve's scatter: slot sp = oc + (i-16)*L - L/2, i in [post0,post1) within [0,56)
    const int L = s.eighth_octave_lines, half = L >> 1, total = s.total_octave_lines;
    f.linesper_log2 = 0;
    while ((1 << f.linesper_log2) < L) f.linesper_log2++;
    f.runinfo.clear(); f.cls_run.clear(); f.slot_rng.assign((size_t)2 * total, 0);
    const int nr = (int)f.run_lo.size();
    std::vector<int> oc(nr);
    for (int r = 0; r < nr; r++) {
      const int ov = s.octave[f.run_hi[r]];
      int band = ov >> s.shiftoc;              // arithmetic shift, as the reference's long >> (lib/psy.c:438)
      if (band >= VB200_P_BANDS) band = VB200_P_BANDS - 1;
      if (band < 0) band = 0;
      oc[r] = ov - s.firstoc;
      f.runinfo.push_back(f.run_lo[r]); f.runinfo.push_back(f.run_hi[r]);
      f.runinfo.push_back(oc[r]); f.runinfo.push_back(band);
    }
    std::vector<int> &cls_off = f.cls_off;
    cls_off.assign(L + 1, 0);
    for (int c = 0; c < L; c++) {
      cls_off[c] = (int)f.cls_run.size() / 2;
      for (int r = 0; r < nr; r++)             // runs are already in increasing oc order
        if ((((oc[r] - half) % L) + L) % L == c) { f.cls_run.push_back(r); f.cls_run.push_back(oc[r]); }
    }
    cls_off[L] = (int)f.cls_run.size() / 2;
    f.runrec.clear();
    for (size_t k = 0; k < f.cls_run.size() / 2; k++) {
      const int r = f.cls_run[2 * k];
      union { float fl; int i; } u; u.fl = s.ath[f.run_hi[r]];
      f.runrec.push_back(f.run_lo[r] | (f.run_hi[r] << 16));
      f.runrec.push_back(f.runinfo[4 * r + 2]); f.runrec.push_back(f.runinfo[4 * r + 3]); f.runrec.push_back(u.i);
    }
    for (int sp = 1; sp < total; sp++) {
      const int c = sp % L;
      int k0 = cls_off[c + 1], k1 = cls_off[c];
      for (int k = cls_off[c]; k < cls_off[c + 1]; k++) {
        const int o = f.cls_run[2 * k + 1];
        if (o >= sp + half - 39 * L && o <= sp + half + 16 * L) { if (k < k0) k0 = k; if (k + 1 > k1) k1 = k + 1; }
      }
      if (k1 < k0) { k0 = 0; k1 = 0; }
      f.slot_rng[2 * sp] = k0; f.slot_rng[2 * sp + 1] = k1;
    }
  }
  {
    long linpos = 0;
    long pos = s.octave[0] - s.firstoc - (s.eighth_octave_lines >> 1);
    while (linpos + 1 < n) {
      long end = ((s.octave[linpos] + s.octave[linpos + 1]) >> 1) - s.firstoc;
      f.grp.push_back((int)pos);
      while (pos + 1 <= end) pos++;
      f.grp.push_back((int)pos);
      end = pos + s.firstoc;
      f.grp.push_back((int)linpos);
      for (; linpos < n && s.octave[linpos] <= end; linpos++) {}
      f.grp.push_back((int)linpos);
    }
    f.tail_lin0 = (int)linpos;
    const int ng = (int)f.grp.size() / 4;
    f.long_grp.clear();
    for (int gi = 0; gi < ng; gi++) if (f.grp[4 * gi + 1] - f.grp[4 * gi] > 16) f.long_grp.push_back(gi);
    f.bin_grp.assign(n, (short)ng);
    for (int gi = 0; gi < ng; gi++)
      for (int i = f.grp[4 * gi + 2]; i < f.grp[4 * gi + 3]; i++) f.bin_grp[i] = (short)gi;
  }
  {
    int i = 0;
    for (; i < n; i++) {
      const int lo = s.bark[i] >> 16, hi = s.bark[i] & 0xffff;
      if (lo >= 0 || -lo >= n || hi >= n) break;
    }
    for (; i < n; i++) {
      const int lo = s.bark[i] >> 16, hi = s.bark[i] & 0xffff;
      if (lo < 0 || lo >= n || hi >= n) break;
    }
    f.bark_first_extra = i;
  }
  {
    const int fixed = s.noisewindowfixed;
    int i = 0;
    if (fixed > 0) {
      for (; i < n; i++) {
        const int hi = i + fixed / 2, lo = hi - fixed;
        if (hi >= n || lo >= 0) break;
      }
      for (; i < n; i++) {
        const int hi = i + fixed / 2, lo = hi - fixed;
        if (hi >= n || lo < 0) break;
      }
    }
    f.fixed_first_extra = i;
  }
}

}  // namespace vb200
