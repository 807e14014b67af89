/* vb_oracle.c — CPU oracle for the libvorbis per-block DSP path.
 *
 * TEST INFRASTRUCTURE ONLY (see vb_oracle.h).  A from-scratch restatement in
 * plain C of what the reference computes, organised as loops over independent
 * work items ("for item u in stage s") instead of the reference's pointer
 * walks.  Every function cites the reference file:line it follows.  The
 * arithmetic DAG per output value (which products, which sums, in fp32 or
 * fp64) is kept identical to the reference so results are bit-identical when
 * both are built with -O2 -ffp-contract=off (tests/test_oracle_vs_ref.py).
 *
 * Paths are relative to the xiph/vorbis tree (libvorbis 1.3.7).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "vb_oracle.h"

#define NEGINF (-9999.f)          /* lib/psy.c:31 */

typedef struct {
  int    N, log2n;
  float *trig;                    /* N + N/4, as mdct_init lib/mdct.c:51-90 */
  int   *bitrev;                  /* N/4 */
  float  scale;
  float *win;                     /* N/2 rising half window */
  /* real FFT (lib/smallft.c) */
  int    nf, fac[32];
  float *wa;                      /* N twiddles, drfti1 lib/smallft.c:37-108 */
} vbo_xform;

typedef struct {
  vb200_psy_setup s;
  float *ath, *tonecurves, *noiseoffset;
  int32_t *octave, *bark;
  /* static structure of seed_loop (lib/psy.c:417-452): runs of equal octave */
  int nruns; int *run_lo, *run_hi;
  /* static structure of max_seeds' second loop (lib/psy.c:522-544) */
  int ngrp; int *grp_pos0, *grp_pos1, *grp_lin0, *grp_lin1; int tail_lin0;
} vbo_psy;

typedef struct {
  vb200_floor1_setup s;
  int quant_q;
  int sorted_index[VB200_VIF_POSIT + 2], forward_index[VB200_VIF_POSIT + 2], reverse_index[VB200_VIF_POSIT + 2];
  int hineighbor[VB200_VIF_POSIT], loneighbor[VB200_VIF_POSIT];
} vbo_floor1;
void vbo_floor1_init(vbo_floor1 *f, const vb200_floor1_setup *s);

struct vbo_ctx {
  vb200_setup setup;
  vbo_xform x[2];
  vbo_psy   psy[4];
  vbo_floor1 floor1[2][VB200_MAX_SUBMAPS];
};

/* ======================================================================= */
/* tables                                                                  */

/* mdct_init, lib/mdct.c:51-90 */
static void xform_init_mdct(vbo_xform *x, int N){
  int i, j;
  int n2 = N >> 1;
  x->N = N;
  x->log2n = (int)rint(log((float)N) / log(2.f));
  x->trig = (float*)malloc(sizeof(float) * (N + N/4));
  x->bitrev = (int*)malloc(sizeof(int) * (N/4));
  for(i = 0; i < N/4; i++){
    x->trig[i*2]        = (float)cos((M_PI / N) * (4*i));
    x->trig[i*2+1]      = (float)-sin((M_PI / N) * (4*i));
    x->trig[n2+i*2]     = (float)cos((M_PI / (2*N)) * (2*i+1));
    x->trig[n2+i*2+1]   = (float)sin((M_PI / (2*N)) * (2*i+1));
  }
  for(i = 0; i < N/8; i++){
    x->trig[N+i*2]      = (float)(cos((M_PI / N) * (4*i+2)) * .5);
    x->trig[N+i*2+1]    = (float)(-sin((M_PI / N) * (4*i+2)) * .5);
  }
  {
    int mask = (1 << (x->log2n - 1)) - 1;
    int msb  = 1 << (x->log2n - 2);
    for(i = 0; i < N/8; i++){
      int acc = 0;
      for(j = 0; msb >> j; j++)
        if((msb >> j) & i) acc |= 1 << j;
      x->bitrev[i*2]   = ((~acc) & mask) - 1;
      x->bitrev[i*2+1] = acc;
    }
  }
  x->scale = 4.f / N;
}

/* drfti1, lib/smallft.c:37-108, for n a power of two: factors 4,...,4 with a
 * single 2 (if any) moved to the front; twiddles in float arithmetic exactly
 * as the reference forms them (argh, argld, arg are floats).                */
static void xform_init_fft(vbo_xform *x, int N){
  int nl = N, nf = 0, k1, j, ii;
  int l1 = 1, is = 0;
  const float tpi = 6.28318530717958648f;
  float argh = tpi / N;
  while(nl % 4 == 0){ x->fac[nf++] = 4; nl /= 4; }
  if(nl == 2){
    for(j = nf; j > 0; j--) x->fac[j] = x->fac[j-1];
    x->fac[0] = 2; nf++; nl = 1;
  }
  x->nf = nf;
  x->wa = (float*)calloc((size_t)N, sizeof(float));
  for(k1 = 0; k1 < nf - 1; k1++){
    int ip = x->fac[k1];
    int ld = 0;
    int l2 = l1 * ip;
    int ido = N / l2;
    for(j = 0; j < ip - 1; j++){
      int i = is;
      float argld, fi = 0.f;
      ld += l1;
      argld = (float)ld * argh;
      for(ii = 2; ii < ido; ii += 2){
        float arg;
        fi += 1.f;
        arg = fi * argld;
        x->wa[i++] = (float)cos(arg);
        x->wa[i++] = (float)sin(arg);
      }
      is += ido;
    }
    l1 = l2;
  }
}

/* window closed form (doc/04-codec.tex:320); the reference tabulates it
 * (lib/window.c:23-2096) — callers normally pass that table in setup.window */
static void xform_init_window(vbo_xform *x, const float *given){
  int i, h = x->N / 2;
  x->win = (float*)malloc(sizeof(float) * h);
  for(i = 0; i < h; i++){
    if(given) x->win[i] = given[i];
    else{
      double s = sin((i + .5) / x->N * M_PI);   /* K = N: (p+.5)*pi/K ... */
      x->win[i] = (float)sin(M_PI * .5 * s * s);
    }
  }
}

static void *dup_mem(const void *p, size_t bytes){
  void *q = malloc(bytes ? bytes : 1);
  memcpy(q, p, bytes);
  return q;
}

static void psy_init(vbo_psy *p, const vb200_psy_setup *s){
  int n = s->n, i;
  p->s = *s;
  p->ath         = (float*)dup_mem(s->ath, sizeof(float)*n);
  p->octave      = (int32_t*)dup_mem(s->octave, sizeof(int32_t)*n);
  p->bark        = (int32_t*)dup_mem(s->bark, sizeof(int32_t)*n);
  p->tonecurves  = (float*)dup_mem(s->tonecurves, sizeof(float)*VB200_P_BANDS*VB200_P_LEVELS*(VB200_EHMER_MAX+2));
  p->noiseoffset = (float*)dup_mem(s->noiseoffset, sizeof(float)*VB200_P_NOISECURVES*n);
  p->s.ath = p->ath; p->s.octave = p->octave; p->s.bark = p->bark;
  p->s.tonecurves = p->tonecurves; p->s.noiseoffset = p->noiseoffset;

  /* runs of bins sharing one octave value: the outer loop of seed_loop,
   * lib/psy.c:430-436, depends only on the octave[] table */
  p->run_lo = (int*)malloc(sizeof(int)*n);
  p->run_hi = (int*)malloc(sizeof(int)*n);
  p->nruns = 0;
  for(i = 0; i < n; ){
    int j = i;
    while(j + 1 < n && p->octave[j+1] == p->octave[i]) j++;
    p->run_lo[p->nruns] = i; p->run_hi[p->nruns] = j; p->nruns++;
    i = j + 1;
  }

  /* the control flow of the gather loop in max_seeds, lib/psy.c:522-538,
   * also depends only on octave[]/firstoc: record, per outer iteration, the
   * seed positions it folds and the bins it updates */
  p->grp_pos0 = (int*)malloc(sizeof(int)*n);
  p->grp_pos1 = (int*)malloc(sizeof(int)*n);
  p->grp_lin0 = (int*)malloc(sizeof(int)*n);
  p->grp_lin1 = (int*)malloc(sizeof(int)*n);
  p->ngrp = 0;
  {
    long linpos = 0;
    long pos = p->octave[0] - s->firstoc - (s->eighth_octave_lines >> 1);
    while(linpos + 1 < n){
      long end = ((p->octave[linpos] + p->octave[linpos+1]) >> 1) - s->firstoc;
      int g = p->ngrp++;
      p->grp_pos0[g] = (int)pos;
      while(pos + 1 <= end) pos++;
      p->grp_pos1[g] = (int)pos;
      end = pos + s->firstoc;
      p->grp_lin0[g] = (int)linpos;
      for(; linpos < n && p->octave[linpos] <= end; linpos++);
      p->grp_lin1[g] = (int)linpos;
    }
    p->tail_lin0 = (int)linpos;
  }
}

static void psy_free(vbo_psy *p){
  free(p->ath); free(p->octave); free(p->bark); free(p->tonecurves); free(p->noiseoffset);
  free(p->run_lo); free(p->run_hi);
  free(p->grp_pos0); free(p->grp_pos1); free(p->grp_lin0); free(p->grp_lin1);
}

vbo_ctx *vbo_create(const vb200_setup *setup){
  vbo_ctx *c = (vbo_ctx*)calloc(1, sizeof(*c));
  int w, i;
  c->setup = *setup;
  for(w = 0; w < 2; w++){
    xform_init_mdct(&c->x[w], setup->blocksizes[w]);
    xform_init_fft(&c->x[w], setup->blocksizes[w]);
    xform_init_window(&c->x[w], setup->window[w]);
  }
  for(i = 0; i < setup->n_psy && i < 4; i++) psy_init(&c->psy[i], &setup->psy[i]);
  for(w = 0; w < 2; w++)
    for(i = 0; i < VB200_MAX_SUBMAPS; i++) vbo_floor1_init(&c->floor1[w][i], &setup->floor1[w][i]);
  return c;
}

void vbo_destroy(vbo_ctx *c){
  int w, i;
  if(!c) return;
  for(w = 0; w < 2; w++){ free(c->x[w].trig); free(c->x[w].bitrev); free(c->x[w].win); free(c->x[w].wa); }
  for(i = 0; i < c->setup.n_psy && i < 4; i++) psy_free(&c->psy[i]);
  free(c);
}

int vbo_table(vbo_ctx *c, int W, int which, void *dst, int cap){
  vbo_xform *x = &c->x[W];
  int N = x->N;
  switch(which){
  case 0: if(cap < N+N/4) return -1; memcpy(dst, x->trig, sizeof(float)*(N+N/4)); return N+N/4;
  case 1: if(cap < N/4) return -1; memcpy(dst, x->bitrev, sizeof(int)*(N/4)); return N/4;
  case 2: if(cap < N/2) return -1; memcpy(dst, x->win, sizeof(float)*(N/2)); return N/2;
  case 3: if(cap < N) return -1; memcpy(dst, x->wa, sizeof(float)*N); return N;
  }
  return -1;
}

/* ======================================================================= */
/* MDCT core shared by forward and backward                                */

#define C1 .92387953251128675613F  /* cos(pi/8)  (cPI1_8, lib/mdct.h:47) */
#define C2 .70710678118654752441F  /* cos(2pi/8) (cPI2_8) */
#define C3 .38268343236508977175F  /* cos(3pi/8) (cPI3_8) */

/* lib/mdct.c:93-115 */
static void fly8(float *x){
  float s62 = x[6] + x[2], d62 = x[6] - x[2];
  float s40 = x[4] + x[0], d40 = x[4] - x[0];
  float d51 = x[5] - x[1], d73 = x[7] - x[3];
  float s51 = x[5] + x[1], s73 = x[7] + x[3];
  x[6] = s62 + s40;  x[4] = s62 - s40;
  x[0] = d62 + d51;  x[2] = d62 - d51;
  x[3] = d73 + d40;  x[1] = d73 - d40;
  x[7] = s73 + s51;  x[5] = s73 - s51;
}

/* lib/mdct.c:117-150 */
static void fly16(float *x){
  float a, b;
  a = x[1] - x[9];  b = x[0] - x[8];
  x[8] += x[0];     x[9] += x[1];
  x[0] = (a + b) * C2;  x[1] = (a - b) * C2;

  a = x[3] - x[11]; b = x[10] - x[2];
  x[10] += x[2];    x[11] += x[3];
  x[2] = a;         x[3] = b;

  a = x[12] - x[4]; b = x[13] - x[5];
  x[12] += x[4];    x[13] += x[5];
  x[4] = (a - b) * C2;  x[5] = (a + b) * C2;

  a = x[14] - x[6]; b = x[15] - x[7];
  x[14] += x[6];    x[15] += x[7];
  x[6] = a;         x[7] = b;

  fly8(x); fly8(x + 8);
}

/* lib/mdct.c:152-214.  Item q = 0..7 pairs (30-2q, 31-2q) with (14-2q, 15-2q). */
static void fly32(float *x){
  int q;
  for(q = 0; q < 8; q++){
    int a = 30 - 2*q, b = 14 - 2*q;
    float hr = x[a], hi = x[a+1], lr = x[b], li = x[b+1];
    float r0, r1, o0, o1;
    x[a] = hr + lr; x[a+1] = hi + li;
    switch(q){
    case 0: r0 = hr - lr; r1 = hi - li; o0 = r0;                 o1 = r1;                 break;
    case 1: r0 = hr - lr; r1 = hi - li; o0 = r0*C1 - r1*C3;      o1 = r0*C3 + r1*C1;      break;
    case 2: r0 = hr - lr; r1 = hi - li; o0 = (r0 - r1)*C2;       o1 = (r0 + r1)*C2;       break;
    case 3: r0 = hr - lr; r1 = hi - li; o0 = r0*C3 - r1*C1;      o1 = r1*C3 + r0*C1;      break;
    case 4: r0 = hr - lr; r1 = li - hi; o0 = r1;                 o1 = r0;                 break;
    case 5: r0 = lr - hr; r1 = li - hi; o0 = r1*C1 + r0*C3;      o1 = r1*C3 - r0*C1;      break;
    case 6: r0 = lr - hr; r1 = li - hi; o0 = (r1 + r0)*C2;       o1 = (r1 - r0)*C2;       break;
    default:r0 = lr - hr; r1 = li - hi; o0 = r1*C3 + r0*C1;      o1 = r1*C1 - r0*C3;      break;
    }
    x[b] = o0; x[b+1] = o1;
  }
  fly16(x); fly16(x + 16);
}

/* mdct_butterflies, lib/mdct.c:316-336 with butterfly_first/_generic
 * (lib/mdct.c:216-314) flattened: stage s works on sub-blocks of P = points>>s
 * elements with twiddle stride 4<<s; inside a sub-block item q = 0..P/4-1
 * rotates the pair at (P-2-2q) against the pair at (P/2-2-2q).              */
static void butterflies(const vbo_xform *X, float *x, int points){
  const float *T = X->trig;
  int nst = X->log2n - 6;         /* radix-2 stages before the 32-point flies */
  int s, blk, q, j;
  for(s = 0; s < nst; s++){
    int P = points >> s, stride = 4 << s;
    for(blk = 0; blk < (1 << s); blk++){
      float *xb = x + P * blk;
      for(q = 0; q < P/4; q++){
        int a = P - 2 - 2*q, b = (P >> 1) - 2 - 2*q;
        float t0 = T[q*stride], t1 = T[q*stride + 1];
        float r0 = xb[a] - xb[b], r1 = xb[a+1] - xb[b+1];
        xb[a] += xb[b]; xb[a+1] += xb[b+1];
        xb[b]   = r1*t1 + r0*t0;
        xb[b+1] = r1*t0 - r0*t1;
      }
    }
  }
  for(j = 0; j < points; j += 32) fly32(x + j);
}

/* mdct_bitreverse, lib/mdct.c:346-394: item m = 0..N/8-1 reads two complex
 * values of the upper half through bitrev[] and writes w[2m],w[2m+1] and
 * w[N/2-2m-2], w[N/2-2m-1] of the lower half.                               */
static void bitreverse(const vbo_xform *X, float *w){
  int N = X->N, n2 = N >> 1, m;
  const float *T = X->trig + N;
  const float *x = w + n2;
  for(m = 0; m < N/8; m++){
    const float *x0 = x + X->bitrev[2*m];
    const float *x1 = x + X->bitrev[2*m+1];
    float r0 = x0[1] - x1[1];
    float r1 = x0[0] + x1[0];
    float r2 = r1*T[2*m]   + r0*T[2*m+1];
    float r3 = r1*T[2*m+1] - r0*T[2*m];
    float h0 = (x0[1] + x1[1]) * .5f;
    float h1 = (x0[0] - x1[0]) * .5f;
    w[2*m]        = h0 + r2;
    w[2*m+1]      = h1 + r3;
    w[n2-2*m-2]   = h0 - r2;
    w[n2-2*m-1]   = r3 - h1;
  }
}

/* mdct_forward, lib/mdct.c:492-562.  w is N floats of scratch. */
static void mdct_forward1(const vbo_xform *X, const float *in, float *out, float *w){
  int N = X->N, n2 = N >> 1, n4 = N >> 2, n16 = N >> 4, p, i;
  const float *T = X->trig;
  float *w2 = w + n2;
  /* fold + first rotation: item p = 0..N/4-1 produces w2[2p], w2[2p+1]
   * (the three loops at lib/mdct.c:511-542 differ only in which input
   * quarter is read and with which sign) */
  for(p = 0; p < n4; p++){
    float t0 = T[n2 - 2*p - 2], t1 = T[n2 - 2*p - 1];
    float r0, r1;
    if(p < n16){
      const float *x0 = in + n2 + n4 - 4*(p+1);
      const float *x1 = in + n2 + n4 + 1 + 4*p;
      r0 = x0[2] + x1[0];
      r1 = x0[0] + x1[2];
    }else if(p < n4 - n16){
      const float *x0 = in + n2 + n4 - 4*(p+1);
      const float *x1 = in + 1 + 4*(p - n16);
      r0 = x0[2] - x1[0];
      r1 = x0[0] - x1[2];
    }else{
      const float *x0 = in + N - 4*(p - (n4 - n16) + 1);
      const float *x1 = in + 1 + 4*(p - n16);
      r0 = -x0[2] - x1[0];
      r1 = -x0[0] - x1[2];
    }
    w2[2*p]   = r1*t1 + r0*t0;
    w2[2*p+1] = r1*t0 - r0*t1;
  }
  butterflies(X, w2, n2);
  bitreverse(X, w);
  /* final rotation and scale, lib/mdct.c:552-561 */
  T = X->trig + n2;
  for(i = 0; i < n4; i++){
    float a = w[2*i], b = w[2*i+1];
    out[i]        = (a*T[2*i]   + b*T[2*i+1]) * X->scale;
    out[n2-1-i]   = (a*T[2*i+1] - b*T[2*i])   * X->scale;
  }
}

/* mdct_backward, lib/mdct.c:396-490.  in: N/2 coefficients, out: N samples;
 * in and out may not alias here (callers copy).                             */
static void mdct_backward1(const vbo_xform *X, const float *in, float *out){
  int N = X->N, n2 = N >> 1, n4 = N >> 2, n16 = N >> 4, j, k;
  const float *T = X->trig;
  /* input rotation, lib/mdct.c:403-430: item j = 0..N/16-1 of each loop */
  for(j = 0; j < n16; j++){
    const float *iX = in + n2 - 7 - 8*j;
    const float *t = T + n4 + 4*j;
    float *oX = out + n2 + n4 - 4*(j+1);
    oX[0] = -iX[2]*t[3] - iX[0]*t[2];
    oX[1] =  iX[0]*t[3] - iX[2]*t[2];
    oX[2] = -iX[6]*t[1] - iX[4]*t[0];
    oX[3] =  iX[4]*t[1] - iX[6]*t[0];
  }
  for(j = 0; j < n16; j++){
    const float *iX = in + n2 - 8 - 8*j;
    const float *t = T + n4 - 4*(j+1);
    float *oX = out + n2 + n4 + 4*j;
    oX[0] = iX[4]*t[3] + iX[6]*t[2];
    oX[1] = iX[4]*t[2] - iX[6]*t[3];
    oX[2] = iX[0]*t[1] + iX[2]*t[0];
    oX[3] = iX[0]*t[0] - iX[2]*t[1];
  }
  butterflies(X, out + n2, n2);
  bitreverse(X, out);
  /* output rotation + mirroring, lib/mdct.c:437-488, per complex item k:
   *   A[n4-1-k] =  re*T1 - im*T0 ,  B[k] = -(re*T0 + im*T1)
   *   out = [ A | -reverse(A) | reverse(B) | B ]                            */
  T = X->trig + n2;
  for(k = 0; k < n4; k++){
    float re = out[2*k], im = out[2*k+1];
    float a = re*T[2*k+1] - im*T[2*k];
    float b = -(re*T[2*k] + im*T[2*k+1]);
    out[n2 + n4 - 1 - k] = a;       /* parked; lower half is still being read */
    out[n2 + n4 + k]     = b;
  }
  for(k = 0; k < n4; k++){
    float a = out[n2 + n4 - 1 - k];
    out[n4 - 1 - k] = a;
    out[n4 + k]     = -a;
  }
  for(k = 0; k < n4; k++)
    out[n2 + n4 - 1 - k] = out[n2 + n4 + k];
}

void vbo_mdct_forward(vbo_ctx *c, int W, int nvec, const float *in, float *out){
  const vbo_xform *X = &c->x[W];
  int N = X->N, v;
  float *w = (float*)malloc(sizeof(float)*N);
  for(v = 0; v < nvec; v++) mdct_forward1(X, in + (size_t)v*N, out + (size_t)v*(N/2), w);
  free(w);
}

void vbo_mdct_backward(vbo_ctx *c, int W, int nvec, const float *in, float *out){
  const vbo_xform *X = &c->x[W];
  int N = X->N, v;
  for(v = 0; v < nvec; v++) mdct_backward1(X, in + (size_t)v*(N/2), out + (size_t)v*N);
}

/* ======================================================================= */
/* window, lib/window.c:2102-2135                                          */
static void apply_window1(const vbo_ctx *c, float *d, int lW, int W, int nW){
  long n, ln, rn, leftbegin, leftend, rightbegin, rightend, i;
  const float *wl, *wr;
  if(!W){ lW = 0; nW = 0; }
  n = c->x[W].N; ln = c->x[lW].N; rn = c->x[nW].N;
  wl = c->x[lW].win; wr = c->x[nW].win;
  leftbegin = n/4 - ln/4;        leftend  = leftbegin + ln/2;
  rightbegin = n/2 + n/4 - rn/4; rightend = rightbegin + rn/2;
  for(i = 0; i < n; i++){
    if(i < leftbegin || i >= rightend) d[i] = 0.f;
    else if(i < leftend)               d[i] *= wl[i - leftbegin];
    else if(i >= rightbegin)           d[i] *= wr[rn/2 - 1 - (i - rightbegin)];
  }
}

void vbo_apply_window(vbo_ctx *c, int W, int nvec, const int32_t *lW, const int32_t *nW, float *data){
  int N = c->x[W].N, v;
  for(v = 0; v < nvec; v++) apply_window1(c, data + (size_t)v*N, lW ? lW[v] : 0, W, nW ? nW[v] : 0);
}

/* ======================================================================= */
/* real FFT, forward: drftf1 + dradf4 + dradf2, lib/smallft.c:572-631,
 * 168-268, 113-166.  One pass per factor, last factor first, ping-pong
 * between two buffers.  Pass (ip, l1, ido): input viewed as cc[q][k][i]
 * (q<ip, k<l1, i<ido), output as ch[k][q][i].  Items: (k) for i=0, (k,i) for
 * the twiddled interior pairs, (k) for the i=ido-1 column when ido is even. */
static void fft_pass4(int ido, int l1, const float *cc, float *ch,
                      const float *w1, const float *w2, const float *w3){
  const float hsqt2 = .70710678118654752f;
  int t0 = l1 * ido, k, i;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0, *c2 = c1 + t0, *c3 = c2 + t0;
    float *o = ch + 4*k*ido;
    float tr1 = c1[0] + c3[0];
    float tr2 = c0[0] + c2[0];
    o[0]           = tr1 + tr2;
    o[4*ido - 1]   = tr2 - tr1;
    o[2*ido - 1]   = c0[0] - c2[0];
    o[2*ido]       = c3[0] - c1[0];
  }
  if(ido < 2) return;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0, *c2 = c1 + t0, *c3 = c2 + t0;
    float *o = ch + 4*k*ido;
    for(i = 2; i < ido; i += 2){
      float cr2 = w1[i-2]*c1[i-1] + w1[i-1]*c1[i];
      float ci2 = w1[i-2]*c1[i]   - w1[i-1]*c1[i-1];
      float cr3 = w2[i-2]*c2[i-1] + w2[i-1]*c2[i];
      float ci3 = w2[i-2]*c2[i]   - w2[i-1]*c2[i-1];
      float cr4 = w3[i-2]*c3[i-1] + w3[i-1]*c3[i];
      float ci4 = w3[i-2]*c3[i]   - w3[i-1]*c3[i-1];
      float tr1 = cr2 + cr4, tr4 = cr4 - cr2;
      float ti1 = ci2 + ci4, ti4 = ci2 - ci4;
      float ti2 = c0[i] + ci3,   ti3 = c0[i] - ci3;
      float tr2 = c0[i-1] + cr3, tr3 = c0[i-1] - cr3;
      int ic = 2*ido - i;
      o[i-1]          = tr1 + tr2;   o[i]          = ti1 + ti2;
      o[ic-1]         = tr3 - ti4;   o[ic]         = tr4 - ti3;
      o[2*ido+i-1]    = ti4 + tr3;   o[2*ido+i]    = tr4 + ti3;
      o[2*ido+ic-1]   = tr2 - tr1;   o[2*ido+ic]   = ti1 - ti2;
    }
  }
  if(ido & 1) return;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0, *c2 = c1 + t0, *c3 = c2 + t0;
    float *o = ch + 4*k*ido;
    float ti1 = -hsqt2 * (c1[ido-1] + c3[ido-1]);
    float tr1 =  hsqt2 * (c1[ido-1] - c3[ido-1]);
    o[ido-1]   = tr1 + c0[ido-1];
    o[3*ido-1] = c0[ido-1] - tr1;
    o[ido]     = ti1 - c2[ido-1];
    o[3*ido]   = ti1 + c2[ido-1];
  }
}

static void fft_pass2(int ido, int l1, const float *cc, float *ch, const float *w1){
  int t0 = l1 * ido, k, i;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0;
    float *o = ch + 2*k*ido;
    o[0]         = c0[0] + c1[0];
    o[2*ido - 1] = c0[0] - c1[0];
  }
  if(ido < 2) return;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0;
    float *o = ch + 2*k*ido;
    for(i = 2; i < ido; i += 2){
      float tr2 = w1[i-2]*c1[i-1] + w1[i-1]*c1[i];
      float ti2 = w1[i-2]*c1[i]   - w1[i-1]*c1[i-1];
      int ic = 2*ido - i;
      o[i]    = c0[i] + ti2;     o[ic]   = ti2 - c0[i];
      o[i-1]  = c0[i-1] + tr2;   o[ic-1] = c0[i-1] - tr2;
    }
  }
  if(ido & 1) return;
  for(k = 0; k < l1; k++){
    const float *c0 = cc + k*ido, *c1 = c0 + t0;
    float *o = ch + 2*k*ido;
    o[ido]   = -c1[ido-1];
    o[ido-1] = c0[ido-1];
  }
}

static void drft_forward1(const vbo_xform *X, float *data, float *scratch){
  int N = X->N, nf = X->nf, k1;
  int l2 = N, iw = N;
  float *src = data, *dst = scratch;
  for(k1 = 0; k1 < nf; k1++){
    int ip = X->fac[nf - 1 - k1];
    int l1 = l2 / ip, ido = N / l2;
    iw -= (ip - 1) * ido;
    if(ip == 4) fft_pass4(ido, l1, src, dst, X->wa + iw - 1, X->wa + iw + ido - 1, X->wa + iw + 2*ido - 1);
    else        fft_pass2(ido, l1, src, dst, X->wa + iw - 1);
    { float *t = src; src = dst; dst = t; }
    l2 = l1;
  }
  if(src != data) memcpy(data, src, sizeof(float)*N);
}

void vbo_drft_forward(vbo_ctx *c, int W, int nvec, float *data){
  const vbo_xform *X = &c->x[W];
  int N = X->N, v;
  float *s = (float*)malloc(sizeof(float)*N);
  for(v = 0; v < nvec; v++) drft_forward1(X, data + (size_t)v*N, s);
  free(s);
}

/* ======================================================================= */
/* dB conversion, lib/scales.h:43-51 */
static float todB(float x){
  union { uint32_t i; float f; } u;
  u.f = x;
  u.i &= 0x7fffffffu;
  return (float)(u.i * 7.17711438e-7f - 764.6161886f);
}

/* log spectra of one channel, lib/mapping0.c:255-346 (FFT) and :384-385 (MDCT).
 * fft: N floats in FFTPACK order; logfft: n floats; returns local_ampmax.  */
static float log_fft1(int N, const float *fft, float *logfft){
  int n = N/2, k;
  float scale = 4.f / N;
  float scale_dB = todB(scale) + .345;         /* double add, rounded once */
  float amax;
  logfft[0] = scale_dB + todB(fft[0]) + .345;
  amax = logfft[0];
  for(k = 1; k < n; k++){
    float re = fft[2*k-1], im = fft[2*k];
    float t = re*re + im*im;
    t = logfft[k] = scale_dB + .5f*todB(t) + .345;
    if(t > amax) amax = t;
  }
  if(amax > 0.f) amax = 0.f;
  return amax;
}

static void log_mdct1(int n, const float *mdct, float *logmdct){
  int j;
  for(j = 0; j < n; j++) logmdct[j] = todB(mdct[j]) + .345;
}

/* ======================================================================= */
/* noise mask: bark_noise_hybridmp lib/psy.c:547-704, _vp_noisemask :706-752 */

/* the five running sums (lib/psy.c:565-602): strictly sequential fp32 */
static void noise_prefix(int n, const float *f, float offset,
                         float *N, float *X, float *XX, float *Y, float *XY){
  float tN = 0.f, tX = 0.f, tXX = 0.f, tY = 0.f, tXY = 0.f;
  int i;
  for(i = 0; i < n; i++){
    float x = (float)i;
    float y = f[i] + offset, w;
    if(y < 1.f) y = 1.f;
    w = y * y;
    if(i == 0){                    /* first bin: half weight, and the       */
      w = w * .5f;                 /* quirk tX += w (not w*x), :571-584     */
      tN += w; tX += w; tY += w * y;
    }else{
      tN += w; tX += w * x; tXX += w * x * x; tY += w * y; tXY += w * x * y;
    }
    N[i] = tN; X[i] = tX; XX[i] = tXX; Y[i] = tY; XY[i] = tXY;
  }
}

/* regression value for a window [lo,hi] (lib/psy.c:604-650); returns 0 and
 * leaves A,B,D untouched when the window falls off the table (those bins
 * reuse the last A,B,D: :652-658)                                          */
typedef struct { float A, B, D; } abd_t;

static int window_abd(int n, int lo, int hi, const float *N, const float *X, const float *XX,
                      const float *Y, const float *XY, abd_t *o){
  float tN, tX, tXX, tY, tXY;
  if(hi >= n) return 0;
  if(lo < 0){
    if(-lo >= n) return 0;
    tN = N[hi] + N[-lo];  tX = X[hi] - X[-lo];  tXX = XX[hi] + XX[-lo];
    tY = Y[hi] + Y[-lo];  tXY = XY[hi] - XY[-lo];
  }else{
    if(lo >= n) return 0;
    tN = N[hi] - N[lo];   tX = X[hi] - X[lo];   tXX = XX[hi] - XX[lo];
    tY = Y[hi] - Y[lo];   tXY = XY[hi] - XY[lo];
  }
  o->A = tY * tXX - tX * tXY;
  o->B = tN * tXY - tX * tY;
  o->D = tN * tXX - tX * tX;
  return 1;
}

/* The reference runs three consecutive loops (mirrored windows, plain windows,
 * extrapolation) each of which stops at the first bin that does not qualify.
 * `first_plain` / `first_extra` are those stopping points; they depend only on
 * the window table, so the per-bin work is independent given them.          */
static void noise_pass(const vbo_psy *p, const float *f, float *noise, float offset, int fixed,
                       float *N, float *X, float *XX, float *Y, float *XY){
  int n = p->s.n, i;
  int first_plain, first_extra;
  abd_t last = {0.f, 0.f, 1.f}, cur;

  noise_prefix(n, f, offset, N, X, XX, Y, XY);

  /* bark windows */
  for(i = 0; i < n; i++){
    int lo = p->bark[i] >> 16, hi = p->bark[i] & 0xffff;
    if(lo >= 0 || -lo >= n || hi >= n) break;
  }
  first_plain = i;
  for(; i < n; i++){
    int lo = p->bark[i] >> 16, hi = p->bark[i] & 0xffff;
    if(lo < 0 || lo >= n || hi >= n) break;
  }
  first_extra = i;
  for(i = 0; i < n; i++){
    float R;
    if(i < first_extra){
      int lo = p->bark[i] >> 16, hi = p->bark[i] & 0xffff;
      window_abd(n, lo, hi, N, X, XX, Y, XY, &cur);
      if(i == first_extra - 1) last = cur;
    }else cur = last;
    R = (cur.A + (float)i * cur.B) / cur.D;
    if(R < 0.f) R = 0.f;
    noise[i] = R - offset;
  }
  (void)first_plain;
  if(fixed <= 0) return;

  /* fixed-width windows (lib/psy.c:660-703); note A,B,D carry over from the
   * bark loops when no fixed window qualifies at all                        */
  for(i = 0; i < n; i++){
    int hi = i + fixed/2, lo = hi - fixed;
    if(hi >= n || lo >= 0) break;
  }
  for(; i < n; i++){
    int hi = i + fixed/2, lo = hi - fixed;
    if(hi >= n || lo < 0) break;
  }
  first_extra = i;
  for(i = 0; i < n; i++){
    float R;
    if(i < first_extra){
      int hi = i + fixed/2, lo = hi - fixed;
      window_abd(n, lo, hi, N, X, XX, Y, XY, &cur);
      if(i == first_extra - 1) last = cur;
    }else cur = last;
    R = (cur.A + (float)i * cur.B) / cur.D;
    if(R - offset < noise[i]) noise[i] = R - offset;
  }
}

static void noisemask1(const vbo_psy *p, const float *logmdct, float *noise, float *scratch){
  int n = p->s.n, i;
  float *work = scratch, *N = work + n, *X = N + n, *XX = X + n, *Y = XX + n, *XY = Y + n;
  noise_pass(p, logmdct, noise, 140.f, -1, N, X, XX, Y, XY);
  for(i = 0; i < n; i++) work[i] = logmdct[i] - noise[i];
  noise_pass(p, work, noise, 0.f, p->s.noisewindowfixed, N, X, XX, Y, XY);
  for(i = 0; i < n; i++){
    float base = logmdct[i] - work[i];
    int dB = (int)(noise[i] + .5);             /* double add, truncation */
    if(dB >= VB200_COMPAND_LEVELS) dB = VB200_COMPAND_LEVELS - 1;
    if(dB < 0) dB = 0;
    noise[i] = base + p->s.noisecompand[dB];
  }
}

void vbo_noisemask(vbo_ctx *c, int look, int nvec, const float *logmdct, float *noise){
  const vbo_psy *p = &c->psy[look];
  int n = p->s.n, v;
  float *scratch = (float*)malloc(sizeof(float)*6*n);
  for(v = 0; v < nvec; v++) noisemask1(p, logmdct + (size_t)v*n, noise + (size_t)v*n, scratch);
  free(scratch);
}

/* ======================================================================= */
/* tone mask: _vp_tonemask lib/psy.c:754-777 and helpers                   */

/* seed_chase, lib/psy.c:454-508 — emulated literally (its pop rule is not a
 * plain sliding maximum)                                     */
static void seed_chase1(float *seeds, int linesper, int n, int *posstack, float *ampstack){
  int stack = 0, pos = 0, i;
  for(i = 0; i < n; i++){
    if(stack >= 2){
      while(!(seeds[i] < ampstack[stack-1]) &&
            i < posstack[stack-1] + linesper &&
            stack > 1 && ampstack[stack-1] <= ampstack[stack-2] &&
            i < posstack[stack-2] + linesper)
        stack--;
    }
    posstack[stack] = i;
    ampstack[stack++] = seeds[i];
  }
  for(i = 0; i < stack; i++){
    int endpos;
    if(i < stack-1 && ampstack[i+1] > ampstack[i]) endpos = posstack[i+1];
    else endpos = posstack[i] + linesper + 1;
    if(endpos > n) endpos = n;
    for(; pos < endpos; pos++) seeds[pos] = ampstack[i];
  }
}

static void tonemask1(const vbo_psy *p, const float *logfft, float *tone,
                      float gmax, float lmax, float *seed, int *posstack, float *ampstack){
  const vb200_psy_setup *s = &p->s;
  int n = s->n, total = s->total_octave_lines, linesper = s->eighth_octave_lines;
  int i, r, g;
  float att = lmax + s->ath_adjatt;
  float dBoffset = s->max_curve_dB - gmax;
  if(att < s->ath_maxatt) att = s->ath_maxatt;
  for(i = 0; i < total; i++) seed[i] = NEGINF;
  for(i = 0; i < n; i++) tone[i] = p->ath[i] + att;

  /* seed_loop + seed_curve, lib/psy.c:417-452, 390-415: one item per run */
  for(r = 0; r < p->nruns; r++){
    int lo = p->run_lo[r], hi = p->run_hi[r];
    float mx = logfft[lo];
    for(i = lo + 1; i <= hi; i++) if(logfft[i] > mx) mx = logfft[i];
    if(mx + 6.f > tone[hi]){
      int oc = p->octave[hi] >> s->shiftoc;
      const float *posts, *curve;
      int choice, post0, post1, seedptr;
      if(oc >= VB200_P_BANDS) oc = VB200_P_BANDS - 1;
      if(oc < 0) oc = 0;
      choice = (int)(((mx + dBoffset) - 30.) * .1f);   /* P_LEVEL_0 is a double */
      if(choice < 0) choice = 0;
      if(choice > VB200_P_LEVELS - 1) choice = VB200_P_LEVELS - 1;
      posts = p->tonecurves + ((size_t)oc*VB200_P_LEVELS + choice)*(VB200_EHMER_MAX+2);
      curve = posts + 2;
      post0 = (int)posts[0]; post1 = (int)posts[1];
      seedptr = (p->octave[hi] - s->firstoc) + (post0 - 16)*linesper - (linesper >> 1);
      for(i = post0; i < post1; i++){
        if(seedptr > 0){
          float lin = mx + curve[i];
          if(seed[seedptr] < lin) seed[seedptr] = lin;
        }
        seedptr += linesper;
        if(seedptr >= total) break;
      }
    }
  }

  /* max_seeds, lib/psy.c:512-545 */
  seed_chase1(seed, linesper, total, posstack, ampstack);
  for(g = 0; g < p->ngrp; g++){
    int pos = p->grp_pos0[g];
    float minV = seed[pos];
    if(minV > s->tone_abs_limit) minV = s->tone_abs_limit;
    while(pos < p->grp_pos1[g]){
      pos++;
      if((seed[pos] > NEGINF && seed[pos] < minV) || minV == NEGINF) minV = seed[pos];
    }
    for(i = p->grp_lin0[g]; i < p->grp_lin1[g]; i++)
      if(tone[i] < minV) tone[i] = minV;
  }
  {
    float minV = seed[total-1];
    for(i = p->tail_lin0; i < n; i++) if(tone[i] < minV) tone[i] = minV;
  }
}

void vbo_tonemask(vbo_ctx *c, int look, int nvec, const float *logfft,
                  const float *gmax, const float *lmax, float *tone){
  const vbo_psy *p = &c->psy[look];
  int n = p->s.n, total = p->s.total_octave_lines, v;
  float *seed = (float*)malloc(sizeof(float)*total);
  float *amp = (float*)malloc(sizeof(float)*total);
  int *pos = (int*)malloc(sizeof(int)*total);
  for(v = 0; v < nvec; v++)
    tonemask1(p, logfft + (size_t)v*n, tone + (size_t)v*n, gmax[v], lmax[v], seed, pos, amp);
  free(seed); free(amp); free(pos);
}

/* ======================================================================= */
/* _vp_offset_and_mix, lib/psy.c:779-835 */
static void offset_and_mix1(const vbo_psy *p, int sel, const float *noise, const float *tone,
                            float *mdct, const float *logmdct, float *logmask){
  const vb200_psy_setup *s = &p->s;
  int n = s->n, i;
  float toneatt = s->tone_masteratt[sel];
  float cx = s->m_val;
  const float *noff = p->noiseoffset + (size_t)sel*n;
  for(i = 0; i < n; i++){
    float val = noise[i] + noff[i];
    float t;
    if(val > s->noisemaxsupp) val = s->noisemaxsupp;
    t = tone[i] + toneatt;
    logmask[i] = val < t ? t : val;                 /* max() of lib/os.h */
    if(sel == 1){
      float coeffi = -17.2f, de;
      val = val - logmdct[i];
      if(val > coeffi){
        de = 1.0 - ((val - coeffi) * 0.005 * cx);   /* evaluated in double */
        if(de < 0) de = 0.0001f;
      }else
        de = 1.0 - ((val - coeffi) * 0.0003 * cx);
      mdct[i] *= de;
    }
  }
}

void vbo_offset_and_mix(vbo_ctx *c, int look, int nvec, int sel, const float *noise, const float *tone,
                        float *mdct, const float *logmdct, float *logmask){
  const vbo_psy *p = &c->psy[look];
  int n = p->s.n, v;
  for(v = 0; v < nvec; v++){
    size_t o = (size_t)v*n;
    offset_and_mix1(p, sel, noise+o, tone+o, mdct+o, logmdct+o, logmask+o);
  }
}

/* ======================================================================= */
/* Phase A: the two per-channel loops of mapping0_forward, lib/mapping0.c:254-470 */

/* _vp_ampmax_decay, lib/psy.c:837-848 */
float vbo_ampmax_decay(vbo_ctx *c, float amp, int W){
  int n = c->setup.blocksizes[W] / 2;
  float secs = (float)n / c->setup.rate;
  amp += secs * c->setup.ampmax_att_per_sec;
  if(amp < -9999) amp = -9999;
  return amp;
}

typedef struct { float *win, *fftw, *scr, *noise, *tone, *seed, *amp, *nscr; int *pos; float *lmax; } pa_scratch;

static void pa_alloc(vbo_ctx *c, int W, pa_scratch *s){
  int N = c->x[W].N, ch = c->setup.channels, n = N/2;
  int total = 0, i;
  for(i = 0; i < 4; i++) if(c->psy[i].s.total_octave_lines > total) total = c->psy[i].s.total_octave_lines;
  s->win = (float*)malloc(sizeof(float)*N*ch);    /* per channel: windowed pcm -> fft -> logfft|logmdct */
  s->fftw = (float*)malloc(sizeof(float)*N);
  s->scr = (float*)malloc(sizeof(float)*N);
  s->noise = (float*)malloc(sizeof(float)*n);
  s->tone = (float*)malloc(sizeof(float)*n);
  s->seed = (float*)malloc(sizeof(float)*(total+1));
  s->amp = (float*)malloc(sizeof(float)*(total+1));
  s->pos = (int*)malloc(sizeof(int)*(total+1));
  s->nscr = (float*)malloc(sizeof(float)*6*n);
  s->lmax = (float*)malloc(sizeof(float)*ch);
}
static void pa_free(pa_scratch *s){
  free(s->win); free(s->fftw); free(s->scr); free(s->noise); free(s->tone);
  free(s->seed); free(s->amp); free(s->pos); free(s->nscr); free(s->lmax);
}

/* first loop (lib/mapping0.c:254-360) for one block: window, MDCT, FFT, logfft,
 * local maxima; returns max over channels of local_ampmax                   */
static float pa_transform(vbo_ctx *c, int W, int blk, const vb200_phaseA_io *io, pa_scratch *s){
  const vbo_xform *X = &c->x[W];
  int N = X->N, n = N/2, ch = c->setup.channels, i;
  float m = NEGINF * 10.f;
  for(i = 0; i < ch; i++){
    size_t row = (size_t)blk*ch + i;
    float *p = s->win + (size_t)i*N;
    float *gm = io->mdct + row*n;
    memcpy(p, io->pcm + row*N, sizeof(float)*N);
    apply_window1(c, p, io->desc[blk].lW, W, io->desc[blk].nW);
    mdct_forward1(X, p, gm, s->scr);
    if(io->tap_mdct_raw) memcpy(io->tap_mdct_raw + row*n, gm, sizeof(float)*n);
    drft_forward1(X, p, s->scr);
    memcpy(s->fftw, p, sizeof(float)*N);
    s->lmax[i] = log_fft1(N, s->fftw, p);              /* logfft into p[0..n) */
    if(io->tap_logfft) memcpy(io->tap_logfft + row*n, p, sizeof(float)*n);
    if(s->lmax[i] > m) m = s->lmax[i];
  }
  return m;
}

/* second loop (lib/mapping0.c:366-470) for one block given global_ampmax */
static void pa_psy(vbo_ctx *c, int W, int blk, float gmax, const vb200_phaseA_io *io, pa_scratch *s){
  int N = c->x[W].N, n = N/2, ch = c->setup.channels, i;
  const vbo_psy *p = &c->psy[io->desc[blk].blocktype + (W ? 2 : 0)];
  for(i = 0; i < ch; i++){
    size_t row = (size_t)blk*ch + i;
    float *logfft = s->win + (size_t)i*N;
    float *gm = io->mdct + row*n;
    float *logmdct = io->logmdct + row*n;
    log_mdct1(n, gm, logmdct);
    noisemask1(p, logmdct, s->noise, s->nscr);
    tonemask1(p, logfft, s->tone, gmax, s->lmax[i], s->seed, s->pos, s->amp);
    if(io->tap_noise) memcpy(io->tap_noise + row*n, s->noise, sizeof(float)*n);
    if(io->tap_tone)  memcpy(io->tap_tone  + row*n, s->tone,  sizeof(float)*n);
    offset_and_mix1(p, 1, s->noise, s->tone, gm, logmdct, io->logmask + row*n);
  }
}

void vbo_phaseA(vbo_ctx *c, int W, int nblocks, const vb200_phaseA_io *io){
  pa_scratch s; int blk;
  pa_alloc(c, W, &s);
  for(blk = 0; blk < nblocks; blk++){
    float g = io->desc[blk].ampmax;
    float m = pa_transform(c, W, blk, io, &s);
    if(m > g) g = m;                                   /* lib/mapping0.c:346 */
    pa_psy(c, W, blk, g, io, &s);
    io->ampmax_out[blk] = g;                           /* lib/mapping0.c:576 */
  }
  pa_free(&s);
}

/* stream mode: the ampmax chain of vorbis_analysis_blockout, lib/block.c:626-628:
 *   in[k] = decay(max(in[k-1], out[k-1])),  out[k] = max(in[k], locals[k])
 * with in[-1] = out[-1] = ampmax0 (g->ampmax and vbi->ampmax both start at
 * -9999, lib/psy.c:42, lib/block.c:92).                                     */
void vbo_phaseA_streams(vbo_ctx *c, int W, int nstreams, int bps, const vb200_phaseA_io *io,
                        const float *ampmax0){
  pa_scratch s; int st, k;
  pa_alloc(c, W, &s);
  for(st = 0; st < nstreams; st++){
    float g = ampmax0 ? ampmax0[st] : -9999.f;   /* g->ampmax */
    float prev_out = g;
    for(k = 0; k < bps; k++){
      int blk = st*bps + k;
      float in, m, out;
      if(prev_out > g) g = prev_out;
      g = vbo_ampmax_decay(c, g, W);
      in = g;
      m = pa_transform(c, W, blk, io, &s);
      out = in; if(m > out) out = m;
      pa_psy(c, W, blk, out, io, &s);
      io->ampmax_out[blk] = out;
      prev_out = out;
    }
  }
  pa_free(&s);
}

/* ======================================================================= */
/* Phase B and decode are in vb_oracle_b.c (included to keep one library)  */
#include "vb_oracle_b.inc"
#include "vb_oracle_floor.inc"
#include "vb_oracle_env.inc"
#include "vb_oracle_res.inc"
#include "vb_oracle_plan.inc"
