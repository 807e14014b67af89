/* vb_oracle.h — CPU oracle: a plain-C restatement of the libvorbis per-block DSP
 * path, written from scratch in "independent work item" form (every stage is a
 * loop over items that touch disjoint data, which is also how the CUDA kernels
 * are organised).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; it is the checker, never
 * the product.  The product (vorbis_b200/csrc) shares no code with it.
 *
 * Pinning: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md §8c), so the oracle is pinned against outputs of the reference
 * itself: oracle/_ref/libvorbis_ref.so (the unmodified reference sources built
 * by oracle/Makefile) in this container, and the fixtures under tests/golden/
 * that tests/golden/make_golden.py generated from it.  The restatement is
 * bit-exact against both (tests/test_oracle_vs_ref.py, tests/test_golden.py).
 *
 * Build flags matter: -O2 -ffp-contract=off, no fast-math (oracle/Makefile).
 */
#ifndef VB_ORACLE_H
#define VB_ORACLE_H
#include "vorbis_b200.h"

typedef struct vbo_ctx vbo_ctx;

vbo_ctx *vbo_create(const vb200_setup *setup);
void     vbo_destroy(vbo_ctx *c);
int      vbo_table(vbo_ctx *c, int W, int which, void *dst, int cap);

void vbo_mdct_forward (vbo_ctx *c, int W, int nvec, const float *in, float *out);
void vbo_mdct_backward(vbo_ctx *c, int W, int nvec, const float *in, float *out);
void vbo_apply_window (vbo_ctx *c, int W, int nvec, const int32_t *lW, const int32_t *nW, float *data);
void vbo_drft_forward (vbo_ctx *c, int W, int nvec, float *data);

void vbo_noisemask(vbo_ctx *c, int look, int nvec, const float *logmdct, float *noise);
void vbo_tonemask (vbo_ctx *c, int look, int nvec, const float *logfft,
                   const float *gmax, const float *lmax, float *tone);
void vbo_offset_and_mix(vbo_ctx *c, int look, int nvec, int sel, const float *noise, const float *tone,
                        float *mdct, const float *logmdct, float *logmask);

void vbo_phaseA(vbo_ctx *c, int W, int nblocks, const vb200_phaseA_io *io);
void vbo_phaseA_streams(vbo_ctx *c, int W, int nstreams, int bps, const vb200_phaseA_io *io,
                        const float *ampmax0);
float vbo_ampmax_decay(vbo_ctx *c, float amp, int W);

void vbo_couple_quantize_normalize(vbo_ctx *c, int W, int blocktype, int blobno, int nblocks,
                                   const float *mdct, int32_t *iwork, int32_t *nonzero);

void vbo_synthesis(vbo_ctx *c, int nstreams, int nblk, const int32_t *Wseq,
                   const int64_t *coef_off, const float *coef,
                   const int64_t *pcm_off, float *pcm, int64_t pcm_stride);
void vbo_decouple(vbo_ctx *c, int W, int nblocks, float *res);
void vbo_floor1_fit(vbo_ctx *c, int W, int floor_sel, int nrows, const float *logmdct, const float *logmask,
                    int32_t *posts, int32_t *fit_nonzero);
void vbo_floor1_render(vbo_ctx *c, int W, int floor_sel, int nrows, int32_t *posts, const int32_t *fit_nonzero,
                       int32_t *ilogmask, int32_t *nonzero);
void vbo_floor1_inverse2(vbo_ctx *c, int W, int floor_sel, int nrows, const int32_t *posts,
                         const int32_t *present, float *data);
void vbo_decode_dsp(vbo_ctx *c, int nstreams, int nblk, const int32_t *Wseq, const int64_t *coef_off,
                    float *res, const int32_t *posts, const int32_t *present,
                    const int64_t *pcm_off, float *pcm, int64_t pcm_stride);
int  vbo_residue_partvals(vbo_ctx *c, int W);
void vbo_residue_classify(vbo_ctx *c, int W, int nblocks, const int32_t *iwork, const int32_t *nonzero,
                          int32_t *classes, int stride);
void vbo_envelope_search(vbo_ctx *c, int nstreams, const float *pcm, int64_t stride, int first_step,
                         int nsteps, int32_t *state, uint8_t *ret);
void vbo_envelope_apply_marks(const uint8_t *ret, int first_step, int nsteps, int32_t *mark);
void vbo_plan_blocks(vbo_ctx *c, int nstreams, const int32_t *mark, int64_t mark_stride, int nsteps,
                     const int64_t *pcm_len, const int64_t *eof, int max_blocks,
                     vb200_stream_block *plan, int32_t *nblocks);
#endif
