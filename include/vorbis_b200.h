/* vorbis_b200.h — C ABI of the B200-native per-block DSP path of libvorbis.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point names the reference interface it replaces (paths relative
 * to the xiph/vorbis tree, libvorbis 1.3.7).  The reference-side binding a
 * libvorbis maintainer would add (a replacement `mapping0_exportbundle`) is
 * shown in INTEGRATION.md and implemented in vorbis_b200/host/.
 *
 * Conventions
 *   - return 0 on success, a negative OV_* style code on failure (never abort;
 *     lib/mapping0.c:498 returns -1, include/vorbis/codec.h:217-233 OV_*);
 *     vb200_last_error() gives a thread-local message.
 *   - N  = block size in samples (vorbis_block.pcmend), n = N/2 spectral lines.
 *   - W  = block-size flag (0 short, 1 long), as vorbis_block.W.
 *   - batches are homogeneous in W; vectors are laid out [block][channel][...]
 *     contiguous fp32 (the reference's vb->pcm[ch][N] stacked block-major).
 *   - functions ending in _dev take DEVICE pointers and a cudaStream_t passed
 *     as void* (NULL = legacy default stream) and are asynchronous;
 *     the others take HOST pointers, copy in/out and synchronise.
 *   - a context keeps grow-only device scratch: use one context per host thread
 *     (the reference is single threaded per vorbis_dsp_state as well);
 *     the host-pointer entry points serialise on a per-context mutex.  _dev calls
 *     that use that scratch (Phase A, encode_dsp, encode_streams, envelope_search)
 *     may be issued on different CUDA streams: each waits (on the device, via an
 *     event) for the previous such call of the same context, so they never share
 *     the scratch in time.  Growing the scratch re-allocates (cudaFree/cudaMalloc
 *     synchronise the device): the first call at a new maximum size is not
 *     asynchronous.
 */
#ifndef VORBIS_B200_H
#define VORBIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_OK        0
#define VB200_EFAULT  (-129)   /* OV_EFAULT: CUDA/runtime failure            */
#define VB200_EIMPL   (-130)   /* OV_EIMPL : configuration not supported      */
#define VB200_EINVAL  (-131)   /* OV_EINVAL: bad argument                     */

#define VB200_P_BANDS        17   /* lib/psy.h:28  */
#define VB200_P_LEVELS        8   /* lib/psy.h:29  */
#define VB200_P_NOISECURVES   3   /* lib/psy.h:31  */
#define VB200_EHMER_MAX      56   /* lib/masking.h:43 */
#define VB200_COMPAND_LEVELS 40   /* lib/psy.h:33  */
#define VB200_PACKETBLOBS    15   /* lib/codec_internal.h:28 */
#define VB200_MAX_CHANNELS  255
#define VB200_VE_BANDS        7   /* lib/envelope.h:29 */
#define VB200_MAX_COUPLING  256

/* One psychoacoustic lookup == vorbis_look_psy + the vorbis_info_psy scalars
 * the per-block code reads (lib/psy.h:35-65, 96-114).  Built by the
 * reference's _vp_psy_init (lib/psy.c:266); we only consume it.            */
typedef struct vb200_psy_setup {
  int32_t n;                      /* vorbis_look_psy.n  (= blocksize/2)      */
  int32_t blockflag;              /* vorbis_info_psy.blockflag               */
  float   ath_adjatt;
  float   ath_maxatt;
  float   tone_masteratt[VB200_P_NOISECURVES];
  float   tone_abs_limit;
  float   noisemaxsupp;
  int32_t noisewindowfixed;
  float   noisecompand[VB200_COMPAND_LEVELS];
  float   max_curve_dB;
  int32_t normal_p;
  int32_t normal_start;
  int32_t normal_partition;
  double  normal_thresh;
  int32_t firstoc;
  int32_t shiftoc;
  int32_t eighth_octave_lines;
  int32_t total_octave_lines;
  float   m_val;
  const float   *ath;             /* [n]                                      */
  const int32_t *octave;          /* [n]   (reference: long)                  */
  const int32_t *bark;            /* [n]   packed ((lo-1)<<16)+(hi-1)         */
  const float   *tonecurves;      /* [P_BANDS][P_LEVELS][EHMER_MAX+2]         */
  const float   *noiseoffset;     /* [P_NOISECURVES][n]                       */
} vb200_psy_setup;

/* Floor 1 configuration of one block size: vorbis_info_floor1 (lib/backends.h:64-91).  The sorted /
 * forward / reverse indices and the decode neighbours of vorbis_look_floor1
 * (lib/codec_internal.h:138-155, built by floor1_look lib/floor1.c:180-259) are re-derived from
 * postlist by the context.                                                                    */
#define VB200_VIF_POSIT 63
#define VB200_FLOOR1_STRIDE (VB200_VIF_POSIT + 2)   /* ints per row in every posts array of this API */
#define VB200_MAX_SUBMAPS 4                          /* libvorbis' encoder setups use 1 or 2 (5.1: LFE) */
typedef struct vb200_floor1_setup {
  int32_t posts;                          /* vorbis_look_floor1.posts (<= VIF_POSIT+2); 0 = absent */
  int32_t postlist[VB200_VIF_POSIT + 2];
  int32_t mult;                           /* 1..4 */
  int32_t n;                              /* spectral lines this floor spans (postlist[1])  */
  float   maxover, maxunder, maxerr;
  float   twofitweight, twofitatten;
} vb200_floor1_setup;

/* Residue partition classification parameters of one submap: vorbis_info_residue0 (lib/backends.h:103-118)
 * + the residue type (codec_setup_info.residue_type).  Only what res{0,1,2}_class read.              */
typedef struct vb200_residue_setup {
  int32_t type;                           /* 0, 1 or 2; -1 = not provided                          */
  int32_t begin, end;                     /* in samples (type 2: interleaved samples of the bundle) */
  int32_t grouping;                       /* samples per partition                                  */
  int32_t partitions;                     /* number of partition classes                            */
  int32_t classmetric1[64];
  int32_t classmetric2[64];
} vb200_residue_setup;

/* Everything the kernels need from codec_setup_info / private_state
 * (lib/codec_internal.h:59-133) for one (channels, rate, quality) setup.    */
typedef struct vb200_setup {
  int32_t channels;
  int32_t rate;
  int32_t blocksizes[2];          /* codec_setup_info.blocksizes              */
  int32_t n_psy;                  /* 0 (decode/transform only) or 4           */
  vb200_psy_setup psy[4];         /* private_state.psy[blocktype+2*W]         */
  /* vorbis_info_psy_global (lib/psy.h:67-85) */
  float   ampmax_att_per_sec;
  int32_t coupling_pointlimit[2][VB200_PACKETBLOBS];
  int32_t coupling_prepointamp[VB200_PACKETBLOBS];
  int32_t coupling_postpointamp[VB200_PACKETBLOBS];
  int32_t sliding_lowpass[2][VB200_PACKETBLOBS];
  /* vorbis_info_mapping0 coupling (lib/backends.h:127-141), per W           */
  int32_t coupling_steps[2];
  int32_t coupling_mag[2][VB200_MAX_COUPLING];
  int32_t coupling_ang[2][VB200_MAX_COUPLING];
  /* optional half-window tables, vwin[] of lib/window.c:23-2096 for
   * blocksizes[0] and [1] (blocksize/2 floats each); NULL = compute from the
   * closed form of doc/04-codec.tex:320                                     */
  const float *window[2];
  /* floors of mode W (lib/mapping0.c:499-506): channel i uses floor1[W][chmux[W][i]], i.e. the
   * floor of its submap (vorbis_info_mapping0.chmuxlist / floorsubmap, lib/backends.h:127-141).
   * posts == 0: not provided (the floor entry points then return -1)                          */
  int32_t submaps[2];
  uint8_t chmux[2][VB200_MAX_CHANNELS + 1];
  vb200_floor1_setup floor1[2][VB200_MAX_SUBMAPS];
  /* envelope / block-switch detector: vorbis_info_psy_global.preecho_thresh, postecho_thresh,
   * stretch_penalty, preecho_minenergy (lib/psy.h:67-85), read by _ve_amp (lib/envelope.c:88-213) */
  float   preecho_thresh[VB200_VE_BANDS];
  float   postecho_thresh[VB200_VE_BANDS];
  float   stretch_penalty;
  float   preecho_minenergy;
  /* residue of submap sm of mode W: residue_param[ mapping.residuesubmap[sm] ] (lib/mapping0.c:663) */
  vb200_residue_setup residue[2][VB200_MAX_SUBMAPS];
} vb200_setup;

/* Per-block inputs of mapping0_forward that are not PCM (lib/mapping0.c:230-252) */
typedef struct vb200_block_desc {
  int32_t lW;                     /* vorbis_block.lW                          */
  int32_t nW;                     /* vorbis_block.nW                          */
  int32_t blocktype;              /* vorbis_block_internal.blocktype (0/1)    */
  float   ampmax;                 /* vorbis_block_internal.ampmax on entry    */
} vb200_block_desc;

typedef struct vb200_ctx vb200_ctx;

/* ---- context ---------------------------------------------------------- */
/* replaces the lookup building of _vds_shared_init (lib/block.c:170-294):
 * mdct_init, drft_init are recomputed here from the block sizes; the psy
 * lookups and windows are uploaded as given.                                */
int  vb200_ctx_create(const vb200_setup *setup, int device, vb200_ctx **out);
void vb200_ctx_destroy(vb200_ctx *ctx);
int  vb200_device_count(void);
const char *vb200_last_error(void);
/* host copies of the tables the context derived itself (for parity tests):
 * which: 0 mdct trig (N+N/4), 1 mdct bitrev (N/4 int32), 2 window (N/2),
 * 3 fft twiddles (N).  Returns element count or <0.                         */
int  vb200_ctx_table(vb200_ctx *ctx, int W, int which, void *dst, int cap);
/* kernels launched through this context since creation (bench evidence)     */
uint64_t vb200_launch_count(vb200_ctx *ctx);
/* measurement aid: when on, the Phase-A entry points bracket each of their three
 * kernels (transform, ampmax, psy) with CUDA events on the launching stream;
 * vb200_phaseA_kernel_ms returns the durations of the last call (ms3[3]).       */
int  vb200_set_profiling(vb200_ctx *ctx, int on);
int  vb200_phaseA_kernel_ms(vb200_ctx *ctx, float *ms3);
/* same for the last vb200_encode_dsp_dev call: transform, ampmax, psy, floor1_fit, floor1_render,
 * couple_quantize_normalize (+ nonzero propagation)                                            */
int  vb200_encode_dsp_kernel_ms(vb200_ctx *ctx, float *ms6);
/* development aid: with env VB200_PHASE_TIMING set, the psy kernel adds the SM cycles each of
 * its 11 barrier-delimited phases took (thread 0 of every CTA) into a 16-slot counter array. */
int  vb200_debug_phase_cycles(vb200_ctx *ctx, unsigned long long *out16, int reset);

/* ---- transforms ------------------------------------- */
/* mdct_forward, lib/mdct.c:492: in [nvec][N] -> out [nvec][N/2]            */
int vb200_mdct_forward_dev (vb200_ctx*, int W, int nvec, const float *d_in, float *d_out, void *stream);
int vb200_mdct_forward     (vb200_ctx*, int W, int nvec, const float *in,   float *out);
/* mdct_backward, lib/mdct.c:396: in [nvec][N/2] -> out [nvec][N]           */
int vb200_mdct_backward_dev(vb200_ctx*, int W, int nvec, const float *d_in, float *d_out, void *stream);
int vb200_mdct_backward    (vb200_ctx*, int W, int nvec, const float *in,   float *out);
/* _vorbis_apply_window, lib/window.c:2102: in place on [nvec][N];
 * lW/nW are per-vector int32 arrays (host) or NULL for W=0                  */
int vb200_apply_window     (vb200_ctx*, int W, int nvec, const int32_t *lW, const int32_t *nW, float *data);
/* drft_forward, lib/smallft.c:1231: in place on [nvec][N], FFTPACK layout   */
int vb200_drft_forward     (vb200_ctx*, int W, int nvec, float *data);

/* ---- psychoacoustic stages, stage-isolated --------- */
/* look = blocktype + 2*W selects vb200_setup.psy[look]                      */
/* _vp_noisemask, lib/psy.c:706: logmdct [nvec][n] -> noise [nvec][n]       */
int vb200_noisemask        (vb200_ctx*, int look, int nvec, const float *logmdct, float *noise);
/* _vp_tonemask, lib/psy.c:754: logfft [nvec][n], specmax per vector         */
int vb200_tonemask         (vb200_ctx*, int look, int nvec, const float *logfft,
                            const float *global_specmax, const float *local_specmax, float *tone);
/* _vp_offset_and_mix, lib/psy.c:779: writes logmask, scales mdct in place   */
int vb200_offset_and_mix   (vb200_ctx*, int look, int nvec, int offset_select,
                            const float *noise, const float *tone,
                            float *mdct, const float *logmdct, float *logmask);

/* ---- encode Phase A: the per-channel loops of mapping0_forward
 *      (lib/mapping0.c:254-470): window, MDCT, FFT, log spectra, ampmax,
 *      noise mask, tone mask, offset_and_mix(select 1).
 * pcm     [nblocks][ch][N]   un-windowed block PCM (vb->pcm), read-only
 * desc    [nblocks]
 * mdct    [nblocks][ch][n]   gmdct after the AoTuV-M1 scaling (mapping0.c:463)
 * logmdct [nblocks][ch][n]   input of floor1_fit (mapping0.c:500)
 * logmask [nblocks][ch][n]   input of floor1_fit
 * ampmax_out [nblocks]       vorbis_block_internal.ampmax on exit (mapping0.c:576)
 * Optional taps (may be NULL): noise, tone, logfft [nblocks][ch][n], raw mdct. */
typedef struct vb200_phaseA_io {
  const float *pcm;
  const vb200_block_desc *desc;
  float *mdct;
  float *logmdct;
  float *logmask;
  float *ampmax_out;
  float *tap_noise;
  float *tap_tone;
  float *tap_logfft;
  float *tap_mdct_raw;
} vb200_phaseA_io;
int vb200_analysis_phaseA_dev(vb200_ctx*, int W, int nblocks, const vb200_phaseA_io *d_io, void *stream);
int vb200_analysis_phaseA    (vb200_ctx*, int W, int nblocks, const vb200_phaseA_io *io);

/* Stream mode: `nstreams` streams of `blocks_per_stream` consecutive blocks
 * (block index = stream*blocks_per_stream + k).  desc[].ampmax is ignored;
 * the ampmax chain of vorbis_analysis_blockout (lib/block.c:626-628,
 * lib/psy.c:837-848) is evaluated on the device between the transform and
 * the psy kernel, starting from `ampmax0[stream]` (NULL = -9999).           */
int vb200_analysis_phaseA_streams_dev(vb200_ctx*, int W, int nstreams, int blocks_per_stream,
                                      const vb200_phaseA_io *d_io, const float *d_ampmax0, void *stream);

/* ---- PCM ingest fused into Phase A -----------------------------------------
 * Each stream's PCM is ONE contiguous buffer; block k of a stream is the N samples that start
 * at sample k*hop - exactly what vorbis_analysis_blockout copies out of v->pcm (lib/block.c:630-643;
 * hop = N/2 for a run of equal-size blocks), so the 50 % overlap is never duplicated in memory.
 *   fmt VB200_PCM_F32_PLANAR : float [stream][ch][stream_stride]
 *   fmt VB200_PCM_S16_INTERLEAVED : int16 [stream][stream_stride][ch], converted sample/32768.f
 *       as examples/encoder_example.c:196-201 does
 * io->pcm is ignored; everything else as vb200_analysis_phaseA_streams_dev (ampmax chain on device,
 * d_ampmax0 may be NULL).  stream_stride counts samples per channel; it and hop must be multiples
 * of 4 for the float format.                                                                 */
#define VB200_PCM_F32_PLANAR       1
#define VB200_PCM_S16_INTERLEAVED  2
int vb200_analysis_phaseA_pcmstream_dev(vb200_ctx*, int W, int nstreams, int blocks_per_stream,
                                        const void *d_pcm, int fmt, int64_t stream_stride, int hop,
                                        const vb200_phaseA_io *d_io, const float *d_ampmax0, void *stream);

/* ---- floor 1 on the device ---------------------------------------------------
 * Rows are (block, channel) pairs.  floor_sel = -1: rows are laid out [block][channel] like the
 * Phase A outputs and row r uses the floor of channel r % channels; floor_sel >= 0: every row uses
 * floor1[W][floor_sel] (what one reference call with one vorbis_look_floor1 does).
 * All posts arrays are [rows][VB200_FLOOR1_STRIDE] int32; entries past the floor's posts are 0.
 *
 * floor1_fit (lib/floor1.c:576-729): logmdct, logmask [rows][n] -> posts exactly as the reference
 * returns them (unused posts carry predicted|0x8000); fit_nonzero[rows] = 0 where the reference
 * returns NULL (silent channel; the row of posts is then all 0).                               */
int vb200_floor1_fit_dev(vb200_ctx*, int W, int floor_sel, int nrows, const float *d_logmdct,
                         const float *d_logmask, int32_t *d_posts, int32_t *d_fit_nonzero, void *stream);
int vb200_floor1_fit    (vb200_ctx*, int W, int floor_sel, int nrows, const float *logmdct,
                         const float *logmask, int32_t *posts, int32_t *fit_nonzero);
/* the part of floor1_encode that is not bit packing (lib/floor1.c:765-832, 919-945): quantise the
 * posts to the multiplier, apply the prediction/flag pass, render the integer floor curve.
 * posts is updated in place to what the reference leaves in post[]; ilogmask [rows][n] is the curve
 * (all zero and nonzero = 0 where fit_nonzero is 0).                                            */
int vb200_floor1_render_dev(vb200_ctx*, int W, int floor_sel, int nrows, int32_t *d_posts,
                            const int32_t *d_fit_nonzero, int32_t *d_ilogmask, int32_t *d_nonzero,
                            void *stream);
int vb200_floor1_render    (vb200_ctx*, int W, int floor_sel, int nrows, int32_t *posts,
                            const int32_t *fit_nonzero, int32_t *ilogmask, int32_t *nonzero);

/* ---- encode Phase B: _vp_couple_quantize_normalize, lib/psy.c:1014 ----
 * mdct  [nblocks][ch][n]  (Phase A output)
 * iwork [nblocks][ch][n]  in: ilogmask from floor1_encode (0..1023 dB index,
 *                         mapping0.c:617), out: quantised residue ints
 * nonzero [nblocks][ch]   in/out                                            */
int vb200_couple_quantize_normalize_dev(vb200_ctx*, int W, int blocktype, int blobno, int nblocks,
                                        const float *d_mdct, int32_t *d_iwork, int32_t *d_nonzero, void *stream);
int vb200_couple_quantize_normalize    (vb200_ctx*, int W, int blocktype, int blobno, int nblocks,
                                        const float *mdct, int32_t *iwork, int32_t *nonzero);

/* ---- residue partition classification: res1_class / res2_class (lib/res0.c:745-778)
 * -> _01class (:412-474) / _2class (:479-532), called per submap as mapping0_forward does (:660-672).
 * iwork   [nblocks][ch][n] quantised residue (what vb200_couple_quantize_normalize leaves)
 * nonzero [nblocks][ch]
 * classes [nblocks][ch][class_stride] int32 partition classes (the reference's partword):
 *   residue type 0/1: row of channel c holds partword of that channel, valid where nonzero[c] (the
 *     reference compacts the used channels of a submap: partword[u] = row of the u-th used channel);
 *   residue type 2: ONE vector per submap, stored in the row of the submap's first channel, valid if any
 *     channel of the submap is nonzero (the reference then classifies the whole bundle, :769-778).
 *   Rows that are not valid and entries past partvals = (end-begin)/grouping are 0.
 * class_stride >= vb200_residue_partvals(ctx, W) (the largest partvals over the mode's submaps).    */
int vb200_residue_partvals(vb200_ctx*, int W);
int vb200_residue_classify_dev(vb200_ctx*, int W, int nblocks, const int32_t *d_iwork, const int32_t *d_nonzero,
                               int32_t *d_classes, int class_stride, void *stream);
int vb200_residue_classify    (vb200_ctx*, int W, int nblocks, const int32_t *iwork, const int32_t *nonzero,
                               int32_t *classes, int class_stride);

/* ---- the whole per-block encode DSP of mapping0_forward in ONE call ---------------------------
 * lib/mapping0.c:230-646 with the bit packing and the residue backend left to the caller:
 *   window, MDCT, FFT, log spectra, ampmax (:254-346)  ->  noise / tone masks, offset_and_mix(1)
 *   (:366-470)  ->  floor1_fit (:500)  ->  the non-bit-packing part of floor1_encode (:617)
 *   ->  _vp_couple_quantize_normalize (:631-646)
 * for un-managed bitrate (only blob `blobno` = PACKETBLOBS/2 is produced, :592-594).  The float
 * spectra never leave the device: per stereo long block 4 KB of int16 PCM go in and 8.4 KB come
 * back (posts, nonzero, quantised residue) instead of 16 KB + 24.6 KB for Phase A alone.
 *
 * Blocks are `nstreams` x `blocks_per_stream`, block index = stream*blocks_per_stream + k.
 *   pcm_fmt VB200_PCM_F32_BLOCKS      float [nblocks][ch][N]  (vb->pcm as blockout leaves it)
 *           VB200_PCM_F32_PLANAR      float [stream][ch][stream_stride], block k starts at k*hop
 *           VB200_PCM_S16_INTERLEAVED int16 [stream][stream_stride][ch], sample/32768.f
 *   desc[nblocks]      lW, nW, blocktype of every block (psy look = blocktype + 2W, :250)
 *   independent != 0   desc[].ampmax is each block's ampmax on entry (what one vorbis_analysis
 *                      call sees); == 0: the decay chain of vorbis_analysis_blockout
 *                      (lib/block.c:626-628) runs per stream from ampmax0[stream] (NULL = -9999)
 * Outputs (what floor1_encode's bit packer and the residue backend consume):
 *   posts   [nblocks][ch][VB200_FLOOR1_STRIDE]  post[] as floor1_encode leaves it (:765-832)
 *   nonzero [nblocks][ch]                       after the coupling propagation (lib/psy.c:1203)
 *   iwork   [nblocks][ch][n]                    quantised, coupled residue ints
 *   ampmax_out [nblocks]                        vorbis_block_internal.ampmax on exit (:576)
 *   mdct, logmdct, logmask [nblocks][ch][n]     optional (NULL = stay in device scratch)
 * iwork_fmt VB200_IWORK_S16 halves the largest output: the residue leaves as int16, saturated to
 * [-32768,32767]; overflow[block] counts the values of that block that were clipped (0 for any
 * realistic signal: |mdct|/floor would have to exceed 32767) - a caller that sees a non-zero
 * count re-runs that block with VB200_IWORK_S32.                                                */
#define VB200_PCM_F32_BLOCKS       0
#define VB200_IWORK_S32            0
#define VB200_IWORK_S16            1
typedef struct vb200_encode_io {
  const void *pcm;
  int32_t pcm_fmt;
  int32_t hop;
  int64_t stream_stride;
  const vb200_block_desc *desc;
  const float *ampmax0;
  int32_t independent;
  int32_t iwork_fmt;              /* VB200_IWORK_S32 (0) or VB200_IWORK_S16 */
  int32_t *posts;
  int32_t *nonzero;
  void    *iwork;                 /* int32 or int16 [nblocks][ch][n], see iwork_fmt */
  float *ampmax_out;
  float *mdct;
  float *logmdct;
  float *logmask;
  int32_t *overflow;              /* VB200_IWORK_S16 only: [nblocks] values that did not fit */
  int32_t *classes;               /* optional [nblocks][ch][class_stride]: vb200_residue_classify of the residue */
  int64_t class_stride;
} vb200_encode_io;
/* every pointer in *d_io is a device pointer; the struct itself is host memory */
int vb200_encode_dsp_dev(vb200_ctx*, int W, int nstreams, int blocks_per_stream, int blobno,
                         const vb200_encode_io *d_io, void *stream);
/* host buffers (pinned memory makes the copies asynchronous); whole streams are cut into chunks
 * that rotate over four device buffer sets: all H2D copies on one stream, all D2H copies on
 * another, the kernels of consecutive chunks on two compute streams, ordered by events         */
int vb200_encode_dsp    (vb200_ctx*, int W, int nstreams, int blocks_per_stream, int blobno,
                         const vb200_encode_io *io);

/* ---- bitrate-managed mode ------------------------------------------------------
 * What mapping0_forward does when vorbis_bitrate_managed(vb) (lib/mapping0.c:507-573, 596-646):
 * Phase A as above, then besides the middle mask (offset_select 1) the low-noise and the
 * high-noise masks (_vp_offset_and_mix selections 2 and 0, lib/psy.c:779-835), a floor1_fit on
 * each (only where the middle fit exists), the twelve floor1_interpolate_fit curves in between
 * (lib/floor1.c:731-757; a curve exists only where both of its ends do), and for EVERY one of the
 * VB200_PACKETBLOBS curves k: the floor render (floor1_encode minus the bits, :765-945) and
 * _vp_couple_quantize_normalize with blob k's coupling parameters and sliding low-pass.
 * Same vb200_encode_io as vb200_encode_dsp with blob-major outputs:
 *   posts   [VB200_PACKETBLOBS][nblocks*ch][VB200_FLOOR1_STRIDE]   (all zero where the curve is NULL)
 *   nonzero [VB200_PACKETBLOBS][nblocks*ch]
 *   iwork   [VB200_PACKETBLOBS][nblocks*ch][n]                     (iwork_fmt must be VB200_IWORK_S32)
 * classes / overflow must be NULL.  The host keeps what it keeps in un-managed mode: the bits of
 * floor1_encode, residue coding, and the bitrate manager's choice among the 15 packets.          */
int vb200_encode_dsp_managed_dev(vb200_ctx*, int W, int nstreams, int blocks_per_stream,
                                 const vb200_encode_io *d_io, void *stream);
int vb200_encode_dsp_managed    (vb200_ctx*, int W, int nstreams, int blocks_per_stream,
                                 const vb200_encode_io *io);

/* ---- envelope / block-switch detector -----------------------------------------
 * The analysis loop of _ve_envelope_search (lib/envelope.c:232-267): for steps j = first_step ..
 * first_step+nsteps-1 of every stream and every channel, _ve_amp (:88-213) on the 128 samples that
 * start at sample 64*j - squared-sine window, mdct_forward(128), near-DC spreading, the seven band
 * amplitudes, their pre-/post-echo deltas against the 17-deep amplitude history - and the stretch
 * logic that couples the channels.  Everything that vorbis_analysis_blockout then does with the
 * result (cursor / curmark walk, :269-327) only reads the marks and stays on the host.
 *
 * pcm     stream PCM, VB200_PCM_F32_PLANAR [stream][ch][stream_stride] or VB200_PCM_S16_INTERLEAVED
 *         [stream][stream_stride][ch]; 64*(first_step+nsteps-1)+128 <= stream_stride
 * state   [nstreams][VB200_VE_STATE_WORDS(ch)] 32-bit words, in/out: word 0 = envelope_lookup.stretch,
 *         then envelope_filter_state[ch*VE_BANDS] with the reference's own layout (lib/envelope.h:32-42:
 *         ampbuf[17], ampptr, nearDC[15], nearDC_acc, nearDC_partialacc, nearptr) so that a binding
 *         can copy ve->filter in and out verbatim; all zero = a fresh vorbis_dsp_state
 * ret     [nstreams][nsteps] the loop's `ret` per step: bit 0 (and 2) pre-echo, bit 1 post-echo
 * vb200_envelope_apply_marks (plain C, no CUDA) replays lib/envelope.c:254-264 on a mark array.   */
#define VB200_VE_FILTER_WORDS 36
#define VB200_VE_STATE_WORDS(ch) (1 + VB200_VE_FILTER_WORDS * VB200_VE_BANDS * (ch))
int vb200_envelope_search_dev(vb200_ctx*, int nstreams, const void *d_pcm, int pcm_fmt, int64_t stream_stride,
                              int first_step, int nsteps, int32_t *d_state, uint8_t *d_ret, void *stream);
int vb200_envelope_search    (vb200_ctx*, int nstreams, const void *pcm, int pcm_fmt, int64_t stream_stride,
                              int first_step, int nsteps, int32_t *state, uint8_t *ret);
void vb200_envelope_apply_marks(const uint8_t *ret, int first_step, int nsteps, int32_t *mark);
/* many streams with DIFFERENT amounts of new data in one call (the multi-stream driver of
 * vorbis_b200/host/vb200_mapping0.c): stream s analyses steps 0 .. steps_per_stream[s]-1 (<= nsteps) of its
 * buffer; ret [nstreams][nsteps], entries past a stream's own count are left untouched.  Host pointers.   */
int vb200_envelope_search_var(vb200_ctx*, int nstreams, const void *pcm, int pcm_fmt, int64_t stream_stride,
                              int nsteps, const int32_t *steps_per_stream, int32_t *state, uint8_t *ret);

/* ---- block planning: what vorbis_analysis_blockout decides per block ----------------
 * lib/block.c:556-615 (nW from the envelope marks, blocktype) with the cursor / curmark walk of
 * _ve_envelope_search (lib/envelope.c:269-327) and _ve_envelope_mark (:329-356), replayed per stream
 * on the stream's timeline buffer: sample 0 is v->pcm[][0] of a fresh vorbis_dsp_state, i.e. the
 * blocksizes[1]/2 samples of (pre-extrapolated) preamble, then the caller's PCM, then - after EOF -
 * the extrapolated tail of vorbis_analysis_wrote(v,0).
 *   mark   [nstreams][mark_stride] int32, mark[j] = envelope_lookup.mark of the 64-sample step j counted
 *          from sample 0 (vb200_envelope_search + vb200_envelope_apply_marks over the whole timeline)
 *   nsteps steps analysed (the reference's `last` = pcm_len/64 - 4 when the whole timeline was searched)
 *   pcm_len[nstreams]  samples present per stream (v->pcm_current without any shift)
 *   eof    [nstreams]  v->eofflag in timeline samples (preamble + samples written) or 0: no EOF yet
 *   plan   [nstreams][max_blocks] blocks in stream order; nblocks[nstreams] how many (the stream stops
 *          where vorbis_analysis_blockout would return 0)
 * block sizes: blocksizes[0]/4 must be a multiple of 64 (the envelope search step).               */
typedef struct vb200_stream_block {
  int32_t pos;        /* first sample of the block on the timeline (centerW - blocksizes[W]/2) */
  int32_t slot;       /* index of the block among the blocks of its size in the whole call */
  int32_t W, lW, nW;  /* vb->W, vb->lW, vb->nW */
  int32_t blocktype;  /* vorbis_block_internal.blocktype (psy look = blocktype + 2W) */
} vb200_stream_block;
int vb200_plan_blocks(vb200_ctx*, int nstreams, const int32_t *mark, int64_t mark_stride, int nsteps,
                      const int64_t *pcm_len, const int64_t *eof, int max_blocks,
                      vb200_stream_block *plan, int32_t *nblocks);

/* ---- whole streams in ONE call (SURVEY §8 a12 + a15): envelope search, block planning, then the
 * per-block encode DSP of both block sizes with the ampmax decay chain carried along each stream
 * across sizes (lib/block.c:626-628).  The blocks of size W of all streams form one batch; block
 * `slot` of that batch writes posts[W][slot], nonzero[W][slot], iwork[W][slot], ampmax_out[W][slot].
 *   pcm       timeline buffers (see vb200_plan_blocks), VB200_PCM_F32_PLANAR [stream][ch][stream_stride]
 *             or VB200_PCM_S16_INTERLEAVED [stream][stream_stride][ch]
 *   cap[W]    capacity (blocks) of the size-W outputs; count[W] (host, out) blocks produced
 *   plan / nblocks as vb200_plan_blocks (device pointers for the _dev form)
 * Un-managed bitrate (blob `blobno`).  Streams start fresh (zero envelope state, ampmax -9999).   */
typedef struct vb200_streams_io {
  const void *pcm;
  int32_t pcm_fmt;
  int32_t max_blocks;
  int64_t stream_stride;
  const int64_t *pcm_len;      /* [nstreams] */
  const int64_t *eof;          /* [nstreams] or NULL (no EOF anywhere) */
  vb200_stream_block *plan;    /* [nstreams][max_blocks] out */
  int32_t *nblocks;            /* [nstreams] out */
  int32_t cap[2];
  int32_t count[2];            /* out (host) */
  int32_t *posts[2];           /* [cap][ch][VB200_FLOOR1_STRIDE] */
  int32_t *nonzero[2];         /* [cap][ch] */
  int32_t *iwork[2];           /* [cap][ch][blocksizes[W]/2] */
  float   *ampmax_out[2];      /* [cap] */
} vb200_streams_io;
int vb200_encode_streams_dev(vb200_ctx*, int nstreams, int blobno, vb200_streams_io *d_io, void *stream);
int vb200_encode_streams    (vb200_ctx*, int nstreams, int blobno, vb200_streams_io *io);

/* ---- decode: mdct_backward (lib/mapping0.c:792-795) fused with the windowed
 *      overlap-add of vorbis_synthesis_blockin (lib/block.c:767-823).
 * `nstreams` independent streams x `ch` channels, each `nblk` blocks long.
 * Wseq  [nstreams][nblk] int32 block-size flags
 * coef_off [nstreams][nblk] int64 offset (floats) of block k's spectra inside
 *       `coef`; channel c of that block is at coef_off + c*n_k
 * pcm_off  [nstreams][nblk] int64 offset (floats) into each channel's output
 *       where the samples finished by block k go (k=0 finishes nothing)
 * pcm   [nstreams][ch][pcm_stride]                                          */
int vb200_synthesis_dev(vb200_ctx*, int nstreams, int nblk, const int32_t *d_Wseq,
                        const int64_t *d_coef_off, const float *d_coef,
                        const int64_t *d_pcm_off, float *d_pcm, int64_t pcm_stride, void *stream);
/* same, but the finished samples leave as interleaved int16 [stream][pcm_stride][ch]:
 * floor(x*32767.f+.5f) clipped to [-32768,32767] (examples/decoder_example.c:250-262)          */
int vb200_synthesis_s16_dev(vb200_ctx*, int nstreams, int nblk, const int32_t *d_Wseq,
                            const int64_t *d_coef_off, const float *d_coef,
                            const int64_t *d_pcm_off, int16_t *d_pcm16, int64_t pcm_stride, void *stream);
int vb200_synthesis    (vb200_ctx*, int nstreams, int nblk, const int32_t *Wseq,
                        const int64_t *coef_off, const float *coef, int64_t coef_len,
                        const int64_t *pcm_off, float *pcm, int64_t pcm_stride);

/* ---- decode: channel de-coupling of mapping0_inverse (lib/mapping0.c:754-779).
 * res [nblocks][ch][n] residue vectors as left by the residue backend; every coupling step
 * (magnitude, angle) -> (left, right) is undone in place, last step first.  Elementwise.  */
int vb200_decouple_dev(vb200_ctx*, int W, int nblocks, float *d_res, void *stream);
int vb200_decouple    (vb200_ctx*, int W, int nblocks, float *res);

/* ---- decode: floor1_inverse2 (lib/floor1.c:1041-1086), the floor curve multiplied into the spectrum.
 * data [rows][n] in place; posts [rows][VB200_FLOOR1_STRIDE] = fit_value[] as floor1_inverse1 returns it
 * (:962-1039; unused posts carry bit 15); present[rows] = 0 where floor1_inverse1 returned NULL (the row
 * is zeroed, :1084).  Rows and floor_sel as for vb200_floor1_fit.                                        */
int vb200_floor1_inverse2_dev(vb200_ctx*, int W, int floor_sel, int nrows, const int32_t *d_posts,
                              const int32_t *d_present, float *d_data, void *stream);
int vb200_floor1_inverse2    (vb200_ctx*, int W, int floor_sel, int nrows, const int32_t *posts,
                              const int32_t *present, float *data);

/* ---- decode: everything of mapping0_inverse after the entropy decoders (lib/mapping0.c:754-795) plus
 * the overlap-add of vorbis_synthesis_blockin, in one call: channel de-coupling, floor multiply,
 * mdct_backward, windowed overlap-add.  Layout as vb200_synthesis; `res` holds the residue vectors as
 * the residue backend leaves them and is modified in place; posts / present are
 * [nstreams][nblk][ch][VB200_FLOOR1_STRIDE] and [nstreams][nblk][ch].  pcm_s16 != 0: the finished
 * samples leave as interleaved int16 (as vb200_synthesis_s16_dev), else planar float.                   */
int vb200_decode_dsp_dev(vb200_ctx*, int nstreams, int nblk, const int32_t *d_Wseq, const int64_t *d_coef_off,
                         float *d_res, const int32_t *d_posts, const int32_t *d_present,
                         const int64_t *d_pcm_off, void *d_pcm, int pcm_s16, int64_t pcm_stride, void *stream);
int vb200_decode_dsp    (vb200_ctx*, int nstreams, int nblk, const int32_t *Wseq, const int64_t *coef_off,
                         float *res, int64_t res_len, const int32_t *posts, const int32_t *present,
                         const int64_t *pcm_off, void *pcm, int pcm_s16, int64_t pcm_stride);

/* ---- device memory helpers for non-CUDA hosts (C callers) -------------- */
int  vb200_malloc_device(vb200_ctx*, size_t bytes, void **dptr);
int  vb200_free_device  (vb200_ctx*, void *dptr);
int  vb200_memcpy_h2d   (vb200_ctx*, void *dptr, const void *src, size_t bytes);
int  vb200_memcpy_d2h   (vb200_ctx*, void *dst, const void *dptr, size_t bytes);
int  vb200_synchronize  (vb200_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* VORBIS_B200_H */
