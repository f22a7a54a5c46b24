/* flite_b200.h -- C ABI of libflite_b200.so: hand-written sm_100a kernels for the F Lite denoise step.
 *
 * The reference (sippycoder/f-lite) has no native layer: every op below replaces a
 * Python call site in /root/reference/f_lite/model.py or the sampler loop in
 * /root/reference/f_lite/pipeline.py; the citation on each entry point is the reference line(s) whose
 * arithmetic it reproduces.  See INTEGRATION.md for the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - all pointers are DEVICE pointers (bf16 unless stated), caller-owned; the library allocates nothing
 *   - ld* are row strides in ELEMENTS; rows must be 16-byte aligned
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never synchronise and are
 *     CUDA-graph capturable
 *   - return 0 on success, a negative FLITE_ERR_* otherwise; flite_last_error() gives the reason
 *   - sm_100a only; there is no fallback path
 */
#ifndef FLITE_B200_H
#define FLITE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLITE_ABI_VERSION 1

#define FLITE_OK 0
#define FLITE_ERR_INVALID (-1)   /* bad argument (shape / alignment / null)             */
#define FLITE_ERR_CUDA (-2)      /* CUDA runtime or driver error                         */
#define FLITE_ERR_UNSUPPORTED (-3) /* device is not sm_100 / required driver symbol missing */
#define FLITE_ERR_WATCHDOG (-4)  /* a kernel barrier wait timed out (protocol bug)        */

/* GEMM epilogues (flite_gemm_bf16) */
#define FLITE_EPI_STORE 0      /* C = act(A W^T + bias)                         model.py:162,189,436,448-454,472,475 */
#define FLITE_EPI_GATED_RES 1  /* C = resid + (A W^T + bias) * gate[sample]     model.py:289,294-297,301              */
#define FLITE_EPI_SWIGLU 2     /* C = silu(A Wg^T) * (A Wu^T), W = interleave64(Wg, Wu)  model.py:267 (LigerSwiGLUMLP) */
#define FLITE_EPI_QKV_ROPE 3   /* +bias, 2-D RoPE, per-head RMSNorm on q/k heads  model.py:162-183,403-414,92-108      */

/* GEMM kernel variants */
#define FLITE_GEMM_AUTO 0
#define FLITE_GEMM_1CTA_N256 1
#define FLITE_GEMM_2CTA_N256 2
#define FLITE_GEMM_1CTA_N128 3
#define FLITE_GEMM_1CTA_N64 4
#define FLITE_GEMM_GEMV 5        /* M <= 8, plain epilogue: one warp per output column, weight streaming (HBM-bound); AUTO picks it */

/* attention kernel variants */
#define FLITE_ATTN_AUTO 0
#define FLITE_ATTN_1WG 1   /* one softmax warpgroup (192 threads)                 */
#define FLITE_ATTN_2WG 2   /* two softmax warpgroups splitting the key columns    */
#define FLITE_ATTN_2CTA_1WG 3 /* cta_group::2 pair sharing K/V, double-buffered, 1 softmax warpgroup */
#define FLITE_ATTN_2CTA_2WG 4 /* same with two softmax warpgroups                                           */
#define FLITE_ATTN_2CTA_1WG_PTMEM 5 /* cta_group::2, P kept in TMEM (A operand of the PV MMA read from TMEM); default */
#define FLITE_ATTN_2CTA_2WG_PTMEM 6
#define FLITE_ATTN_QTMEM_1WG 7 /* cta_group::2, Q and P both TMEM operands, 64-key tiles, 6-stage K/V ring */
#define FLITE_ATTN_QTMEM_2WG 8
#define FLITE_ATTN_PERSISTENT 10  /* one wave of 2-CTA clusters, whole (sequence, head, 256-query tile) units round-robin, ragged lengths: variant 5's arithmetic without the per-unit launch cost (bit-identical to it) */
#define FLITE_ATTN_XRES 9        /* persistent cross-attention, K/V (<= 256 keys per sequence, checked on the device) resident in the CTA pair */

/* tuning knobs (A/B switches used by the benchmarks; defaults are the measured best) */
#define FLITE_TUNE_RMSNORM_KERNEL 0  /* 0 auto | 1 two-pass | 2 register-resident | 3 streaming (persistent warps, next-row prefetch) */
#define FLITE_TUNE_ATTN_VARIANT 1    /* default attention variant when the call passes FLITE_ATTN_AUTO */
#define FLITE_TUNE_GEMM_VARIANT 2    /* default GEMM variant when the call passes FLITE_GEMM_AUTO (0 = heuristic) */
#define FLITE_TUNE_ATTN_DEBUG 3      /* profiling experiments only: bit0 skip softmax math, bit1 skip K/V reloads */
#define FLITE_TUNE_GEMM_TAIL_SPLIT 4 /* 0 = run a last partial wave as half-width tiles (default), 1 = off */
#define FLITE_TUNE_QKV_STAGED_STORES 6 /* 1 = QKV epilogue stores whole row segments via a shared-memory transpose (always on for peer stores) */
#define FLITE_TUNE_ATTN_STAGED_STORES 7 /* 1 = 2-CTA attention stores whole output rows via a shared-memory transpose (always on for peer stores) */
#define FLITE_TUNE_PDL 8             /* 1 = GEMM / attention / rmsnorm kernels use programmatic dependent launch; default 0 = off:
                                        measured 1 % SLOWER at C2 (profiles/r1e_ab_pdl.json) -- the step is power-capped, idle gaps are free */
#define FLITE_TUNE_GEMM_HINT_A 9      /* L2 eviction hint of the GEMM's A-tile TMA loads: 0 auto | 1 none | 2 evict_first | 3 evict_last */
#define FLITE_TUNE_GEMM_HINT_B 10     /* same for the W-tile loads */
#define FLITE_TUNE_PATCH_EMBED 11     /* 0 auto: patchify = gather + tcgen05 GEMM when C*P*P % 64 == 0 | 1 CUDA-core patch_embed kernel */
#define FLITE_TUNE_ATTN_VARIANT_SHORT_K 12 /* attention variant for FLITE_ATTN_AUTO calls with <= 512 keys per sequence on average (cross-attention); 0 = same as the long-key default (the persistent kernel); 9 (FLITE_ATTN_XRES) = the host model requests the round-1 resident-K/V kernel when the padded context has <= 256 tokens (A/B only: slower than the default now) */
#define FLITE_TUNE_P2P_TIMEOUT_S 13    /* seconds a cross-rank flag wait (flite_p2p_wait) may spin before it aborts; 0 = default 120 */
#define FLITE_TUNE_GEMM_NARROW_M 14    /* 0 = a last M-tile with <= 128 valid rows runs as M = 128 MMAs (default), 1 = padded 256-row tile */
#define FLITE_TUNE_ATTN_SK_MODE 15     /* schedule of flite_attention_streamk: 0 = stream-K shares | 1 = whole units round-robin (persistent, never splits a unit) | 2 = hybrid (whole rounds in lock step, stream-K over the last 1..2 units per cluster) */
#define FLITE_TUNE_ATTN_TMA_OUT 16     /* 0 = the 2-CTA attention kernels store whole 128-row output tiles through shared memory + TMA (default), 1 = per-thread row stores */
#define FLITE_TUNE_GEMM_DEBUG 17       /* profiling experiments only: bit0 = gated-residual epilogue without its global loads / stores */
#define FLITE_TUNE_GEMM_BAND 5       /* 0 = L2-aware band rasterisation for large M (default), 1 = single band */
int flite_set_tuning(int key, int value);
int flite_get_tuning(int key);   /* current value of a knob (0 for an unknown key) */

int flite_abi_version(void);
const char* flite_last_error(void);

/* 0 if cuda:current is sm_100 and the driver exports cuTensorMapEncodeTiled, else FLITE_ERR_UNSUPPORTED */
int flite_check_device(void);

/* Synchronises the device and reports (and clears) the kernel watchdog word; 0 = no barrier timed out. */
int flite_watchdog_status(unsigned int* code_out);

/* Sampler: fused CFG combine + Euler update.            pipeline.py:290,296-297 / train.py:596,599
 *   v = u + g*(c-u) (skipped if !do_cfg: v = c);  acc += dt*v;  lat_out = bf16(acc)
 *   acc is bf16 (acc_is_fp32 = 0, FLitePipeline semantics) or fp32 (= 1, train.py sample_images). */
int flite_cfg_euler(void* acc, int acc_is_fp32, const void* v_uncond, const void* v_cond, float guidance,
                    float dt, int do_cfg, void* lat_out, int64_t numel, void* stream);

/* Sampler: Augmented Parallel Guidance combine + Euler update.           pipeline.py:276-287,296-297
 *   dy = c; dd = c-u; par = (dy.dd)/(dy.dy)*dy (global sums over the whole tensor); orth = dd-par;
 *   orth *= min(1, threshold/std(orth)); v = dy + (g-1)*orth; acc += dt*v; lat_out = bf16(acc)
 *   every torch rounding point of the model dtype (bf16) is reproduced; the three global reductions are evaluated on
 *   the device (3 stream-ordered launches, no host sync -- the reference syncs in `min(1, tensor)`).
 *   workspace: flite_apg_workspace_bytes() bytes of device memory, 8-byte aligned, owned by the caller. */
int flite_apg_workspace_bytes(void);
int flite_apg_euler(void* acc, int acc_is_fp32, const void* v_uncond, const void* v_cond, float guidance, float dt,
                    float orthogonal_threshold, void* lat_out, int64_t numel, void* workspace, void* stream);

/* Pipeline tail.  latent_unscale: out = lat/scaling_factor + shift_factor (bf16)                 pipeline.py:304
 *   image_to_uint8: (x/2+0.5).clamp(0,1)*255 -> round -> uint8, decoded [B,C,H,W] (bf16 or fp32) -> out [B,H,W,C]
 *   (pipeline.py:324-327; the reference keeps NCHW and permutes each image on the host) */
int flite_latent_unscale(const void* latents, void* out, float scaling_factor, float shift_factor, int64_t numel,
                         void* stream);
int flite_image_to_uint8(const void* decoded, int in_is_fp32, void* out_u8, int B, int C, int H, int W, void* stream);

/* GroupNorm (+ SiLU) on channels-last activations x[N, HW, C] (bf16): the norm -> activation pairs of the VAE decoder the
 * pipeline decodes with (diffusers AutoencoderKL.decode behind f_lite/pipeline.py:299-307: ResnetBlock2D.norm1/norm2 +
 * nonlinearity, Attention.group_norm, Decoder.conv_norm_out + conv_act).  Two launches (per-slice partial sums in a fixed
 * order, then normalise): deterministic, statistics in fp32 / double, GroupNorm output rounded to bf16 before the SiLU like
 * the torch op pair.  `partials` is caller-owned scratch of flite_groupnorm_partials_bytes(N, groups, splits) bytes.
 * Requires C % 8 == 0, (C / groups) % 4 == 0, groups <= 64, C <= 2048, 16-byte aligned pointers. */
int64_t flite_groupnorm_partials_bytes(int N, int groups, int splits);
/* y[rows, C] (channels-last conv output, bf16) = bf16(y + bias[C]); if residual != NULL then y = bf16(residual + y): the
 * bias of the decoder's convolutions (torch applies it as a separate broadcast add_ on the cuDNN output) and the
 * ResnetBlock2D skip connection (diffusers resnet.py: output = input + hidden) in one in-place pass. */
int flite_bias_residual_add_nhwc(void* y, const void* bias, const void* residual, int64_t rows, int C, void* stream);
/* y[N, 2H, 2W, C] = nearest-neighbour 2x upsampling of the channels-last x[N, H, W, C] (bf16; diffusers Upsample2D:
 * F.interpolate(scale_factor=2.0, mode="nearest") before its convolution).  C % 8 == 0. */
int flite_upsample_nearest2x_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream);
int flite_groupnorm_silu_nhwc(const void* x, void* y, const void* gamma, const void* beta, int N, int64_t HW, int C,
                              int groups, float eps, int apply_silu, void* partials, int splits, void* stream);

/* y = RMSNorm(x)[*w] [*(1+scale[s]) + shift[s]], s = row / rows_per_sample.   model.py:238,283-284,292-293,299-300,437,579-580
 *   weight_mode 0 none | 1 Liger "llama" casting | 2 reference RMSNorm (fp32 weight multiply)
 *   scale/shift may be NULL (no modulation); they index a [B, ld_mod] modulation matrix */
int flite_rmsnorm_modulate(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w, int weight_mode,
                           const void* scale, const void* shift, int64_t ld_mod, int rows_per_sample,
                           int rows, int d, float eps, void* stream);

/* In-place RoPE + QK-RMSNorm on head slots [0, n_slots) of buf[rows, ld] (head_dim 256).  model.py:166-180
 *   cos/sin: bf16 [rows_per_sample, 128] (the values the reference's bf16 buffers hold) or NULL (no rotation,
 *   cross-attention, model.py:197) */
int flite_rope_qknorm(void* buf, int64_t ld, int rows, int n_slots, const void* cos_t, const void* sin_t,
                      int rows_per_sample, float eps, void* stream);

/* Patch embedding + register tokens -> tokens[B*tok_count, d].                             model.py:318-328,535
 *   Rows are (sample, local token) for sequence positions [tok_offset, tok_offset + tok_count) of the
 *   n_reg + hw tokens (tok_count <= 0: the whole sequence; a slice is what a sequence-parallel rank owns). */
int flite_patch_embed(const void* x, const void* w, const void* bias, const void* reg_tokens, void* out,
                      int B, int C, int H, int W, int P, int d, int n_reg, int tok_offset, int tok_count,
                      void* stream);

/* Patchify on the tensor cores: builds the im2col matrix A [B * n_img, C*P*P] (k order (c, p1, p2), n_img = image tokens
 * of the slice [tok_offset, tok_offset + tok_count)) and copies the register-token rows of the slice into `out`
 * [B * tok_count, d]; the caller then runs flite_gemm_bf16(A_b, conv weight viewed as [d, C*P*P], +bias) into the image
 * rows of every sample.  Replaces Conv2d k = s = P + cat(register_tokens) (model.py:324-328,535) like flite_patch_embed,
 * which does the projection on CUDA cores and is kept for C*P*P not a multiple of 64. */
int flite_patch_gather(const void* x, const void* reg_tokens, void* A, void* out, int B, int C, int H, int W, int P,
                       int d, int n_reg, int tok_offset, int tok_count, void* stream);

/* dst[n1, n0, n2] = src[n0, n1, n2] (bf16, n2 % 8 == 0): layout transform around the Ulysses all-to-alls. */
int flite_permute_021(const void* src, void* dst, int n0, int n1, int n2, void* stream);

/* Sinusoidal timestep embedding; t is fp32 on device.                                    model.py:20-28,551
 *   t_is_bf16: 0 = fp32 timesteps; 1 = bf16 timesteps (reproduces `timesteps*1000` rounded to bf16);
 *              2 = t already holds float(timesteps*1000) computed in the caller's dtype */
int flite_timestep_embed(const float* t, int t_is_bf16, const float* freqs, void* out, int B, int d,
                         void* stream);

/* tokens[B*L, ldt] (first P*P*C columns) -> out (B, C, H, W), register rows dropped.      model.py:577,583-590 */
int flite_unpatchify(const void* tok, int64_t ldt, void* out, int B, int C, int H, int W, int P, int n_reg,
                     void* stream);

/* Varlen packing of context rows by a 0/1 mask, no host sync.                            model.py:31-64,530
 *   mask fp32 [B, Lc]; pos_ws int32 [B*Lc] and seqlens_ws int32 [B] are workspaces;
 *   cu_seqlens int32 [B+1] out; dst [>= B*Lc, ldd] must be zero-initialised by the caller */
int flite_pack_context(const void* src, int64_t lds, void* dst, int64_t ldd, const float* mask, int B, int Lc,
                       int d, int* pos_ws, int* seqlens_ws, int* cu_seqlens, void* stream);

/* C[M, *] = epilogue(A[M,K] W[N,K]^T).  bf16 in, fp32 accumulate in TMEM (tcgen05), bf16 out.
 *   Requirements: K % 64 == 0, N % 64 == 0 (N % 256 == 0 for SWIGLU / QKV_ROPE), lda/ldw/ldc % 8 == 0.
 *   bias [N] or NULL; act: 0 none, 1 SiLU (EPI_STORE only)
 *   EPI_GATED_RES: resid [M, ldr], gate row = gate + (row / rows_per_sample) * ld_gate
 *   EPI_SWIGLU   : C has N/2 columns
 *   EPI_QKV_ROPE : columns [0, qk_cols) are 256-wide heads that get RoPE (rope_cos/sin bf16
 *                  [rows_per_sample, 128], may be NULL) and RMSNorm(eps); the rest is bias only.
 *                  sp_ranks > 0 (Ulysses): N = 3*d; head h of q|k|v is written into the all-to-all send layout
 *                  C[(sample*sp_ranks + h / sp_heads_per_rank)*rows_per_sample + local_row,
 *                    which*sp_heads_per_rank*256 + (h % sp_heads_per_rank)*256 ...], ldc = 3*sp_heads_per_rank*256 */
int flite_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                    int K, const void* bias, int act, int epilogue, const void* resid, int64_t ldr,
                    const void* gate, int64_t ld_gate, int rows_per_sample, const void* rope_cos,
                    const void* rope_sin, int qk_cols, float eps, int sp_ranks, int sp_heads_per_rank, int variant,
                    void* stream);

/* Varlen non-causal flash attention, head_dim 256.                                        model.py:203-211
 *   q[rows_q, ldq] head h at columns q_col0 + 256 h (same for k, v); cu_q / cu_k int32 [B+1] on device;
 *   out[rows_q, ldo] head-major columns; max_q = longest query sequence (host value).
 *   variant FLITE_ATTN_AUTO = FLITE_ATTN_PERSISTENT: one wave of 2-CTA clusters walks whole (sequence, head, 256-query
 *   tile) units round-robin with the per-sequence lengths of cu_q / cu_k (bit-identical to variant 5, one cluster per
 *   unit); whole 128-row output tiles leave through shared memory + TMA stores when `out` is 16-byte aligned. */
int flite_attention_varlen(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                           int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out,
                           int64_t ldo, const int* cu_q, const int* cu_k, int B, int H, int max_q,
                           float softmax_scale, int variant, void* stream);

/* Same attention for UNIFORM sequence lengths (every sequence q_len queries / k_len keys -- the DiT's image stream),
 * as ONE persistent wave of 2-CTA clusters; FLITE_TUNE_ATTN_SK_MODE picks the schedule: 1 = whole units round-robin
 * (what flite_b200.DiT uses: never splits a unit, bit-identical to flite_attention_varlen), 0 = stream-K work shares
 * (units x key-tiles cut into equal contiguous shares), 2 = hybrid (whole rounds, then shares of 1..2 units).  A unit
 * split between two clusters is merged through `workspace` (flite_attention_streamk_workspace_bytes() bytes, 16-byte
 * aligned, zero-filled ONCE by the caller when allocated; the kernel leaves it zeroed).                model.py:203-211 */
int64_t flite_attention_streamk_workspace_bytes(void);
int flite_attention_streamk(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                            int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out,
                            int64_t ldo, const int* cu_q, const int* cu_k, int B, int H, int q_len, int k_len,
                            float softmax_scale, void* workspace, int64_t workspace_bytes, void* stream);
/* ... with the fused Ulysses return path of flite_attention_varlen_p2p (rows stored into the token owner's buffer). */
int flite_attention_streamk_p2p(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                                int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0,
                                void* const* peer_out, int n_peers, int tokens_per_rank, int head0, int64_t ldo,
                                const int* cu_q, const int* cu_k, int B, int H, int q_len, int k_len,
                                float softmax_scale, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused compute + exchange over NVLink peer memory (Ulysses sequence parallelism) ----
 * flite_gemm_qkv_p2p: the QKV projection (+bias, RoPE, QK-norm) whose epilogue stores every head straight into the
 *   receive buffer of the rank that owns it: peer_recv[g][(sample*seq_len + sp_rank*tokens_per_sample + t), q|k|v][..]
 *   (row stride 3*sp_heads_per_rank*256).  Replaces GEMM -> all_to_all.
 * flite_attention_varlen_p2p: attention over the full sequence for this rank's heads whose epilogue stores each query
 *   row into the buffer of the rank that owns the token: peer_out[l / tokens_per_rank][(b*tokens_per_rank + l %
 *   tokens_per_rank), (head0 + h)*256 ..] (row stride ldo).  Replaces attention -> all_to_all -> permute.
 * peer_* are HOST arrays of device pointers valid in this process (own allocation + cudaIpcOpenMemHandle of peers). */
int flite_gemm_qkv_p2p(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const void* bias,
                       int tokens_per_sample, const void* rope_cos, const void* rope_sin, float eps, int sp_ranks,
                       int sp_heads_per_rank, int sp_rank, int seq_len, void* const* peer_recv, int variant,
                       void* stream);
int flite_attention_varlen_p2p(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                               int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0,
                               void* const* peer_out, int n_peers, int tokens_per_rank, int head0, int64_t ldo,
                               const int* cu_q, const int* cu_k, int B, int H, int max_q, float softmax_scale,
                               int variant, void* stream);

/* Symmetric peer allocations (cudaMalloc + CUDA IPC) and stream-ordered cross-GPU completion flags.
 *   signal: after all prior work of `stream`, write `value` into slot my_slot of each peer's flag array (uint32[8]);
 *   wait:   block `stream` until slots [0, n) of the local flag array are >= value (monotonic epochs).  The peers are
 *           paced by their hosts, so the wait tolerates FLITE_TUNE_P2P_TIMEOUT_S seconds (default 120) before it sets
 *           the sticky watchdog word; flite_poison_on_abort then fails the result closed.
 *   poison_on_abort: stream-ordered, no host sync: if the watchdog word of this process is set, overwrite the bf16
 *           buffer with NaNs (called on the output of every sequence-parallel forward). */
int flite_p2p_alloc(int64_t bytes, void** out);
int flite_p2p_free(void* p);
int flite_ipc_get_handle(const void* p, void* handle64);
int flite_ipc_open(const void* handle64, void** out);
int flite_ipc_close(void* p);
int flite_p2p_signal(void* const* peer_flags, int n, int my_slot, unsigned int value, void* stream);
int flite_p2p_wait(const void* my_flags, int n, unsigned int value, void* stream);
int flite_poison_on_abort(void* buf, int64_t numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLITE_B200_H */
