/*
 * mst.h — C-ABI of the B200-native motion-style-transfer sampler ("mst").
 *
 * This is the drop-in boundary for ONE hot path of
 * hlcdyy/diffusion-based-motion-style-transfer: the MDM-style denoising loop
 * (p_sample_loop / ddim_sample_loop + respacing + inpainting blend + CFG)
 * around the transformer-encoder denoiser.  The reference is pure
 * Python/PyTorch and has no FFI of its own; each entry point below cites the
 * Python function (file:line, relative to the reference root) whose device
 * work it replaces.  The Python host layer that mirrors the reference's class
 * API lives in `diffusion-based-motion-style-transfer_b200/` and binds these
 * symbols with ctypes.
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer
 *     owned by the caller (PyTorch allocates; the library never frees it);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream()
 *     .cuda_stream); all launches are asynchronous on it and are legal inside
 *     CUDA-graph capture (no allocation, no sync, no host reads);
 *   - every function returns 0 on success, non-zero on error;
 *     mst_last_error() returns a thread-local message for the last failure;
 *   - tensors are dense, row-major, fp32 unless stated.  The motion state is
 *     [B, F, 1, T] (reference layout, model/mdm_forstyledataset.py:317) which
 *     is addressed here as [B][F][T].
 *   - there is NO CPU fallback: a missing GPU / wrong arch is an error.
 */
#ifndef MST_H_
#define MST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MST_OK 0
#define MST_ERR_INVALID 1
#define MST_ERR_CUDA 2
#define MST_ERR_UNSUPPORTED 3

#define MST_MAX_LAYERS 32

/* precision of the denoiser forward */
#define MST_PREC_FP32 0 /* SIMT fp32 everywhere: parity mode (<=1e-4 rel)   */
#define MST_PREC_BF16 1 /* tcgen05 bf16 x bf16 -> fp32 TMEM accumulate      */

/* sampler kind for mst_update_step */
#define MST_SAMPLER_DDPM 0 /* p_sample,   inpainting_gaussian_diffusion.py:25  */
#define MST_SAMPLER_DDIM 1 /* ddim_sample, inpainting_gaussian_diffusion.py:125 */

/* noise source for mst_update_step */
#define MST_NOISE_NONE 0   /* no noise term (t==0 everywhere / eta==0)        */
#define MST_NOISE_TENSOR 1 /* read eps from `noise` (torch RNG or injected)   */
#define MST_NOISE_PHILOX 2 /* Philox4x32-10 + Box-Muller generated in-kernel  */

/* inpainting-mask storage */
#define MST_MASK_NONE 0
#define MST_MASK_FULL 1 /* [B,F,T] fp32, exactly what the reference reads     */
#define MST_MASK_FT 2   /* [F,T]  fp32, batch-invariant mask                  */
#define MST_MASK_F 3    /* [F]    fp32, batch- and time-invariant mask        */

const char* mst_version(void);
const char* mst_last_error(void);

/* Number of SMs / compute capability of the current device (diagnostics). */
int mst_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Number of kernels this library has launched in this process (all threads,
 * all streams).  Launches recorded into a CUDA graph are counted once, at
 * capture; a replay re-issues them without passing through the host library,
 * so callers multiply the captured count by their replays (bench.py does). */
uint64_t mst_launch_count(void);

/* Per-launch device timing of everything this thread launches through the
 * library on `stream` between begin and end: a CUDA event is recorded after
 * every kernel launch.  mst_profile_end synchronises, writes the duration of
 * launch i to ms[i] (i < cap) and the '\n'-separated kernel names to `names`,
 * and returns the number of launches in *n.  Not legal during graph capture.
 * This is the live source of bench.py's roofline numbers.                   */
int mst_profile_begin(void* stream);
int mst_profile_end(float* ms, char* names, int32_t cap, size_t names_cap, int32_t* n);
/* Deferred form for CUDA graphs: begin/end may bracket launches that are being CAPTURED (the events become
 * event-record nodes); call mst_profile_end(NULL, NULL, 0, 0, &n), replay the graph as often as wanted, then
 * mst_profile_collect reads the per-launch times of the LAST replay - kernel time plus the in-graph gap to
 * the next kernel, in the thermal / power state of a long run.                                              */
int mst_profile_collect(float* ms, char* names, int32_t cap, size_t names_cap, int32_t* n);

/* sizeof() of the argument structs as this library was compiled - lets a
 * foreign-language binding verify its own struct layout (no GPU needed).   */
int mst_abi_sizes(size_t* model_desc, size_t* weights, size_t* forward_args, size_t* update_args);

/* ------------------------------------------------------------------------ *
 * Denoiser description — model/mdm_forstyledataset.py:184-270 (MDM.__init__),
 * :495-567 (StyleDiffusion.__init__); utils/model_util.py:108-167.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_feats;   /* F = njoints*nfeats (181 stylexia, 263 humanml)       */
  int32_t d_model;   /* latent_dim (512)                                     */
  int32_t n_heads;   /* 4                                                    */
  int32_t d_ff;      /* 1024                                                 */
  int32_t n_layers;  /* 8                                                    */
  int32_t clip_dim;  /* 512                                                  */
  int32_t pe_len;    /* rows of the positional-encoding table (5000)         */
  int32_t precision; /* MST_PREC_*                                           */
} mst_model_desc;

/* One nn.TransformerEncoderLayer (post-norm, GELU) — fp32 device pointers in
 * the reference's own state_dict layout (mdm_forstyledataset.py:231-238).   */
typedef struct {
  const float* qkv_w; /* self_attn.in_proj_weight [3d, d] */
  const float* qkv_b; /* self_attn.in_proj_bias   [3d]    */
  const float* o_w;   /* self_attn.out_proj.weight [d, d] */
  const float* o_b;   /* self_attn.out_proj.bias   [d]    */
  const float* w1;    /* linear1.weight [ff, d]           */
  const float* b1;    /* linear1.bias   [ff]              */
  const float* w2;    /* linear2.weight [d, ff]           */
  const float* b2;    /* linear2.bias   [d]               */
  const float* ln1_g; /* norm1.weight [d]                 */
  const float* ln1_b; /* norm1.bias   [d]                 */
  const float* ln2_g; /* norm2.weight [d]                 */
  const float* ln2_b; /* norm2.bias   [d]                 */
} mst_layer_weights;

typedef struct {
  const float* in_w;  /* input_process.poseEmbedding.weight [d, F]  (:431) */
  const float* in_b;  /* input_process.poseEmbedding.bias   [d]            */
  const float* pe;    /* sequence_pos_encoder.pe [pe_len, d]        (:392) */
  const float* t_w1;  /* embed_timestep.time_embed.0.weight [d, d]  (:415) */
  const float* t_b1;
  const float* t_w2;  /* embed_timestep.time_embed.2.weight [d, d]         */
  const float* t_b2;
  const float* txt_w; /* embed_text.weight [d, clip_dim]            (:258) */
  const float* txt_b;
  const float* out_w; /* output_process.poseFinal.weight [F, d]     (:460) */
  const float* out_b;
  mst_layer_weights layers[MST_MAX_LAYERS];
} mst_weights;

typedef struct mst_engine_s* mst_engine_t;

int mst_engine_create(const mst_model_desc* desc, mst_engine_t* out);
int mst_engine_destroy(mst_engine_t e);

/* Bytes of device memory the engine needs for its packed copy of the weights
 * (bf16 tiles padded for TMA in MST_PREC_BF16; 0-copy aliases in FP32).     */
int mst_engine_packed_weight_bytes(mst_engine_t e, size_t* bytes);

/* (Re)pack weights.  Call after load_state_dict and after every optimizer
 * step.  `packed_dev` must stay alive while the engine is used.             */
int mst_engine_load_weights(mst_engine_t e, const mst_weights* w, void* packed_dev,
                            size_t packed_bytes, void* stream);

/* Workspace (activations) needed for one forward over n_seqs sequences of
 * n_frames frames (n_seqs = 2B under batched CFG).                          */
int mst_engine_workspace_bytes(mst_engine_t e, int n_seqs, int n_frames, size_t* bytes);

/* TimestepEmbedder.forward (mdm_forstyledataset.py:421): rows of
 * time_embed(pe[t]).  t_dev: int64[n] device.  out: [n, d] device.          */
int mst_time_embed(mst_engine_t e, const int64_t* t_dev, int n, float* out_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* embed_text (mdm_forstyledataset.py:327 / :615): feat [n, clip_dim] -> [n, d] */
int mst_text_embed(mst_engine_t e, const float* feat_dev, int n, float* out_dev, void* stream);

/* ------------------------------------------------------------------------ *
 * Denoiser forward — MDM.forward (:315-364) / StyleDiffusion.forward
 * (:602-625), optionally with the two passes of ClassifierFreeSampleModel
 * .forward (model/cfg_sampler.py:36-43) batched into one launch sequence.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t batch;     /* B                                                    */
  int32_t n_frames;  /* T  (S = T+1 tokens)                                  */
  int32_t cfg;       /* 0: one pass;  1: cond + uncond passes batched (2B)   */
  int32_t uncond;    /* cfg==0 only: force_mask (y['uncond']) for all rows   */
  const float* x;    /* [B, F, T] device                                     */
  /* token-0 ingredients */
  const float* temb; /* [n_rows, d] device: time_embed rows                  */
  const int32_t* temb_row_dev; /* device scalar: row of `temb` shared by all
                                  samples (graph-replayable); NULL => row b  */
  int32_t temb_row_offset;     /* added to *temb_row_dev (or to b)           */
  const float* text_emb;       /* [B, d] device: embed_text(clip(text)); may
                                  be NULL when uncond or cond_mode has no text*/
  /* outputs */
  float* out_cond;   /* [B, F, T] device (the only output when cfg==0)       */
  float* out_uncond; /* [B, F, T] device, cfg==1                             */
  void* workspace;   /* device, >= mst_engine_workspace_bytes                */
  size_t workspace_bytes;
  int32_t use_graph; /* mst_denoiser_forward_train only: 1 = the caller keeps
                        every pointer of this call stable across calls, so
                        the launch sequence may be captured into a CUDA
                        graph once (keyed by the argument values) and
                        replayed - ~150 launches become one              */
  float dropout_p;   /* mst_denoiser_forward_train only: p of the dropouts of
                        nn.TransformerEncoderLayer / PositionalEncoding (0.1
                        in the reference once model.train() was called); 0 =
                        identity (eval mode)                              */
  const uint64_t* dropout_seed; /* device, one Philox key PER SEQUENCE of this
                        call (read by the kernels, so a graph replay sees
                        new values); the backward must get the same keys  */
  int32_t tape_seqs;        /* mst_denoiser_forward_train only: the tape holds
                        this many sequences (0 = batch) and this call
                        records sequences [tape_seq_offset, +batch) of it:
                        forwards recorded one by one can be back-propagated
                        by ONE mst_denoiser_backward over the whole tape   */
  int32_t tape_seq_offset;
} mst_forward_args;

int mst_denoiser_forward(mst_engine_t e, const mst_forward_args* a, void* stream);

/* ------------------------------------------------------------------------ *
 * Fused per-step update — replaces, in ONE memory-bound kernel:
 *   cfg lerp          model/cfg_sampler.py:43
 *   inpainting blend  diffusion/gaussian_diffusion.py:341-349
 *   x0 clamp          :390-396
 *   posterior mean    :287-309 (coef1*x0 + coef2*x_t), FIXED_SMALL var :375-388
 *   masked noise      diffusion/inpainting_gaussian_diffusion.py:51-63
 *   (DDIM variant     :125-174)
 * All coefficient tables are fp32[N] device arrays produced on the host in
 * float64 and cast once (gaussian_diffusion.py:1605-1618 casts per call).
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t batch, n_feats, n_frames;
  int32_t sampler;          /* MST_SAMPLER_*                                 */
  int32_t clip_denoised;    /* clamp x0 to [-1,1]                            */
  /* model output(s) */
  const float* out_cond;    /* [B,F,T]                                       */
  const float* out_uncond;  /* [B,F,T] or NULL (no CFG)                      */
  const float* cfg_scale;   /* [B] or NULL                                   */
  /* state */
  const float* x_t;         /* [B,F,T]                                       */
  float* x_prev;            /* [B,F,T] out: "sample"                         */
  float* pred_xstart;       /* [B,F,T] out or NULL                           */
  /* inpainting */
  int32_t mask_kind;        /* MST_MASK_*                                    */
  const float* mask;        /* per mask_kind                                 */
  const float* x_inpaint;   /* [B,F,T] (y['inpainted_motion']) or NULL       */
  int32_t mask_noise;       /* 1: eps *= (1-mask) (Inpainting override);
                               0: base GaussianDiffusion.p_sample            */
  /* timestep: exactly one of the three sources */
  const int64_t* t_vec;     /* [B] device, per-sample t, or NULL             */
  int32_t* t_scalar_dev;    /* device scalar shared by the batch, or NULL    */
  int32_t t_imm;            /* host immediate, used when both are NULL       */
  int32_t advance_t;        /* 1: after the step, *t_scalar_dev -= 1 (last
                               block does it) so a captured graph can replay */
  int32_t* block_counter;   /* device int32 scratch (zeroed), for advance_t  */
  /* schedule tables, fp32[N] device */
  const float* coef1;       /* DDPM: posterior_mean_coef1 | DDIM: sqrt(abar_prev)          */
  const float* coef2;       /* DDPM: posterior_mean_coef2 | DDIM: sqrt(1-abar_prev-sig^2)  */
  const float* sigma;       /* DDPM: exp(.5*post_log_var_clipped) | DDIM: eta-sigma        */
  const float* recip;       /* DDIM: sqrt_recip_alphas_cumprod    (NULL for DDPM)          */
  const float* recipm1;     /* DDIM: sqrt_recipm1_alphas_cumprod  (NULL for DDPM)          */
  /* noise */
  int32_t noise_kind;       /* MST_NOISE_*                                   */
  const float* noise;       /* [B,F,T] when MST_NOISE_TENSOR                 */
  int32_t const_noise;      /* 1: every sample uses sample 0's noise (:571)  */
  uint64_t philox_seed;
  uint64_t philox_sample_offset; /* global index of sample 0 of this shard   */
} mst_update_args;

int mst_update_step(const mst_update_args* a, void* stream);

/* q_sample with the inpainting override (inpainting_gaussian_diffusion.py:6-23):
 * x_t = sqrt_ab[t]*x0 + sqrt_1m_ab[t]*(eps*(1-mask)).  `noise` is read-only
 * here; the in-place mutation the reference performs on its `noise` argument is
 * reproduced by the Python layer when it matters.  mask may be NULL.        */
int mst_q_sample(const float* x_start, const float* noise, int32_t mask_kind, const float* mask,
                 const int64_t* t_vec, int32_t t_imm, const float* sqrt_ab, const float* sqrt_1m_ab,
                 float* x_t, int32_t batch, int32_t n_feats, int32_t n_frames, void* stream);

/* out = out_u + scale[b]*(out_c - out_u)  (model/cfg_sampler.py:43)         */
int mst_cfg_combine(const float* out_cond, const float* out_uncond, const float* scale,
                    float* out, int32_t batch, int64_t per_sample, void* stream);

/* Fill `out` [B,F,T] with the N(0,1) stream MST_NOISE_PHILOX uses for (seed,
 * step key `t`): lets tests and the oracle check the in-kernel generator.   */
int mst_philox_normal(float* out, int32_t batch, int64_t per_sample, uint64_t seed,
                      uint64_t sample_offset, int32_t t, void* stream);

/* ------------------------------------------------------------------------ *
 * Training path - fp32, MST_PREC_FP32 engines.
 * The reference lets torch autograd differentiate StyleDiffusion.forward
 * (model/mdm_forstyledataset.py:602-625) inside
 * few_shot_style_finetune_losses (diffusion/gaussian_diffusion.py:1317-1399);
 * here the forward records an activation tape and the backward is explicit.
 * Dropout is the identity.
 * ------------------------------------------------------------------------ */
/* Gradient buffers of one encoder layer, same order and shapes as
 * mst_layer_weights.  Gradients are ACCUMULATED (+=) into non-NULL buffers
 * (torch .grad semantics); a NULL pointer skips that gradient.             */
typedef struct {
  float* qkv_w; float* qkv_b; float* o_w; float* o_b; float* w1; float* b1; float* w2; float* b2;
  float* ln1_g; float* ln1_b; float* ln2_g; float* ln2_b;
} mst_layer_grads;

/* Bytes of the activation tape / of the backward scratch for n_seqs sequences
 * of seq_len TOKENS (T+1 for the denoiser, T+2 for the MotionEncoder).      */
int mst_train_sizes(mst_engine_t e, int32_t n_seqs, int32_t seq_len, size_t* tape_bytes, size_t* scratch_bytes);

/* mst_denoiser_forward (cfg must be 0) that also fills `tape`.              */
int mst_denoiser_forward_train(mst_engine_t e, const mst_forward_args* a, void* tape, size_t tape_bytes, void* stream);

typedef struct {
  int32_t batch, n_frames;
  const float* d_out;       /* [B,F,T] gradient w.r.t. the model output      */
  float* d_x;               /* [B,F,T] gradient w.r.t. x, or NULL            */
  const mst_layer_grads* layer_grads; /* host array of n_layers entries      */
  void* tape;               /* filled by mst_denoiser_forward_train; the
                               backward reuses dead slots as scratch, so a
                               tape can be back-propagated ONCE              */
  size_t tape_bytes;
  void* scratch;
  size_t scratch_bytes;
  int32_t use_graph;        /* as in mst_forward_args: pointers (incl. the
                               gradient buffers) are stable across calls    */
  float dropout_p;          /* the forward's values: the masks are          */
  const uint64_t* dropout_seed; /* recomputed, not stored (one key / sequence) */
  int32_t tape_seqs;        /* as in mst_forward_args: back-propagates      */
  int32_t tape_seq_offset;  /* sequences [tape_seq_offset, +batch) of the tape */
} mst_backward_args;

int mst_denoiser_backward(mst_engine_t e, const mst_backward_args* a, void* stream);

/* sizeof() of the training structs as compiled (binding self-check, no GPU). */
int mst_abi_sizes_train(size_t* layer_grads, size_t* backward_args);

/* MotionEncoder.forward (model/mdm_forstyledataset.py:89-124): tokens
 * [muQuery, sigmaQuery, InputProcess(x)] + pe, encoder stack of THIS engine
 * with a key-padding mask (key_valid [B, T+2] bytes, 1 = attend; NULL = all),
 * mu_out [B, d] = output token 0.  The engine's in_w/in_b/pe must be the
 * frozen mdm_model's.  The backward yields only d_x: every MotionEncoder
 * parameter is frozen in the finetune loss.                                 */
int mst_motion_encoder_forward(mst_engine_t e, const float* x, const uint8_t* key_valid, const float* mu_query,
                               const float* sigma_query, int32_t batch, int32_t n_frames, float* mu_out, void* tape,
                               size_t tape_bytes, float dropout_p, const uint64_t* dropout_seed, int32_t use_graph,
                               void* stream);
int mst_motion_encoder_backward(mst_engine_t e, const float* d_mu, int32_t batch, int32_t n_frames, float* d_x,
                                void* tape, size_t tape_bytes, void* scratch, size_t scratch_bytes, float dropout_p,
                                const uint64_t* dropout_seed, int32_t use_graph, void* stream);

/* Test hook: the multiplier (0 or 1/(1-p)) the training kernels apply to element i of dropout site `site`
 * (0 = token sequence after the positional encoding; 8*(layer+1) + {1: attention probabilities [seq][head][q][k],
 * 2: out-proj output, 3: GELU output, 4: linear2 output}, token-major [seq][token][column]).  n % 4 == 0.       */
int mst_test_dropout_scale(float* out, int64_t n, float p, const uint64_t* seed_dev, int32_t site, void* stream);

/* masked_l2 (diffusion/gaussian_diffusion.py:223-235) over `rows` rows of
 * b [rows,F,T]; a ([a_rows,F,T]) and mask ([mask_rows,T]) rows are taken
 * modulo their row counts (the reference expands them over the stacked
 * steps).  Forward: loss [rows] (grad_* NULL).  Backward: grad_b [rows,F,T]
 * = grad_loss[row] * d loss / d b (loss NULL).                              */
int mst_masked_l2(const float* a, const float* b, const float* mask, float* loss, const float* grad_loss,
                  float* grad_b, int32_t rows, int32_t a_rows, int32_t mask_rows, int32_t n_feats, int32_t n_frames,
                  void* stream);

/* Gradient of mst_update_step w.r.t. the (non-CFG) model output:
 * d_out = (d_pred_xstart + k_table[t] * d_sample) * (1 - mask), zeroed where
 * the x0 clamp was active.  k_table: d sample / d x0 per timestep (DDPM:
 * posterior_mean_coef1; DDIM: sqrt(abar_prev) - sqrt(1-abar_prev-sig^2) /
 * sqrt_recipm1).  Either gradient may be NULL.
 * (inpainting_gaussian_diffusion.py:66-123, :176-239)                       */
int mst_update_step_backward(const float* d_pred_xstart, const float* d_sample, const float* k_table,
                             const int64_t* t_vec, int32_t mask_kind, const float* mask, const float* pred_xstart,
                             int32_t clip_denoised, float* d_out, int32_t batch, int32_t n_feats, int32_t n_frames,
                             void* stream);

/* torch.optim.AdamW step over flat fp32 arenas (train/training_loop.py:97-99);
 * grads are multiplied by grad_scale first (1/world_size after a SUM
 * all-reduce).  `step` is the 1-based step count.                           */
int mst_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   void* stream);

/* out2[0] = sum x^2, out2[1] = sum y^2 (y may be NULL) as float64 on the
 * device: MixedPrecisionTrainer._compute_norms (diffusion/fp16_util.py:
 * 215-223) without its 192 host syncs.                                      */
int mst_sumsq2(const float* x, const float* y, int64_t n, double* out2, void* stream);

/* Post-sampling decode (scope row N2), fused: inv_transform x*std+mean
 * (data_loaders/humanml/data/dataset.py:478-479; mean/std [F] device, both
 * NULL = already de-normalised) + recover_from_ric
 * (data_loaders/humanml/scripts/motion_process.py:389-411, :444-461).
 * x: [B,F,T] in the sampler's layout; joints: [B,T,J,3] out.               */
int mst_recover_from_ric(const float* x, const float* mean, const float* stdv, float* joints, int32_t batch,
                         int32_t n_feats, int32_t n_frames, int32_t joints_num, void* stream);

/* ------------------------------------------------------------------------ *
 * CLIP text tower (scope row N1) - what `self.clip_model.encode_text(texts)`
 * computes in MDM.encode_text (model/mdm_forstyledataset.py:298-313, loaded
 * by load_and_freeze_clip :275-286).  The arithmetic is third-party:
 * openai/CLIP @ a9b1bf59 (requirements.txt:26), clip/model.py
 * CLIP.encode_text: token_embedding + positional_embedding -> `layers`
 * pre-norm residual blocks (x += MHA(ln_1 x, causal mask); x += c_proj(
 * QuickGELU(c_fc(ln_2 x)))) -> ln_final -> the row of the highest token id
 * (end-of-text) times text_projection.  Tokenising stays with the caller
 * (clip.tokenize needs the BPE vocabulary file, which is not redistributed
 * here).  fp32 weights in the CLIP state_dict layout.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t vocab;     /* 49408                                               */
  int32_t ctx;       /* context length 77 (<= 80)                           */
  int32_t width;     /* transformer_width 512                               */
  int32_t n_heads;   /* 8  (head_dim must be 64)                            */
  int32_t n_layers;  /* 12                                                  */
  int32_t d_ff;      /* 4 * width                                           */
  int32_t d_out;     /* embed_dim 512 (= MDM clip_dim)                      */
  int32_t precision; /* MST_PREC_*                                          */
} mst_clip_text_desc;

typedef struct {
  const float* ln1_g;  /* transformer.resblocks.{l}.ln_1.weight [w]          */
  const float* ln1_b;
  const float* qkv_w;  /* .attn.in_proj_weight [3w, w]                       */
  const float* qkv_b;
  const float* o_w;    /* .attn.out_proj.weight [w, w]                       */
  const float* o_b;
  const float* ln2_g;  /* .ln_2.weight [w]                                   */
  const float* ln2_b;
  const float* fc_w;   /* .mlp.c_fc.weight [ff, w]                           */
  const float* fc_b;
  const float* proj_w; /* .mlp.c_proj.weight [w, ff]                         */
  const float* proj_b;
} mst_clip_text_layer;

typedef struct {
  const float* token_embedding;      /* token_embedding.weight [vocab, w]    */
  const float* positional_embedding; /* [ctx, w]                             */
  const float* lnf_g;                /* ln_final.weight [w]                  */
  const float* lnf_b;
  const float* text_projection;      /* [w, d_out]                           */
  mst_clip_text_layer layers[MST_MAX_LAYERS];
} mst_clip_text_weights;

typedef struct mst_clip_text_s* mst_clip_text_t;

int mst_clip_text_create(const mst_clip_text_desc* desc, mst_clip_text_t* out);
int mst_clip_text_destroy(mst_clip_text_t h);
int mst_abi_sizes_clip_text(size_t* desc, size_t* weights);
/* device bytes for the bf16 weight packs (0 in MST_PREC_FP32) and for the
 * activations of `batch` token rows                                         */
int mst_clip_text_packed_weight_bytes(mst_clip_text_t h, size_t* bytes);
int mst_clip_text_workspace_bytes(mst_clip_text_t h, int32_t batch, size_t* bytes);
int mst_clip_text_load_weights(mst_clip_text_t h, const mst_clip_text_weights* w, void* packed_dev,
                               size_t packed_bytes, void* stream);
/* tokens int32 [batch, ctx] device (ids outside [0, vocab) are an error the
 * host mirror checks; the kernel clamps) -> features fp32 [batch, d_out].   */
int mst_clip_text_encode(mst_clip_text_t h, const int32_t* tokens, int32_t batch, float* features,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ *
 * Kernel-level test hooks (used by tests/ and bench.py roofline legs only).
 * ------------------------------------------------------------------------ */
/* C[M,N] (fp32) = A[M,K](bf16) * W[N,K](bf16)^T + bias[N], tcgen05 path.    */
int mst_test_gemm_bf16(const void* a_bf16, const void* w_bf16, const float* bias, float* c,
                       int32_t m, int32_t n, int32_t k, void* stream);
/* bf16-output epilogues of the same GEMM: epi 0 = bias, 1 = bias+GELU,
 * 2 = LayerNorm(acc + bias + residual)*g + b (n must be 512).  out: bf16 [m,n]. */
int mst_test_gemm_epi_bf16(int32_t epi, const void* a_bf16, const void* w_bf16, const float* bias,
                           const void* residual_bf16, const float* ln_g, const float* ln_b, void* out_bf16,
                           int32_t m, int32_t n, int32_t k, void* stream);
/* Developer hook: when set to a device buffer of >= 6*1024 int64, cluster 0 of the pair GEMM records clock64()
 * timelines of its producer / MMA / epilogue roles there (tools/gemm_timeline.py).  NULL switches it off.  */
int mst_test_set_gemm_debug(void* dev_buf_int64);
/* softmax(QK^T/sqrt(dh))V for qkv [n_seqs*S, 3d] bf16 -> out [n_seqs*S, d] bf16 */
int mst_test_attention_bf16(mst_engine_t e, const void* qkv_bf16, void* out_bf16, int32_t n_seqs,
                            int32_t seq_len, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_H_ */
