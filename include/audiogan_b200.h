/*
 * audiogan_b200 -- C ABI of the B200 (sm_100a) kernels behind the audiogan GAN training step.
 *
 * The reference (BarclayII/audiogan) has no FFI of its own: its hot path is Python calling
 * torch.nn modules (audiogan.py).  Each entry point below replaces the PyTorch call site(s)
 * cited next to it (file:line in /root/reference); INTEGRATION.md shows the ctypes stubs a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - the library never allocates, frees or retains caller memory (workspaces are passed in);
 *   - every call launches on the cudaStream_t passed as `stream` (void* here) and returns
 *     0 on success or a negative AG_E* code; nothing throws or exits across the ABI;
 *     ag_last_error_string() describes the last failure on the calling thread;
 *   - fp32 unless a parameter says otherwise; sizes are int64_t.
 */
#ifndef AUDIOGAN_B200_H
#define AUDIOGAN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AG_OK 0
#define AG_EINVAL (-1)   /* bad argument / unsupported shape */
#define AG_ECUDA (-2)    /* CUDA runtime error (see ag_last_error_string) */
#define AG_ENOTSUP (-3)  /* device is not sm_100 / feature not available */

int ag_version(void);
const char* ag_last_error_string(void);
/* cudaStreamSynchronize + cudaGetLastError: surfaces asynchronous faults. */
int ag_sync_check(void* stream);
/* SM count, max opt-in shared memory per block, compute capability major*10+minor. */
int ag_device_info(int* sm_count, int* smem_optin, int* cc);

/* ------------------------------------------------------------------------------------------
 * View-GEMM: one engine for Linear / Conv1d / ConvTranspose1d forward, data-gradient and
 * weight-gradient.  Replaces F.linear (audiogan.py:260, :409-410, :509-511), NN.Conv1d
 * (:272, :406, :490) and NN.ConvTranspose1d (:275) plus their autograd backward, with the bias,
 * LeakyReLU (:261, :277, :532), skip (:264, :282) and length mask (:534) fused in the epilogue.
 *
 * Operands are *views*: a row index m splits into (batch = m / rpb, t = m % rpb) and a column
 * index into (outer = k / kin, inner = k % kin) so that an im2col matrix of a channel-last
 * activation is addressed in place (no im2col copy):
 *   A(m,k)  at  A + batch*a_bs + t*a_rs + outer*a_k1s + inner
 *   B(n,k)  at  B + n*ldb + k                                  (packed weights, K contiguous)
 *   C(m,n)  at  C + batch*c_bs + t*c_rs + (n / c_nin)*c_n1s + n % c_nin
 * ------------------------------------------------------------------------------------------ */
typedef struct ag_gemm_desc {
  int64_t M, N, K;
  const void* A; int64_t a_rpb, a_bs, a_rs, a_kin, a_k1s;
  const void* B; int64_t ldb;
  void* C;       int64_t c_rpb, c_bs, c_rs, c_nin, c_n1s;
  /* epilogue, applied in this order (null pointer / zero flag = skipped) */
  float alpha;                  /* acc *= alpha (0 means 1) */
  const float* bias;  int64_t bias_mod;     /* += bias[n % bias_mod] */
  const float* rowbias; int64_t rowbias_ld; /* += rowbias[batch*rowbias_ld + n] */
  const void* skip;             /* += skip[C-addressing]   (skip == C gives C += ...) */
  int32_t act;                  /* 1: LeakyReLU(slope) */
  const void* dact;             /* *= (dact[C-addressing] > 0 ? 1 : slope)  (LeakyReLU') */
  float slope;
  const int32_t* mask_len;      /* pos = t*mask_tmul + (n/c_nin)*mask_n1mul + mask_toff; zero unless 0 <= pos < mask_len[batch] */
  int64_t mask_tmul, mask_n1mul, mask_toff;
  int32_t a_dtype, b_dtype, c_dtype, aux_dtype;   /* 0 = fp32, 1 = bf16 (skip/dact use aux_dtype) */
  int32_t a_layout;   /* 0: B / dW columns in A's column order.  1 (tensor-core path, bf16, channel-prefix views with
                         a_kin % 8 == 0): columns ordered (channel group g of 8, tap j padded to a multiple of 8, channel c),
                         column = ((g*KT + j/8)*8 + j%8)*8 + c with KT = ceil(taps/8), taps = K/a_kin; ldb / ldw >=
                         (a_kin/8)*KT*64 (+1 with ones_col, the bias column comes last); padded taps hold zeros.
                         2 (same conditions): columns ordered (tap j, channel group g of 64, channel c), column =
                         (j*G + g)*64 + c with G = ceil(a_kin/64); ldb / ldw >= taps*G*64 (+1 with ones_col); the columns of
                         channels >= a_kin hold zeros (B) / receive zeros (dW).  128-byte TMA requests: the layout for prefixes
                         of 32 channels and more. */
} ag_gemm_desc;

/* C = epilogue(A . B^T).  fp32 FFMA path ("fp32 mode", <=1e-5 parity). */
int ag_gemm_nt_f32(const ag_gemm_desc* d, void* stream);
/* Weight gradient: C[n, k] += sum_m Y(m,n) * A(m,k), C plain [N, ldc] fp32, accumulated with
 * atomics (zero it first).  Y uses the C-addressing fields of the descriptor (d->C = Y, read
 * only), the result goes to `dw`.  If `ones_col` != 0, column K of dw receives sum_m Y(m,n)
 * (the bias gradient). */
int ag_gemm_tn_f32(const ag_gemm_desc* d, float* dw, int64_t ldw, int32_t ones_col, void* stream);

/* bf16 tcgen05 / TMA path ("bf16 mode", <=2e-2 parity): same descriptor, a_dtype = b_dtype = 1,
 * fp32 accumulation in TMEM.  Requires 16-byte aligned row strides. */
int ag_gemm_nt_tc(const ag_gemm_desc* d, void* stream);
int ag_gemm_tn_tc(const ag_gemm_desc* d, float* dw, int64_t ldw, int32_t ones_col, void* stream);
/* Profiling aid for ag_gemm_nt_tc (tools/gemm_phases.py): enable (on != 0, also resets) / read 16 per-phase cycle
 * totals summed over CTAs: [0] CTAs, [1] producer wait-for-loads, [2] wait-slot-free, [3] store+fence+arrive,
 * [4] main loop, [5] epilogue, [6] k-blocks, [7] MMA thread wait-full. */
int ag_gemm_dbg_enable(int on);
int ag_gemm_dbg_read(unsigned long long* out16);

/* ------------------------------------------------------------------------------------------
 * Persistent recurrent kernels (cooperative launch, one CTA per slice of hidden units, its
 * weight slice resident in shared memory for the whole sequence when it fits, one grid barrier
 * per dependent phase).  Replaces NN.LSTMCell + proj + stopper + stop sampling per frame
 * (audiogan.py:437-460) and NN.LSTM(bidirectional) under dynamic_rnn (:214-229, :498-503, :543),
 * and their backward (BPTT).  Gate order i,f,g,o (torch).  Layouts (fp32); Tcap is the allocated
 * number of steps, T <= Tcap the number to run:
 *   pre    [B, Tcap, ndir*4H]   input projections + both biases (hoisted GEMM)
 *   w1     [ndir, 4H, H+F]      rows in torch gate order; columns [whh | wx]
 *   hbuf   [B, Tcap+2, ndir*H]  row t+1 holds h_t; rows 0 and T+1 must be zero on entry
 *   gates  [B, Tcap, ndir*4H]   post-activation i,f,g,o (saved for backward; may be NULL)
 *   cbuf   [B, Tcap, ndir*H]    cell state c_t (saved for backward; may be NULL)
 *   len    [B] int32 or NULL (all T): steps t >= len[b] emit h = 0; direction 1 runs
 *          t = T-1 .. 0 and therefore starts at each sample's own last frame.
 * Feedback variant (generator, ndir == 1, F > 0): the gates also get wx . x_{t-1} where
 *   x_t = tanh(wp . h_t + bp), logit_t = ws . h_t + bs;  w2 = [wp ; ws] is [F+1, H], b2 [F+1].
 *   xbuf [B, Tcap+1, F] (row t+1 holds x_t, row 0 must be zero) and sbuf [B, Tcap] (logits).
 *   Stop sampling: stop[b,t] = u[b,t] < sigmoid(logit) (u NULL: never), glen[b] = frames until
 *   the first stop inclusive, *t_end = number of steps run (the loop ends early once every
 *   sample has stopped, audiogan.py:458-460).
 * Backward: dgates [B, Tcap, ndir*4H] = grad wrt the pre-activation gates (= grad wrt `pre`),
 *   dpx [B, Tcap, FP] = grad wrt the proj pre-activation, column F = grad wrt the stop logit,
 *   FP = F+1 rounded up to a multiple of 8 (pad columns zero).  Weight gradients are batched
 *   GEMMs over these two buffers (ag_gemm_tn_*), outside the recurrence.
 *   w1t [ndir, H, 4H+FP] rows j: [whh[:, j] | wp[:, j], ws[j], 0-pad];  wxt [F, 4H] = wx^T.
 * H % 4 == 0, F % 4 == 0.  `barrier`: >= 8 uint32 of device memory (zeroed by the call).
 * ------------------------------------------------------------------------------------------ */
typedef struct ag_lstm_desc {
  int32_t B, T, Tcap, H, ndir, F;
  const float* pre; const float* w1; const float* w2; const float* b2;
  float* hbuf; float* gates; float* cbuf;
  const int32_t* len;
  float* xbuf; float* sbuf;
  const float* u; int32_t* stop; int32_t* glen; int32_t* t_end;
  /* backward only */
  const float* dh_ext;   /* [B, Tcap, ndir*H] grad wrt emitted h (NULL = 0); batch stride dh_ext_bs if != 0 */
  const float* dx_ext;   /* [B, Tcap, F] grad wrt emitted frames (NULL = 0) */
  const float* ds_ext;   /* [B, Tcap] grad wrt stop logits (NULL = 0) */
  float* dgates; float* dpx;
  const float* w1t; const float* wxt;
  unsigned int* barrier;
  int64_t dh_ext_bs;
  /* bf16 mode (prec = 1): the recurrent products run on tensor cores (bf16 operands, fp32 accumulate, fp32 state).
   * The kernels keep bf16 shadow copies of the per-step operands, same shapes as their fp32 twins:
   * hbuf16 / xbuf16 (forward writes, rows 0 / T+1 zero on entry), dgates16 / dpx16 (backward writes). */
  int32_t prec, flags;       /* flags bit 0 (AG_LSTM_GRID_ONLY): use the grid-barrier kernels only (no cluster / TMEM-resident
                              * kernels); bit 1 (AG_LSTM_ALLOW_TMEM): allow the TMEM-resident generator kernel (needs ll_ws) */
  void* hbuf16; void* xbuf16; void* dgates16; void* dpx16;
  long long* dbg;        /* optional [gridDim][8] cycle counters per CTA: gemm, cell, barrier, phase2/A, total (profiling aid) */
  /* Workspace of the TMEM-resident generator kernel (bf16 mode, F > 0): the per-step h_t / x_t exchange between the CTAs
   * of a batch group runs over it ("LL" words: payload + step tag).  >= 16 * B_pad * (H/2 + F) + 256 bytes with
   * B_pad = B rounded up to 32; NULL -> the grid-barrier kernels are used.  The call zeroes what it needs. */
  void* ll_ws; int64_t ll_ws_bytes;
} ag_lstm_desc;

#define AG_LSTM_GRID_ONLY 1
#define AG_LSTM_ALLOW_TMEM 2
#define AG_LSTM_BF16_H_ONLY 8         /* forward on the cluster / TMEM-resident kernels: h_t goes to hbuf16 only (fp32 `hbuf` untouched) */
#define AG_LSTM_BF16_DGATES_ONLY 4   /* BPTT on the cluster / TMEM-resident kernels: write the gate gradients to dgates16 only (the
                                        fp32 `dgates` is then left untouched; bf16 mode's GEMMs read the bf16 copy) */
int ag_lstm_fwd(const ag_lstm_desc* d, void* stream);
int ag_lstm_bwd(const ag_lstm_desc* d, void* stream);
/* Bytes of `ll_ws` the TMEM-resident generator kernels need for this descriptor's (B, H, F) -- forward (bwd == 0) or BPTT
 * (bwd != 0); 0 when the shape runs on a path that needs no workspace.  Only B, H, F, ndir and prec are read. */
int64_t ag_lstm_workspace_bytes(const ag_lstm_desc* d, int32_t bwd);
/* Largest batch ONE launch of the TMEM-resident generator kernels takes for this descriptor's (H, F) on this device
 * (whole batch groups that are co-resident: 128 samples forward / 64 BPTT for the default net on 148 SMs); 0 when the shape
 * does not run there.  Samples are independent, so a caller with a larger batch runs the recurrence in chunks of this many
 * samples over row slices of the batch-major buffers instead of dropping to the grid-barrier kernels. */
int32_t ag_lstm_batch_cap(const ag_lstm_desc* d, int32_t bwd);
/* Which kernel family the last ag_lstm_fwd / ag_lstm_bwd call ON THIS THREAD ran on, and -- when a bf16-mode call fell off
 * the fast (cluster / TMEM-resident) kernels -- why: e.g. "cluster", "tmem", "grid-bf16 (tmem declined: 8 groups x 32
 * slices > 148 SMs)".  With AUDIOGAN_VERBOSE=1 in the environment every distinct decline is also printed to stderr once. */
const char* ag_lstm_last_path(void);
/* Non-feedback sequences in bf16 mode (prec >= 1, F == 0, H in {128, 256, 512}: the discriminator's BiLSTM) run on
 * cluster-resident kernels: one thread-block cluster of H/32 CTAs keeps a full copy of one direction's recurrent
 * weights in distributed shared memory and carries a slice of <= 32 samples through all T steps with tcgen05 MMAs,
 * exchanging h_t (forward) / partial dh sums (backward) through DSMEM -- no grid barrier.  Same buffers and layouts
 * as above.  Returns how many such clusters can be co-resident on this device (< ndir: the calls above fall back to
 * the grid-barrier kernels), or -1 on error. */
int ag_lstm_cluster_max_active(int H, int bwd);

/* ------------------------------------------------------------------------------------------
 * Weight-norm (audiogan.py:77-80 -> torch.nn.utils.weight_norm dim 0), multi-tensor.
 * One table entry per parameter tensor; kind 0: w = g*v/||v||_row, kind 1: w = v (plain
 * parameter, e.g. Discriminator.rnn).  The table and row_start live in DEVICE memory.
 * ------------------------------------------------------------------------------------------ */
typedef struct ag_wn_entry {
  const float* v; const float* g;   /* g unused for kind 1 */
  float* w;                         /* forward out [rows*cols] */
  float* norm;                      /* [rows] forward out / backward in */
  const float* dw; float* dv; float* dg;   /* backward */
  int32_t rows, cols, kind, reserved;
} ag_wn_entry;
int ag_wn_fwd_multi(const ag_wn_entry* table, const int32_t* row_start, int32_t ntensors,
                    int32_t total_rows, void* stream);
int ag_wn_bwd_multi(const ag_wn_entry* table, const int32_t* row_start, int32_t ntensors,
                    int32_t total_rows, void* stream);

/* dst[i] = idx[i] >= 0 ? src[idx[i]] : 0 ; dst_dtype 0 fp32 / 1 bf16.  Packs canonical weights
 * into GEMM operand layouts and un-packs weight gradients (index maps built once on the host). */
int ag_gather(void* dst, const float* src, const int32_t* idx, int64_t n, int32_t dst_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Waveform framing + noise (audiogan.py:462-464, :724-725, :750-751, :842-843):
 * dst[b, pad_l + i] = src[b*src_ld + i] + noise_scale*noise[b*L + i] (noise may be NULL) for i < L,
 * pads [0,pad_l) and [pad_l+L, dst_ld) zeroed.  dst_dtype 0/1.
 * ------------------------------------------------------------------------------------------ */
int ag_frame_noise(void* dst, int64_t dst_ld, int64_t pad_l, const float* src, int64_t src_ld,
                   const float* noise, float noise_scale, int64_t B, int64_t L, int32_t dst_dtype, void* stream);

/* Masked BCE-with-logits per sample (audiogan.py:187-197): loss[b] = sum_t w*(x - x*tgt + max(-x,0)
 * + log(exp(-max) + exp(-x-max))).  weight may be NULL (=1).  Backward: dx = gout[b]*w*(sigmoid(x)-tgt). */
int ag_bce_fwd(const float* x, const float* tgt, const float* w, float* loss, int64_t B, int64_t T, void* stream);
int ag_bce_bwd(const float* x, const float* tgt, const float* w, const float* gout, float* dx,
               int64_t B, int64_t T, void* stream);
/* Fused training-loop form (audiogan.py:739-740, :766, :780, :864, :897): per-sample masked BCE
 * against a constant target, / len[b], mean over B accumulated into *loss_mean (zero it first);
 * dlogits = (sigmoid(x)-target)/(len[b]*B) inside the mask, 0 outside.  Also counts
 * correct = sum(mask * (sign*x > 0)) and num = sum(mask) into stats[0..1] (:741-742, :781-782). */
int ag_bce_const_fused(const float* x, int64_t ld, const int32_t* len, float target, float sign,
                       float* loss_mean, float* loss_ps, float* dlogits, float* stats,
                       int64_t B, int64_t T, void* stream);

/* REINFORCE gradient of the generator's stop head (audiogan.py:873-908: reward = -loss, EMA(0.5) baseline,
 * fake_stop.reinforce(reward[:, i]), stopper-only backward).  reward_b = -loss_ps[b]; *baseline_out = mean_b(reward) when
 * baseline_in is NULL, else 0.5 * *baseline_in + 0.5 * mean_b(reward) (device scalars, must not alias);
 * out[b*out_ld + t] = -(reward_b - *baseline_out) * (stop[b,t] - sigmoid(s[b,t])) for t < glen[b] (frames, stop inclusive),
 * 0 for glen[b] <= t < T: the gradient of -sum (reward - baseline) * log p(stop_bt) with respect to the stop logits. */
int ag_reinforce_dlogit(const float* s, int64_t s_ld, const int32_t* stop, int64_t stop_ld, const float* loss_ps,
                        const int32_t* glen, const float* baseline_in, float* baseline_out, float* out, int64_t out_ld,
                        int64_t B, int64_t T, void* stream);

/* calc_dists time moments (audiogan.py:341-348: the per-(sample, channel) statistics over time of a discriminator conv
 * activation that the feature-matching penalty :848-855 compares between real and generated batches).
 * h: channel-last activation, element (b, t, c) at h[b*h_bs + t*h_rs + c] (dtype 0 fp32 / 1 bf16), rows t >= len[b] hold zeros.
 * fwd writes S1[b*C+c] = sum_t h and Q[q*B*C + b*C + c] = sum_{t<len} (h - S1/len)^(2+q), q = 0..2 (two streaming passes; the
 *   caller finishes m = S1/l, s = sqrt(Q2)/l, f = Q4^(1/4)/l on the (B, C) arrays).
 * bwd writes dh[b, t, c] (packed [B, T, C], dh_dtype) from the gradients gm, gs, gf (B, C) of m, s, f. */
int ag_time_moments_fwd(const void* h, int32_t dtype, int64_t h_bs, int64_t h_rs, const int32_t* len, int64_t B, int64_t T,
                        int64_t C, float* S1, float* Q, void* stream);
int ag_time_moments_bwd(const void* h, int32_t dtype, int64_t h_bs, int64_t h_rs, const int32_t* len, int64_t B, int64_t T,
                        int64_t C, const float* S1, const float* Q, const float* gm, const float* gs, const float* gf,
                        void* dh, int32_t dh_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Activation-gradient assembly (backward of LeakyReLU + length mask + dense skip,
 * audiogan.py:261-264, :277-283, :532-534), strided element-wise:
 *   v[b,t,c] = (g1[b,t,c] + g2[b,t,c]) * (act[b,t,c] > 0 ? 1 : slope) * (t < len[b])
 *   out[b, pad_l + t, c] = v, the pad_l rows before and pad_r rows after each sequence zeroed
 *   (out is a packed channel-last buffer [B, pad_l + T + pad_r, C] ready to be a GEMM operand);
 *   acc[b,t,c] += v when acc != NULL (the dense-net skip path);
 *   colsum[c] += sum_{b,t} v[b,t,c] when colsum != NULL (the bias gradient of the layer that produced the activation, fused:
 *   needs unit channel strides, 16-byte aligned rows and C = 4 * 2^k).
 * g1/g2/act/len may be NULL (0 / 1 / all).
 * ------------------------------------------------------------------------------------------ */
typedef struct ag_ew_desc {
  int64_t B, T, C;
  const void* g1; int64_t g1_bs, g1_rs, g1_cs;
  const void* g2; int64_t g2_bs, g2_rs, g2_cs;
  const void* act; int64_t a_bs, a_rs;
  float slope; int32_t reserved;
  const int32_t* len;
  void* out; int64_t pad_l, pad_r;
  void* acc; int64_t acc_bs, acc_rs;
  int32_t g1_dtype, g2_dtype, act_dtype, acc_dtype, out_dtype, reserved2;   /* 0 = fp32, 1 = bf16 (bf16 mode stores the conv
                                                                               stacks' activations and gradients as bf16) */
  float* colsum;
} ag_ew_desc;
int ag_ew_grad(const ag_ew_desc* d, void* stream);
/* out[c] += sum_{b,t} in[b*bs + t*rs + c]  (bias gradients); out must be initialised.  in: dtype 0 fp32 / 1 bf16. */
int ag_colsum(const void* in, int32_t dtype, int64_t bs, int64_t rs, int64_t B, int64_t T, int64_t C, float* out, void* stream);
/* Rank-1 data gradient with LeakyReLU' (backward of the classifier's 512 -> 1 layer, audiogan.py:508-512):
 * out[m, n] = g[m] * w[n] * (act[m, n] > 0 ? 1 : slope); act / out packed [M, N], dtype 0 fp32 / 1 bf16, N % 4 == 0. */
int ag_outer_dact(const float* g, const float* w, const void* act, int32_t act_dtype, void* out, int32_t out_dtype, int64_t M,
                  int64_t N, float slope, void* stream);

/* ------------------------------------------------------------------------------------------
 * The discriminator's first layer, Conv1d(1 -> C, k <= 8, stride s) + bias + LeakyReLU + length mask on the raw waveform
 * (audiogan.py:527-536 with C_in = 1), and its weight / bias gradient.  HBM-bound direct kernels (K = k is too small
 * for tensor cores).  x: zero-padded waveform rows (tap j of output t at x[b*x_ld + s*t + j]); out / dy: channel-last
 * [B, rows, C] with batch stride out_bs / dy_bs (pointers at row t = 0); w [C, k]; dw [C, k + 1] (column k = bias
 * gradient), accumulated with atomics (zero it first).  C % 4 == 0.  out / dy: dtype 0 fp32 / 1 bf16.
 * ------------------------------------------------------------------------------------------ */
int ag_conv1in_fwd(const float* x, int64_t x_ld, const float* w, const float* bias, void* out, int32_t out_dtype, int64_t out_bs,
                   int32_t k, int32_t s, int64_t C, int64_t B, int64_t T, const int32_t* len, float slope, void* stream);
int ag_conv1in_wgrad(const void* dy, int32_t dy_dtype, int64_t dy_bs, const float* x, int64_t x_ld, float* dw, int32_t k, int32_t s, int64_t C,
                     int64_t B, int64_t T, void* stream);
/* Data gradient of the same layer = gradient of the raw waveform (audiogan.py:769 x_grad_norm, :145 / :131 FGSM, and the
 * generator update): dx[b*dx_ld + u] = sum_{t,j: s*t + j == u} sum_c dy[b,t,c] * w[c*k + j] for p <= u < p + Tin (padded
 * coordinates, p = left zero pad of the forward input), 0 elsewhere; every one of the dx_ld entries is written. */
int ag_conv1in_dgrad(const void* dy, int32_t dy_dtype, int64_t dy_bs, const float* w, float* dx, int64_t dx_ld, int32_t k, int32_t s,
                     int32_t p, int64_t C, int64_t B, int64_t T, int64_t Tin, void* stream);
/* Weight + bias gradient of a Linear(K -> 1) (the classifier's last layer, audiogan.py:508-512 backward):
 * out[k] += sum_m g[m] * X[m*K + k], out[K] += sum_m g[m]; X packed [M, K], dtype 0 fp32 / 1 bf16, K % 4 == 0, K <= 1024. */
int ag_wcolsum(const float* g, const void* X, int32_t x_dtype, int64_t M, int64_t K, float* out, void* stream);
/* Forward of the same Linear(K -> 1) (audiogan.py:508-512, the logits of :549): out[m] = sum_k X[m*K + k] * w[k] + bias[0]
 * (bias may be null); X packed [M, K], dtype 0 fp32 / 1 bf16, K % 4 == 0, K <= 1024; w / bias fp32.  One warp per row. */
int ag_rowdot(const void* X, int32_t x_dtype, const float* w, const float* bias, int64_t M, int64_t K, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Last generator layer, Conv1d(C -> 1, k) over the dense channel-last buffer (audiogan.py:403-407, :467): HBM-bound
 * streaming kernels (no tensor-core work with one output channel).  X points at the first tap row of output 0,
 * batch stride x_bs floats, C % 4 == 0, weights w[j*C + c]:
 *   fwd:   out[b,t] = bias + sum_{j<k,c<C} X[b, t+j, c] * w[j*C+c]
 *   dgrad: dX[b, t', c] = sum_j g[b, t'-j] * w[j*C+c]        for t' in [0, T+k-1)   (plain store)
 *   wgrad: dw[j*C+c] += sum_{b,t} g[b,t] * X[b, t+j, c];  dw[k*C] += sum_{b,t} g[b,t]
 * X / dX: dtype 0 fp32 / 1 bf16 (strides in elements); out, g, w, dw fp32.
 * ------------------------------------------------------------------------------------------ */
int ag_conv1out_fwd(const void* X, int32_t x_dtype, int64_t x_bs, int64_t C, int32_t k, const float* w, const float* bias,
                    float* out, int64_t B, int64_t T, void* stream);
int ag_conv1out_dgrad(const float* g, const float* w, void* dX, int32_t dx_dtype, int64_t dx_bs, int64_t C, int32_t k, int64_t B,
                      int64_t T, void* stream);
int ag_conv1out_wgrad(const float* g, const void* X, int32_t x_dtype, int64_t x_bs, int64_t C, int32_t k, float* dw, int64_t B,
                      int64_t T, void* stream);

/* dst[b*d_bs + t*d_rs + c*d_cs] (+)= src[b*s_bs + t*s_rs + c*s_cs]: frame assembly into the dense
 * generator buffer (audiogan.py:462-464) and its gradient read-back. */
int ag_copy3d(void* dst, int64_t d_bs, int64_t d_rs, int64_t d_cs, const void* src, int64_t s_bs, int64_t s_rs,
              int64_t s_cs, int64_t B, int64_t T, int64_t C, int32_t accumulate, int32_t src_dtype, int32_t dst_dtype,
              void* stream);
/* Frame assembly (audiogan.py:462-464) into the first channel slot of the generator's dense channel-last buffer:
 * dst[b*d_bs + t*d_rs + 0] = src[b*s_bs + t], dst[.. + 1 .. slot-1] = 0 for t < L (dst dtype 0 fp32 / 1 bf16). */
int ag_frames_to_slot(void* dst, int32_t dst_dtype, int64_t d_bs, int64_t d_rs, int32_t slot, const float* src, int64_t s_bs, int64_t B,
                      int64_t L, void* stream);
/* Zero the pad rows [0, head) and [tail0, rows) of every batch of a packed channel-last buffer [B, rows, row_bytes] (the zero
 * padding every conv view relies on, audiogan.py:272 / :490 `padding=`): one launch instead of two strided fills.
 * 16-byte stores when row_bytes % 16 == 0 and buf is 16-byte aligned, else 4- / 2-byte stores (row_bytes % 2 == 0). */
int ag_zero_pads(void* buf, int64_t B, int64_t rows, int64_t row_bytes, int64_t head, int64_t tail0, void* stream);
/* The same for up to 8 buffers in ONE launch (a conv stack allocates 2-5 padded activation buffers per pass; `e` is a host array). */
typedef struct { void* buf; int64_t B, rows, row_bytes, head, tail0; } ag_pad_entry;
int ag_zero_pads_multi(const ag_pad_entry* e, int32_t n, void* stream);
/* out[b, n] = sum_t in[b, t, n]; in: dtype 0 fp32 / 1 bf16 (N even) */
int ag_rowgroup_sum(const void* in, int32_t dtype, float* out, int64_t B, int64_t T, int64_t N, void* stream);
/* dst[b, t, c] (channel-last, row stride dst_rs, batch stride dst_bs) <-> src[b, c, t] */
int ag_transpose_bct(const float* src, float* dst, int64_t B, int64_t C, int64_t T,
                     int64_t dst_bs, int64_t dst_rs, int32_t to_channel_last, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor optimizer with the reference's per-tensor clip (audiogan.py:232-253,
 * :693-694, :786-788, :909-921).  Table entries in DEVICE memory.
 *   pass 1 (ag_mt_sqnorm): sqnorm[i] = sum g^2, flags[0] |= any NaN, flags[1] |= any |g| > big (big <= 0: the reference's
 *     1e5, audiogan.py:240; data-parallel callers holding an un-normalised SUM over ranks pass 1e5 * world)
 *   pass 2 (ag_mt_rmsprop / ag_mt_adam): g' = g * min(1, clip/||g||) (clip <= 0: none), then
 *     RMSprop: sq = alpha*sq + (1-alpha)*g'^2 ; p -= lr * g' / (sqrt(sq) + eps)
 *     Adam   : m,v moments with bias correction from `step`.
 * ------------------------------------------------------------------------------------------ */
typedef struct ag_mt_entry {
  float* p; const float* g; float* s1; float* s2;   /* s1 = sq (RMSprop) / m (Adam); s2 = v (Adam) */
  int64_t n;
} ag_mt_entry;
int ag_mt_sqnorm(const ag_mt_entry* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                 int32_t nchunks, int32_t chunk, float* sqnorm, int32_t* flags, float big, void* stream);
/* clip_grad alone (audiogan.py:243-253): g *= clip/||g|| in place where ||g|| > clip. */
int ag_mt_clip(const ag_mt_entry* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
               int32_t nchunks, int32_t chunk, const float* sqnorm, float clip, void* stream);
int ag_mt_rmsprop(const ag_mt_entry* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                  int32_t nchunks, int32_t chunk, const float* sqnorm, float clip, float gscale,
                  double lr, double alpha, double eps, void* stream);   /* doubles: 1 - alpha is rounded once, as torch.optim does */
int ag_mt_adam(const ag_mt_entry* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
               int32_t nchunks, int32_t chunk, const float* sqnorm, float clip, float gscale,
               double lr, double beta1, double beta2, double eps, int32_t step, void* stream);

/* ------------------------------------------------------------------------------------------
 * Step-wise generator recurrence (audiogan.py:428-460 for hidden sizes whose recurrent weights cannot stay resident on chip,
 * e.g. --gstatesize 2048): the per-frame gate / projection / transposed products are ag_gemm_nt_tc launches issued by the host
 * side, the entry points below are the point-wise parts between them.  All pointers are already offset to frame t; *_bs are
 * batch strides in elements; `hx` / `dgp` are the bf16 operand row blocks ([B, hx_ld] = [h_t | x_t], [B, dgp_ld] =
 * [dgates_t | dpx_{t-1}]) of the NEXT frame's GEMM.
 *   cell_fwd:     gates = act(gpre + pre_t) (order i, f, g, o), c_t = f c_{t-1} + i g, h_t = o tanh(c_t)
 *   proj_finish:  x_t = tanh(px[:, :F]) (fp32 + bf16 + hx[:, hx_off:]), stop logit px[:, F] -> sbuf
 *   dpx:          dpx_t = (dx_ext_t + dxpre)(1 - x_t^2), column F = ds_ext_t, pad columns 0 (fp32 + bf16 + dgp[:, dgp_off:])
 *   cell_bwd:     dh_t -> dgates_t (fp32 optional, bf16, dgp[:, :4H]); `dc` [B, H] carries dc between frames
 * ------------------------------------------------------------------------------------------ */
int ag_lstm_step_cell_fwd(const float* gpre, const float* pre, int64_t pre_bs, const float* cprev, int64_t c_bs, float* gates, int64_t g_bs,
                          float* cout, float* h32, int64_t h_bs, void* h16, void* hx, int64_t hx_ld, int32_t B, int32_t H, void* stream);
int ag_gen_step_proj_finish(const float* px, int32_t FP, float* x32, void* x16, int64_t x_bs, void* hx, int64_t hx_ld, int32_t hx_off,
                            float* sbuf, int64_t s_bs, int32_t B, int32_t F, void* stream);
int ag_gen_step_dpx(const float* dxpre, int64_t dxpre_ld, const float* dx_ext, int64_t dx_bs, const float* ds_ext, int64_t ds_bs, const float* xt,
                    int64_t x_bs, float* dpx, void* dpx16, int64_t dpx_bs, void* dgp, int64_t dgp_ld, int32_t dgp_off, int32_t B, int32_t F,
                    int32_t FP, void* stream);
int ag_lstm_step_cell_bwd(const float* dh, const float* gates, int64_t g_bs, const float* c, const float* cprev, int64_t c_bs, float* dc,
                          float* dg32, void* dg16, int64_t dg_bs, void* dgp, int64_t dgp_ld, int32_t B, int32_t H, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient all-reduce over NVLink peer memory (replaces NN.DataParallel's gradient gather, audiogan.py:379-410,
 * and the NCCL all-reduce between backward and the optimizer step).  buf_ptrs_dev / sig_ptrs_dev: DEVICE arrays of `world`
 * pointers -- rank r's gradient buffer (n floats, n % 4 == 0, 16-byte aligned) and rank r's signal pad (>= 64 int32, zeroed
 * once) as mapped into THIS process (symmetric memory).  ag_peer_allreduce: barrier, rank `rank` sums its 1/world range of all
 * ranks' buffers and stores the sums into all ranks' buffers, barrier; on return (in stream order) every buffer holds the
 * sum.  Every rank must issue the same sequence of calls.  nblocks <= 0: 2 x SM count.  Plain kernel launches: capturable.
 * ------------------------------------------------------------------------------------------ */
int ag_peer_barrier(void* const* sig_ptrs_dev, int32_t rank, int32_t world, void* stream);
int ag_peer_allreduce(void* const* buf_ptrs_dev, void* const* sig_ptrs_dev, int32_t rank, int32_t world, int64_t n,
                      int32_t nblocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOGAN_B200_H */
