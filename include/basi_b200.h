/*
 * basi_b200.h -- C ABI of the B200-native BAIS PSPNet hot path.
 *
 * The reference (alisure-ml/Instance-Segment-BASI) has no FFI: the boundary its
 * hot path sits behind is tf.Session.run(fetches, feed_dict)
 *   train     back/2AddClass/BAISRunnerTrain.py:158-168
 *   inference back/4BorderClass/BAISRunnerOne.py:53-55, BAISRunnerGUI.py:75-76
 * and every TF op reached from there (call sites in back/2AddClass/BAISPSPNet.py:118-253).
 * Each entry point below replaces one of those TF call sites (cited per function) so
 * that a ctypes stub inside the reference's Network / Train / Runner classes can call
 * it.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch / C++ types.
 *  - every data pointer is a DEVICE pointer owned by the caller; nothing is allocated
 *    or freed inside except through basi_tc_conv_create / _destroy.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host
 *    synchronisation happens inside (safe under CUDA-graph capture).
 *  - return value: 0 on success, a negative BASI_E_* code otherwise;
 *    basi_last_error() gives the text for the calling thread.
 *  - activations are NHWC; `ld` is the element distance between two consecutive pixels
 *    (== c for a dense tensor, larger for a channel slice of the PSP concat buffer).
 *  - convolution weights are HWIO float32 ([kh][kw][Cin][Cout]), the TF variable layout.
 */
#ifndef BASI_B200_H_
#define BASI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BASI_OK 0
#define BASI_E_INVALID (-1)  /* bad argument / unsupported shape */
#define BASI_E_CUDA (-2)     /* CUDA runtime error, see basi_last_error() */
#define BASI_E_NOGPU (-3)    /* no usable sm_100 device */

#define BASI_F32 0
#define BASI_BF16 1

/* Per-channel batch-norm accumulators are replicated: sums / dsums are double [BASI_BN_REPLICAS][2*C]; blocks add
 * into replica (block % BASI_BN_REPLICAS) and the finalize step sums the replicas (shorter same-address atomic
 * chains at L2). */
#define BASI_BN_REPLICAS 8

typedef struct basi_tensor {
  void* ptr;     /* device pointer to element (0,0,0,0) */
  int32_t n, h, w, c;
  int32_t ld;    /* elements between consecutive pixels (>= c) */
  int32_t dtype; /* BASI_F32 | BASI_BF16 */
} basi_tensor;

/* tf.nn.conv2d / tf.pad + tf.nn.atrous_conv2d geometry (BAISPSPNet.py:118-146).
 * TF 'SAME' is expressed by the caller as explicit pad_t / pad_l (the "before" pads);
 * bottom/right padding is implied by the output size. */
typedef struct basi_conv_desc {
  int32_t kh, kw;
  int32_t stride;
  int32_t dil;
  int32_t pad_t, pad_l;
  int32_t relu; /* fuse ReLU into the fprop epilogue (class_attention_conv) */
} basi_conv_desc;

const char* basi_last_error(void);
int basi_version(void);
/* 0 = this build's 16-bit storage type (dtype BASI_BF16 in basi_tensor) is bfloat16 (libbasi_b200.so), 1 = IEEE fp16
 * (libbasi_b200_f16.so: the same kernels compiled with -DBASI_HALF_FP16; engine precision "f16"). */
int basi_half_format(void);
/* number of SMs of the current device (148 on B200), negative on error */
int basi_sm_count(void);
/* SM partition for the backward pass (no reference counterpart: TF's stream executor schedules its own kernels).
 * While main_sms > 0 every grid-size rule of the library (plan creation and launch) sees main_sms SMs instead of the
 * hardware count; wgrad_ctas > 0 is the CTA target of the tensor-core weight-gradient plans created meanwhile.
 * (0, 0) restores the defaults.  Process-wide, not thread-safe: the engine sets it around plan creation / enqueue. */
int basi_set_sm_budget(int main_sms, int wgrad_ctas);
int basi_memset(void* ptr, int value, int64_t bytes, void* stream);

/* ---- A3 / F3: label encodings and click sampling on the device (integer work, bit-exact) ----
 * ann: uint8 [B][per_image] instance-id maps at P x P (0 background, 255 border, k instance k); nums[b]: the attended
 * instance.  mode 0 binary (back/2AddClass/BAISData.py:150-151); mode 1 the 4-class border encoding with its uint8
 * wrap-around (back/4BorderClass/BAISData.py:143-160, has_255=True: 0 other, 1 attention, 2 border, 3 background);
 * mode 2 the 3-class encoding (same file, has_255=False); mode 3 COCO (back/5COCO/BAISData.py:361-369: ann = uint8 sum
 * of all instance masks, attention = the attended mask: 0 background, 1 other, 2 attention).  Writes int32 and / or
 * float32 labels (either pointer may be NULL). */
int basi_label_encode(const uint8_t* ann, const uint8_t* attention, const int32_t* nums, int mode, int32_t* out_i32,
                      float* out_f32, int B, int64_t per_image, void* stream);
/* Click sampling (back/2AddClass/BAISData.py:63-66: where = np.argwhere(ann == 1); where[np.random.randint(0, len)] * ratio):
 * basi_click_count gives len(where) per image (the host draws k with the reference's own RNG call),
 * basi_click_select writes ratio * (row, col) of the k-th matching pixel in row-major order. */
int basi_click_count(const void* labels, int labels_are_f32, int target, int B, int per_image, int32_t* counts,
                     void* stream);
int basi_click_select(const void* labels, int labels_are_f32, int target, int B, int H, int W, const int32_t* k,
                      int ratio, int32_t* clicks, void* stream);

/* ---- A1/A2: Data._mask_gaussian + np.concatenate (back/2AddClass/BAISData.py:189-202, :79-80) ----
 * img: uint8 [B,H,W,3] (img_is_f32 = 0) or float32 [B,H,W,3] already /255 (img_is_f32 = 1).
 * clicks: int32 [B][2] = (y0, x0).  lut: float32 table indexed by d2 = (x-x0)^2 + (y-y0)^2,
 * built on the host in float64 and rounded once (bit-exact with the numpy expression).
 * out: float32 [B,H,W,4]. */
int basi_clickmap_pack(const void* img, int img_is_f32, const int32_t* clicks, const float* lut,
                       int64_t lut_len, float* out, int B, int H, int W, void* stream);

/* ---- A4/A5: Network.conv / Network.atrous_conv (BAISPSPNet.py:122-146) ----
 * Generic CUDA-core implicit-GEMM path (fp32 accumulate); x/y may be f32 or bf16. */
int basi_conv_fprop(const basi_conv_desc* d, const basi_tensor* x, const float* w, const float* bias,
                    const basi_tensor* y, void* stream);
/* adjoint w.r.t. the input (what tf.gradients derives for train_op, BAISRunnerTrain.py:117).
 * accumulate != 0: dx += result. */
/* conv1_1 (fp32 NHWC4 input, 3x3, 32 or 64 output channels) with the batch-norm statistics of its output fused into
 * the kernel (same protocol as basi_tc_conv_set_bn_stats: sums = double[BASI_BN_REPLICAS][2][C], zeroed by the caller;
 * the last block writes bnp = [mean | istd | gamma*istd | beta] when bnp != NULL). */
int basi_stem_fprop_stats_supported(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y);
int basi_stem_fprop_stats(const basi_conv_desc* d, const basi_tensor* x, const float* w, const basi_tensor* y,
                          double* sums, const float* gamma, const float* beta, double count, float eps, float* bnp,
                          uint32_t* counter, void* stream);
int basi_conv_dgrad(const basi_conv_desc* d, const basi_tensor* dy, const float* w, const basi_tensor* dx,
                    int accumulate, void* stream);
/* adjoint w.r.t. the HWIO weights; ADDS into dw (and dbias when non-NULL). */
int basi_conv_wgrad(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* dy, float* dw,
                    float* dbias, void* stream);

/* x[:, ::stride, ::stride, :] and its adjoint (dx (+)= dy scattered to the sampled pixels, zero elsewhere): the
 * 1x1 stride-2 VALID convolutions conv3_1_1x1_proj / conv3_1_1x1_reduce (BAISPSPNet.py:310,314) are computed as the
 * stride-1 1x1 convolution of the subsampled tensor. */
int basi_subsample_fwd(const basi_tensor* x, int stride, const basi_tensor* y, void* stream);
int basi_subsample_bwd(const basi_tensor* dy, int stride, const basi_tensor* dx, int accumulate, void* stream);

/* ---- A6/A7: Network.batch_normalization (+relu, +add) (BAISPSPNet.py:204-236, :148-150, :171-173) ----
 * sums: double [BASI_BN_REPLICAS][2*C] (sum x, sum x^2), ADDED into (caller zeroes); counter: one zeroed uint32
 * per launch.
 * When gamma != NULL the last block to finish also writes bnp (fused finalize):
 * bnp: float [4*C] = [mean | istd | gamma*istd | beta] (batch mean, biased variance, eps inside the sqrt). */
int basi_bn_stats(const basi_tensor* x, double* sums, const float* gamma, const float* beta, double count, float eps,
                  float* bnp, uint32_t* counter, void* stream);
int basi_bn_finalize(const double* sums, const float* gamma, const float* beta, double count, float eps,
                     float* bnp, int C, void* stream);
/* out = act((x-mean)*scale+beta [+ res | + (res-res_mean)*res_scale+res_beta]); res / res_bnp may be NULL. */
int basi_bn_apply(const basi_tensor* x, const float* bnp, const basi_tensor* res, const float* res_bnp,
                  int relu, const basi_tensor* out, void* stream);
/* dsums (double [BASI_BN_REPLICAS][2*C]) += (sum dy, sum dy*xhat).  dy = dout * (out > 0) when out != NULL; else, when
 * relu_from_x, dy = dout * ((x-mean)*scale+beta > 0) (plain BN+ReLU: the mask is recomputed, out is not read).
 * When coef != NULL the last block also does the finalize: dgamma += sum dy*xhat, dbeta += sum dy,
 * coef (float [2*C]) = dsums / count. */
int basi_bn_bwd_reduce(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x, const float* bnp,
                       int relu_from_x, double* dsums, double count, float* dgamma, float* dbeta, float* coef,
                       uint32_t* counter, void* stream);
int basi_bn_bwd_finalize(const double* dsums, double count, float* dgamma, float* dbeta, float* coef, int C,
                         void* stream);
/* dx = gamma*istd*(dy - coef0 - xhat*coef1); dres (optional) (+)= dy. */
int basi_bn_bwd_apply(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x, const float* bnp,
                      const float* coef, int relu_from_x, const basi_tensor* dx, const basi_tensor* dres,
                      int dres_accumulate, void* stream);

/* Packed ReLU mask for the residual junctions (relu(bn(x) + res), BAISPSPNet.py:171-173): basi_bn_apply_bits also
 * writes one bit per output element (maskbits[row * c/8 + g] bit i <=> out[row][8*g + i] > 0) and the backward pair
 * reads those bits instead of the stored output (1/16 of the bytes).  bf16, c % 128 == 0; query first. */
int basi_bn_maskbits_supported(const basi_tensor* x);
int basi_bn_apply_bits(const basi_tensor* x, const float* bnp, const basi_tensor* res, const float* res_bnp, int relu,
                       const basi_tensor* out, unsigned char* maskbits, void* stream);
int basi_bn_bwd_reduce_bits(const basi_tensor* dout, const unsigned char* maskbits, const basi_tensor* x,
                            const float* bnp, double* dsums, double count, float* dgamma, float* dbeta, float* coef,
                            uint32_t* counter, void* stream);
int basi_bn_bwd_apply_bits(const basi_tensor* dout, const unsigned char* maskbits, const basi_tensor* x,
                           const float* bnp, const float* coef, const basi_tensor* dx, const basi_tensor* dres,
                           int dres_accumulate, void* stream);

/* Cooperative streamed backward: basi_bn_bwd_reduce[_bits] + basi_bn_bwd_apply[_bits] in ONE cooperative launch (grid
 * barrier between the two passes, second pass in reverse row order for L2 reuse).  For the tensors too large for
 * basi_bn_bwd_fused: residual junctions (mask = `out` or `maskbits`, optional residual gradient `dres`) and the
 * 160x160 layers.  `barrier` is a zero-initialised word.  Query with basi_bn_bwd_coop_supported (1 = yes). */
int basi_bn_bwd_coop_supported(const basi_tensor* x, int has_out, int has_bits, int dres_accumulate);
int basi_bn_bwd_coop(const basi_tensor* dout, const basi_tensor* out, const unsigned char* maskbits,
                     const basi_tensor* x, const float* bnp, int relu_from_x, double* dsums, double count,
                     float* dgamma, float* dbeta, float* coef, uint32_t* barrier, const basi_tensor* dx,
                     const basi_tensor* dres, int dres_accumulate, void* stream);

/* Resident backward: basi_bn_bwd_reduce + basi_bn_bwd_apply (no stored-output mask, no residual) in ONE cooperative
 * launch that keeps dout and x in shared memory between the two phases (reads each once).  Needs dense rows
 * (ld == c) and a tensor small enough that 2 * bytes(x) / #SMs fits in shared memory; query with
 * basi_bn_bwd_fused_supported (1 = yes).  `barrier` is a zero-initialised grid-barrier word; `coef` may be NULL. */
int basi_bn_bwd_fused_supported(const basi_tensor* x);
int basi_bn_bwd_fused(const basi_tensor* dout, const basi_tensor* x, const float* bnp, int relu_from_x, double* dsums,
                      double count, float* dgamma, float* dbeta, float* coef, uint32_t* barrier,
                      const basi_tensor* dx, void* stream);

/* F1: adjoint of the slim conv2d epilogue y = relu(conv + bias): dy *= (y > 0) in place (relu != 0) and
 * dbias[c] += sum of the masked dy over the pixels (dbias may be NULL). */
int basi_bias_relu_bwd(const basi_tensor* dy, const basi_tensor* y, int relu, float* dbias, void* stream);
/* F2: Net.add of the top-level LinkNet (BAISNet.py:244, tf.add_n of two tensors, no batch norm) and its adjoint:
 * da (+)= dout, db (+)= dout (either may be NULL). */
int basi_add_fwd(const basi_tensor* a, const basi_tensor* b, const basi_tensor* out, void* stream);
int basi_add_bwd(const basi_tensor* dout, const basi_tensor* da, int acc_a, const basi_tensor* db, int acc_b,
                 void* stream);
/* F4: Net.relu on a MATERIALISED residual sum (back/8AttentionU/BAISNet.py:163-165: the hand-unrolled trunk reads both
 * the sum `net_input = Net.add(...)` -- next 1x1_reduce / 1x1_proj -- and its ReLU -- next shortcut) and its adjoint
 * dx (+)= dy * (y > 0).  x / y / dy / dx: same shape and dtype, channel count a multiple of the vector width. */
int basi_relu_fwd(const basi_tensor* x, const basi_tensor* y, void* stream);
int basi_relu_bwd(const basi_tensor* dy, const basi_tensor* y, const basi_tensor* dx, int accumulate, void* stream);
/* F1: tf.one_hot(labels, depth=2) as float32 pairs (targets of the 2-channel weighted CE of variant B,
 * back/90AttentionSingle2/BAISRunnerTrain.py:128-131); labels float32 {0,1}, out float32 [n][2]. */
int basi_onehot2_f32(const float* labels, float* out, int64_t n, void* stream);
/* F1 (variant B trunk): slim.max_pool2d(net, [2, 2]) of vgg_16 (slim/nets/vgg.py:188-196): 2x2 / stride 2 VALID,
 * y = [N, H/2, W/2, C]; argmax: one byte per output element (window index of the first maximum). */
int basi_maxpool2s2_fwd(const basi_tensor* x, const basi_tensor* y, uint8_t* argmax, void* stream);
int basi_maxpool2s2_bwd(const basi_tensor* dy, const uint8_t* argmax, const basi_tensor* dx, int accumulate,
                        void* stream);
/* ---- A8: Network.max_pool 3x3 s2 SAME (:152-155, :269), Network.avg_pool k=s VALID (:157-160) ---- */
int basi_maxpool3s2_fwd(const basi_tensor* x, const basi_tensor* y, uint8_t* argmax, void* stream);
int basi_maxpool3s2_bwd(const basi_tensor* dy, const uint8_t* argmax, const basi_tensor* dx, int accumulate,
                        void* stream);
int basi_avgpool_fwd(const basi_tensor* x, int k, const basi_tensor* y, void* stream);
int basi_avgpool_bwd(const basi_tensor* dy, int k, const basi_tensor* dx, int accumulate, void* stream);

/* All pyramid pools of one tensor together (BAISPSPNet.py:683-710: windows 40/20/13/6 of conv5_3 at 320^2): ks[p]
 * is the window (= stride) of pool p, ys[p] its output ([n, h/k, w/k, c]).  Forward: one pass over x writes the
 * per-row window sums of every pool to `scratch` (basi_avgpool_multi_scratch_floats() floats, no initialisation
 * needed), a second small pass adds the rows of each cell in a fixed order -- no atomics, bit-reproducible.  The
 * adjoint adds (accumulate != 0) or writes every pool's contribution to dx in one read-modify-write pass.  At most
 * 4 pools; basi_avgpool_multi_scratch_floats returns -1 for a group it cannot take (the caller then uses
 * basi_avgpool_fwd / _bwd per pool). */
int64_t basi_avgpool_multi_scratch_floats(const basi_tensor* x, int n_pools, const int* ks);
int basi_avgpool_multi_fwd(const basi_tensor* x, int n_pools, const int* ks, const basi_tensor* const* ys,
                           float* scratch, void* stream);
int basi_avgpool_multi_bwd(const basi_tensor* const* dys, int n_pools, const int* ks, const basi_tensor* dx,
                           int accumulate, void* stream);

/* ---- A9: Network.resize_bilinear, align_corners=True (:242-244) ---- */
int basi_bilinear_ac_fwd(const basi_tensor* x, const basi_tensor* y, void* stream);
int basi_bilinear_ac_bwd(const basi_tensor* dy, const basi_tensor* dx, int accumulate, void* stream);

/* ---- A11: Network.multiply (relu(conv5_3) * logits[..., att]) (:246-249; 4BorderClass :246-252) ----
 * logits: float32 [n,h,w,nseg].  out = relu(feat) * logits[...,att]. */
int basi_gate_mul_fwd(const basi_tensor* feat, const float* logits, int nseg, int att, const basi_tensor* out,
                      void* stream);
/* dfeat (+)= dout*gate*(feat>0);  dlogits[...,att] += sum_c dout*relu(feat). */
int basi_gate_mul_bwd(const basi_tensor* dout, const basi_tensor* feat, const float* logits, int nseg, int att,
                      const basi_tensor* dfeat, int dfeat_accumulate, float* dlogits, void* stream);

/* ---- A12-A14: variant-B gating ops (click-gated features and the attention cascade) ----
 * A12  tf.multiply(feature, mask) with the mask broadcast over channels: back/7OLD/BAISNet.py:535,
 *      back/90AttentionSingle2/BAISNet.py:748,773-776; cascade gate back/8AttentionU/BAISNet.py:586-591.
 *      mask: float32 [n,h,w,nch], channel ch is used.  dmask may be NULL (the click map needs no gradient). */
int basi_mask_mul_fwd(const basi_tensor* feat, const float* mask, int nch, int ch, const basi_tensor* out,
                      void* stream);
int basi_mask_mul_bwd(const basi_tensor* dout, const basi_tensor* feat, const float* mask, int nch, int ch,
                      const basi_tensor* dfeat, int dfeat_accumulate, float* dmask, void* stream);
/* A13  softmax over C logits, channel `sel`, hard gate tf.where(p > thr, p, 0)
 *      (back/90AttentionSingle2/BAISNet.py:743-746, thr = 0.9; thr < 0 gives the plain softmax channel of
 *      back/8AttentionU/BAISNet.py:586).  gate: float32 [rows].  bwd: dlogits (+)= dgate * d gate / d logits. */
int basi_softmax_gate_fwd(const float* logits, int64_t rows, int C, int sel, float thr, float* gate, void* stream);
int basi_softmax_gate_bwd(const float* logits, const float* dgate, int64_t rows, int C, int sel, float thr,
                          float* dlogits, int accumulate, void* stream);
/* A14  tf.image.resize_nearest_neighbor (TF1 default): src = min(floor(dst * in / out), in - 1)
 *      (back/90AttentionSingle2/BAISNet.py:750,752,773; labels in BAISRunnerTrain.py:92-93). */
int basi_resize_nearest_fwd(const basi_tensor* x, const basi_tensor* y, void* stream);
int basi_resize_nearest_bwd(const basi_tensor* dy, const basi_tensor* dx, int accumulate, void* stream);

/* ---- A11: class_attention_conv (5x5 s5 on a 5x5 map) and Network.fc (:175-189) as skinny GEMMs ----
 * y[m][n] = act(sum_k a[m][k] w[k][n] + bias[n]). a may be f32/bf16 (dtype_a), y float32.  Any M (= batch): rows
 * are processed in chunks that fit the kernels' register / shared-memory budgets.  basi_skinny_supported is the
 * single predicate the host lowering asks (1 = the three entry points below accept this shape).
 * When N is a multiple of 4 the forward runs 128-bit kernels: w (and the workspace) must then be 16-byte aligned
 * (BASI_E_INVALID otherwise); dgrad / wgrad fall back to their scalar kernels on unaligned pointers. */
int basi_skinny_supported(int M, int K, int N);
int basi_skinny_fwd(const void* a, int dtype_a, int64_t lda, const float* w, const float* bias, float* y,
                    int M, int K, int N, int relu, void* stream);
/* The same with a caller-provided workspace of basi_skinny_fwd_workspace_floats(M, K, N) floats: the k-split partial
 * sums are written there and added in a fixed order, so the result is bit-reproducible (basi_skinny_fwd adds them
 * with fp32 atomics, whose order -- and therefore last-bit rounding -- varies from run to run).  workspace == NULL
 * is basi_skinny_fwd. */
int64_t basi_skinny_fwd_workspace_floats(int M, int K, int N);
int basi_skinny_fwd_ws(const void* a, int dtype_a, int64_t lda, const float* w, const float* bias, float* y,
                       int M, int K, int N, int relu, float* workspace, void* stream);
/* da[m][k] (+)= sum_n dy[m][n] w[k][n]   (da dtype_a) */
int basi_skinny_dgrad(const float* dy, const float* w, void* da, int dtype_a, int64_t lda, int M, int K, int N,
                      int accumulate, void* stream);
/* dw[k][n] += sum_m a[m][k] dy[m][n]; dbias[n] += sum_m dy[m][n] */
int basi_skinny_wgrad(const void* a, int dtype_a, int64_t lda, const float* dy, float* dw, float* dbias, int M,
                      int K, int N, void* stream);
/* dy *= (y > 0)  (ReLU adjoint for the skinny path), n elements float32 */
int basi_relu_bwd_f32(float* dy, const float* y, int64_t n, void* stream);

/* ---- A15: tf.reduce_mean(weighted_cross_entropy_with_logits) (2AddClass/BAISRunnerTrain.py:104-105) ----
 * loss_acc[0] += scale * sum(loss_i) (double);  dlogits = grad_scale * dloss_i/dx. */
int basi_wbce_fwd_bwd(const float* logits, const float* labels, float pos_weight, double scale,
                      float grad_scale, int64_t n, double* loss_acc, float* dlogits, void* stream);
/* F4 (back/8AttentionU/BAISRunnerTrain.py:176-180): the same loss on channel `sel` of C-channel rows
 * (tf.split(segment, 2, axis=3)[1]); dlogits [rows][C] gets the gradient in channel `sel` and zeros elsewhere. */
int basi_wbce_sel_fwd_bwd(const float* logits, int C, int sel, const float* labels, float pos_weight, double scale,
                          float grad_scale, int64_t rows, double* loss_acc, float* dlogits, void* stream);
/* F4: Net.sigmoid of the decoder logits (back/8AttentionU/BAISNet.py:527) -- cal_loss consumes the sigmoid outputs
 * as "logits" -- and its adjoint dx (+)= dy * y * (1 - y).  float32, n elements. */
int basi_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream);
int basi_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, int accumulate, void* stream);
/* ---- A15/A16: tf.reduce_mean(sparse_softmax_cross_entropy_with_logits) (4BorderClass :111-116) ---- */
int basi_softmax_ce_fwd_bwd(const float* logits, const int32_t* labels, int64_t rows, int C, double scale,
                            float grad_scale, double* loss_acc, float* dlogits, void* stream);

/* ---- (e) multi-GPU: averaged all-reduce of the flat gradient over NVLink peer memory ----
 * The data-parallel exchange of SURVEY section 8(e) as ONE full-machine kernel per rank instead of NCCL launches that
 * the persistent backward kernels starve.  bufs[q] / flags[q]: this process's mappings of rank
 * q's symmetric gradient buffer and flag array (uint32[world], zero at start-up); seq_dev: device uint32, 0 at
 * start-up on every rank.  Collective: every rank calls it once per step on its stream. */
int basi_p2p_allreduce_mean(void* const* bufs, void* const* flags, int rank, int world, int64_t n, uint32_t* seq_dev,
                            void* stream);

/* ---- A17: GradientDescentOptimizer (2AddClass/BAISRunnerTrain.py:115-117): w -= lr*g ----
 * lr is read from device memory so a captured graph can be replayed with a new rate.
 * w_bf16 (optional) receives the bf16 copy of the updated weights. */
int basi_sgd_step(float* w, const float* g, const float* lr_dev, int64_t n, void* w_bf16, void* stream);

/* ---- A18: predictions ---- */
/* out = logits > thr (2AddClass/BAISRunnerTrain.py:88) */
int basi_threshold(const float* logits, float thr, int32_t* out, int64_t n, void* stream);
/* first-max argmax over the last axis (BAISRunnerOne.py:41-45) */
int basi_argmax(const float* logits, int64_t rows, int C, int32_t* out, void* stream);
/* TF1 resize_bilinear(align_corners=False) to SxS then argmax(sigmoid) (BAISRunnerGUI.py:29-31) */
int basi_upsample_legacy_argmax(const float* logits, int B, int P_h, int P_w, int C, int S_h, int S_w,
                                int32_t* out, void* stream);
/* Split-operand (fp32-grade) mode: x float32 [N,H,W,C] -> y bf16 [N,H,W,3C] = [hi | mid | lo], x == hi + mid + lo to
 * 2^-24 (or [N,H,W,2C] = [hi | mid], x == hi + mid to 2^-17): the operand format of basi_tc_conv_create_split (A4/A5 at the reference's float32 precision on tcgen05). */
int basi_split3_bf16(const basi_tensor* x, const basi_tensor* y, void* stream);
/* dst[i] = (bf16) src[i] / (f32) src[i] helpers */
int basi_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int basi_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);

/* ---- tcgen05 / TMEM / TMA implicit-GEMM convolutions (bf16 in, fp32 accumulate) ----
 * One plan per (layer, pass); owns the TMA descriptors for fixed device pointers.
 * kind: 0 fprop (y = conv(x,w)), 1 dgrad (dx = conv^T(dy,w)), 2 wgrad (dw += x^T dy).
 * Weights are bf16 copies produced by basi_tc_pack_weights. */
#define BASI_TC_FPROP 0
#define BASI_TC_DGRAD 1
#define BASI_TC_WGRAD 2
typedef struct basi_tc_conv basi_tc_conv;
/* returns 1 if the tcgen05 path supports this geometry / channel counts, else 0 */
int basi_tc_conv_supported(int kind, const basi_conv_desc* d, const basi_tensor* in, const basi_tensor* out);
/* w_hwio f32 [taps][Cin][Cout] -> bf16 [taps][Cin][Cout] and bf16 [taps][Cout][Cin] */
int basi_tc_pack_weights(const float* w, void* w_io_bf16, void* w_oi_bf16, int taps, int cin, int cout,
                         void* stream);
/* The same for many layers in ONE launch.  table_dev: device array of n_layers entries
 *   struct { const float* w; void* w_io; void* w_oi; int32 taps, cin, cout, block_start, tiles_co, tiles_ci, pad0, pad1; }
 * (56 bytes each) with tiles_ci = ceil(cin / 32), tiles_co = ceil(cout / 32), block_start = running sum of
 * taps x tiles_ci x tiles_co, and
 * total_blocks = that sum over all layers. */
int basi_tc_pack_weights_multi(const void* table_dev, int n_layers, int total_blocks, void* stream);
/* ---- the same convolutions at the reference's float32 precision ("bf16x6"): operands are the bf16 [hi|mid|lo] parts
 * of float32 tensors (basi_split3_bf16), six bf16 MMAs per product accumulate in fp32 TMEM, outputs are float32.
 * Table entries with pad0 = 1 make basi_tc_pack_weights_multi write the matching split weight layouts
 * (w_oi: [taps][Cout][basi_tc_split_kcols(Cin)] for fprop, w_io: [taps][Cin][basi_tc_split_kcols(Cout)] for dgrad;
 * the buffers must be zero-initialised once).  x / y of *_supported_split are the logical float32 tensors.
 *   FPROP: a = x3, b = y (f32)    DGRAD: a = dy3, b = dx (f32, written / accumulated)    WGRAD: a = x3, b = dy3
 * parts = 3: [hi|mid|lo] operands, six products (float32-exact); parts = 2: [hi|mid] operands, three products
 * (hi*hi, hi*mid, mid*hi: 2^-17 unbiased representation error per operand), table entries with pad0 = 2. */
int basi_tc_conv_supported_split(int kind, const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y);
int basi_tc_split_kcols(int c, int parts);
int basi_tc_conv_create_split(int kind, const basi_conv_desc* d, const basi_tensor* a, const basi_tensor* b,
                              const void* w_split, float* dw, int accumulate, int parts, basi_tc_conv** out);
int basi_tc_conv_create(int kind, const basi_conv_desc* d, const basi_tensor* a, const basi_tensor* b,
                        const void* w_bf16, float* dw, int accumulate, basi_tc_conv** out);
/* fprop plans only: also accumulate the batch-norm statistics of the produced tensor in the epilogue (same
 * contract as basi_bn_stats: sums += (sum y, sum y^2) in double, counter = one zeroed uint32 per launch, and the
 * last CTA writes bnp when it is non-NULL), which removes the separate statistics pass over the conv output. */
int basi_tc_conv_set_bn_stats(basi_tc_conv* plan, double* sums, const float* gamma, const float* beta, double count,
                              float eps, float* bnp, uint32_t* counter);
/* A6 fused into A4: an fprop plan with fused statistics additionally writes out = [relu](BN(y)) -- normalised from the
 * fp32 TMEM accumulators behind a grid barrier (cooperative launch) -- and publishes bnp.  Returns 1 if the plan was
 * switched (then no basi_bn_apply is needed for this layer), 0 if the layer cannot be fused. */
int basi_tc_conv_set_bn_apply(basi_tc_conv* plan, const basi_tensor* out, int relu);
/* A6 backward fused into the A4/A5 input adjoint: a dgrad plan whose destination dx is the gradient wrt the raw conv
 * output x of the BN(+ReLU) layer feeding this convolution computes dA in TMEM, reduces sum(g) and sum(g * xhat) (mask
 * recomputed from x), meets the grid at a barrier (cooperative launch) and writes dx = BN'(dA); dgamma / dbeta are
 * added into their slots.  bnp = [mean | istd | gamma*istd | beta] of that layer, dsums = [BASI_BN_REPLICAS][2C]
 * doubles and *counter zero at the start of a step.  Returns 1 if the plan was switched (then no basi_bn_bwd_* call is
 * needed for that layer), 0 if it cannot be fused. */
int basi_tc_conv_set_bn_bwd(basi_tc_conv* plan, const basi_tensor* x, const float* bnp, int relu, double* dsums,
                            double count, float* dgamma, float* dbeta, uint32_t* counter);
/* fprop plans: y = [relu](conv(x, w) + bias[c]) in the epilogue (slim/nets/vgg.py:187-196 convolutions of variant B) */
int basi_tc_conv_set_bias(basi_tc_conv* plan, const float* bias, int relu);
int basi_tc_conv_run(basi_tc_conv* plan, void* stream);
void basi_tc_conv_destroy(basi_tc_conv* plan);

#ifdef __cplusplus
}
#endif
#endif /* BASI_B200_H_ */
