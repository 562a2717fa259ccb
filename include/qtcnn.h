/* qtcnn.h — C ABI of the B200-native QuadtreeCNN hot path (libqtcnn.so).
 *
 * The reference (Avirup221/Multimodal-Hierarchical-CNN-for-Sun-Salutation-Pose-Classification) has no
 * FFI of its own: its only boundary is the nn.Module surface of `models.py`. Each entry point below
 * replaces the ATen/cuDNN primitive that one reference call site reaches; the citation names that call
 * site (paths relative to the reference root, "QS" = "Quadtree_from scratch").
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator); the library
 *     never allocates, frees or retains device memory; scratch is passed in as `ws` / `ws_bytes`
 *   - activations are channels-last bf16 (NHWC / NDHWC), statistics, biases and weight gradients fp32
 *   - `stream` is a cudaStream_t; launches are asynchronous, never synchronise, and are graph-capturable
 *   - return 0 on success, <0 invalid argument / unsupported shape (text via qt_last_error()),
 *     >0 a cudaError_t. There is no CPU, cuDNN or Triton fallback: unsupported shapes fail.
 */
#ifndef QTCNN_H_
#define QTCNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* qt_stream_t; /* cudaStream_t */

/* Convolution geometry (2-D convs use *_d = 1). Strides are in ELEMENTS of the (n, d, h, w) axes of a
 * channels-last view, so the four quadrant views of the level-1 split (QS/models.py:277-282) are plain
 * descriptors: groups = 4 with per-group element offsets. */
typedef struct qt_conv_desc {
  int n;
  int in_d, in_h, in_w, in_c;
  int out_c;
  int k_d, k_h, k_w;
  int stride_d, stride_h, stride_w;
  int pad_d, pad_h, pad_w;
  int groups;
  long long x_stride[4];
  long long y_stride[4];
  long long x_group_off[4];
  long long y_group_off[4];
} qt_conv_desc;

/* epilogue flags of qt_conv_fprop / qt_linear_fprop */
#define QT_EPI_BIAS 1
#define QT_EPI_RELU 2
#define QT_EPI_STATS 4
#define QT_EPI_OUT_F32 16

int qt_version(void);
const char* qt_last_error(void);
/* 1 if any bounded barrier wait inside a kernel timed out since the last call (debug aid); resets it. */
int qt_take_timeout_flag(void);
/* Tuning/debug switch: 0 routes 3x3 stride-1 convolutions through the generic gather kernel instead of the
 * persistent input-reuse kernel (both are tcgen05 paths). */
void qt_set_conv3x3_enabled(int on);
/* Developer knobs (value 0 = default everywhere). key 0 / 1: pipeline shape of the generic weight-gradient / K-major
 * kernels; 2: generic gather for the stem; 3: generic weight-gradient kernel for 3x3 convs; 4: 7x7 maps through the
 * generic kernels; 5: staged (coalesced) conv3x3 write-out 1 = never, 2 = always (default: maps at least 20 wide);
 * 6: weight tiles by cp.async instead of TMA; 7: Conv3d through the generic gather kernels; 8: linear weight gradients always
 * through the split-K workspace; 9: ablation of the first-layer Conv3d weight-gradient kernel (1 = MMAs skipped, 2 = slab
 * copies skipped: results are then meaningless, profiles/r02_conv3d.md). Also settable through the QTCNN_TUNE="k=v,..."
 * environment variable of the Python binding. */
void qt_set_tuning(int key, int value);

/* ---- layout / packing ------------------------------------------------------------------------- */
/* images.to(device) feeding base_cnn.conv1 (QS/Quadtree_train.py:61, QS/models.py:222):
 * NCHW fp32 [n,3,h,w] -> zero-padded NHWC4 bf16 [n,h+7,w+8,4]. */
int qt_stem_pack_input(const float* x, void* xp, int n, int c, int h, int w, qt_stream_t stream);
/* The same packing from fp32 / bf16 / uint8 NCHW input. uint8 = decoded pixels: transforms.ToTensor + Normalize
 * (QS/dataloader.py:35-36) is applied on the device as v * scale[c] + shift[c] (scale = 1/(255 std), shift = -mean/std),
 * so one byte per value crosses PCIe instead of four. scale/shift may be NULL for the float types. */
#define QT_DTYPE_F32 0
#define QT_DTYPE_BF16 1
#define QT_DTYPE_U8 2
int qt_stem_pack_input_ex(const void* x, int dtype, const float* scale, const float* shift, void* xp, int n, int c, int h, int w,
                          qt_stream_t stream);
int qt_nchw_to_nhwc_bf16_ex(const void* x, int dtype, const float* scale, const float* shift, void* out, int n, int c, long long hw,
                            int c_pad, qt_stream_t stream);
int qt_nchw_f32_to_nhwc_bf16(const float* x, void* out, int n, int c, long long hw, int c_pad, qt_stream_t stream);
int qt_nhwc_bf16_to_nchw_f32(const void* x, float* out, int n, int c, long long hw, int c_pad, qt_stream_t stream);
/* fp32 parameter [cout][cin][taps] -> bf16 GEMM operands (forward [cout][taps][cin], dgrad [cin][taps][cout]). */
int qt_wpack_fprop(const float* w, void* wf, int cout, int cin, int taps, qt_stream_t stream);
int qt_wpack_dgrad(const float* w, void* wd, int cout, int cin, int taps, qt_stream_t stream);
/* both GEMM layouts from one read of the parameter (wd may be NULL). */
int qt_wpack_both(const float* w, void* wf, void* wd, int cout, int cin, int taps, qt_stream_t stream);
int qt_wpack_stem(const float* w, void* w8, int cout, int cin, int r, int s, qt_stream_t stream);
/* Conv3d forward weights with 32 input channels ([cout][32][3][3][3] fp32) -> bf16 [cout][2][9][64]: the operand of the
 * slab kernel's pair mode (qt_conv_plan(d, 0) == 2), where one 128-byte slab row carries two depth planes. */
int qt_wpack_conv3d_pair(const float* w, void* wp, int cout, qt_stream_t stream);
/* First Conv3d layer (3dcnn/models.py:107: Conv3d(3 -> 32)) on channel-padded NDHWC8 input: weights [32][cin <= 8][3][3][3] fp32
 * -> the resident operand tiles of the dedicated kernel (qt_conv_plan(d, 0) == 3 for in_c == 8, out_c == 32). */
int qt_wpack_conv3d_c8(const float* w, void* wb, int cout, int cin, qt_stream_t stream);
/* Every weight of a model repacked in ONE launch (what the training loop needs after optimizer.step(),
 * QS/Quadtree_train.py:72): fill w/wf/wd/cout/cin/taps of each item, call qt_wpack_item_plan (fills co_tile and
 * ci_tiles, returns the item's block count or -1), set first_block to the running sum of the block counts, copy the
 * table to device memory and pass it with the total block count and the largest taps value. */
typedef struct qt_wpack_item {
  const float* w;
  void* wf;
  void* wd; /* may be NULL */
  int cout, cin, taps;
  int co_tile, ci_tiles, first_block;
} qt_wpack_item;
int qt_wpack_item_plan(qt_wpack_item* item);
int qt_wpack_multi(const void* items_dev, int nitems, int total_blocks, int max_taps, qt_stream_t stream);
int qt_f32_to_bf16(const float* x, void* out, long long n, qt_stream_t stream);

/* ---- tensor-core implicit GEMM ------------------------------------------------------------------ */
/* nn.Conv2d / nn.Conv3d forward (torchvision resnet BasicBlock convs reached from QS/models.py:222-243,
 * quadrant_processor.0 QS/models.py:234-238 & 284-287, 3dcnn/models.py:107-139).
 * y = conv(x, wf) [+ bias] [ReLU]; with QT_EPI_STATS also writes per-tile column sum / sum-of-squares
 * partials [qt_conv_stat_rows][2][out_c] for the train-mode BatchNorm that follows. */
/* kernel selection of a pass (0 = fprop, 1 = dgrad, 2 = wgrad): 0 generic gather GEMM, 1 persistent slab kernel,
 * 2 (fprop of a 32-channel Conv3d only) slab kernel reading pair-packed weights (qt_wpack_conv3d_pair),
 * 3 (fprop of the 8 -> 32 channel first Conv3d only) dedicated kernel reading qt_wpack_conv3d_c8 weights. */
int qt_conv_plan(const qt_conv_desc* d, int pass);
int qt_conv_stat_rows(const qt_conv_desc* d);
size_t qt_conv_fprop_workspace_bytes(const qt_conv_desc* d);
int qt_conv_fprop(const qt_conv_desc* d, const void* x, const void* wf, void* y, const float* bias, float* stats,
                  int flags, void* ws, size_t ws_bytes, qt_stream_t stream);
/* convolution_backward, data gradient: dx = conv_transpose(dy, w) [+ dx when accumulate]. */
int qt_conv_dgrad(const qt_conv_desc* d, const void* dy, const void* wd, void* dx, int accumulate,
                  qt_stream_t stream);
/* convolution_backward, weight gradient: dw[cout][cin][taps] fp32 (+= when accumulate). */
size_t qt_conv_wgrad_workspace_bytes(const qt_conv_desc* d);
int qt_conv_wgrad(const qt_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, void* ws,
                  size_t ws_bytes, qt_stream_t stream);
/* base_cnn.conv1 (7x7, stride 2, pad 3, 3 channels; torchvision resnet.py:197) on the packed input. */
int qt_stem_stat_rows(int n, int h, int w);
int qt_stem_fprop(const void* xp, const void* w8, void* y, float* stats, int n, int h, int w, int cout,
                  qt_stream_t stream);
/* r3d_18 stem (3dcnn/models.py:224,268 -> torchvision video/resnet.py BasicStem: Conv3d(3,64,(3,7,7), stride (1,2,2), pad (1,3,3)))
 * on packed frames [n*t][h+7][w+8][4] (qt_stem_pack_input_ex per frame) with weights [cout][3][8][32] (qt_wpack_stem per depth
 * tap); y is NDHWC bf16 [n][t][h/2][w/2][cout]. The reference keeps this stem frozen: forward only. */
int qt_stem3d_stat_rows(int n, int t, int h, int w);
int qt_stem3d_fprop(const void* xp, const void* w24, void* y, float* stats, int n, int t, int h, int w, int cout,
                    qt_stream_t stream);
size_t qt_stem_wgrad_workspace_bytes(int n, int h, int w, int cout);
int qt_stem_wgrad(const void* xp, const void* dy, float* dw, int accumulate, int n, int h, int w, int cout, int cin,
                  void* ws, size_t ws_bytes, qt_stream_t stream);
/* nn.Linear on the tensor cores (classifier.0, QS/models.py:266-271): x bf16 [b][k] (row stride ldx),
 * w bf16 [n][k]; out bf16 or fp32 [b][n]. */
size_t qt_linear_workspace_bytes(int b, int n, int k);
int qt_linear_fprop(const void* x, long long ldx, const void* w, const float* bias, void* out, long long ldo,
                    int flags, int b, int n, int k, void* ws, size_t ws_bytes, qt_stream_t stream);
int qt_linear_dgrad(const void* dy, long long ldy, const void* wt, void* dx, long long ldx, int b, int n, int k,
                    void* ws, size_t ws_bytes, qt_stream_t stream);
int qt_linear_wgrad(const void* x, long long ldx, const void* dy, long long ldy, float* dw, int accumulate, int b,
                    int n, int k, void* ws, size_t ws_bytes, qt_stream_t stream);

/* ---- BatchNorm (train-mode batch statistics; torchvision resnet.py bn1/bn2, 3dcnn/models.py:109) ---- */
size_t qt_bn_workspace_bytes(int c);
int qt_bn_stats(const void* y, long long m, int c, float* partial, int partial_rows, qt_stream_t stream);
int qt_bn_finalize(const float* partial, int partial_rows, int c, double count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                   qt_stream_t stream);
/* qt_bn_finalize that also increments nn.BatchNorm's `num_batches_tracked` (int64 device scalar, may be NULL) in the same launch */
int qt_bn_finalize_tracked(const float* partial, int partial_rows, int c, double count, const float* gamma, const float* beta,
                           float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                           float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes, qt_stream_t stream);
/* eval mode: coefficients from the running statistics (mean/invstd are also returned for the backward). */
int qt_bn_eval_coeffs(int c, const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float eps, float* mean, float* invstd, float* scale, float* shift,
                      qt_stream_t stream);
int qt_bn_apply(const void* y, const float* scale, const float* shift, const void* residual, void* out, long long m,
                int c, int relu, qt_stream_t stream);
/* native_batch_norm_backward fused with the ReLU mask: dz = dout*(act>0) when act != NULL;
 * dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); optionally stores dz (identity-branch gradient).
 * eval_mode != 0: statistics were constants (model.eval(), e.g. Grad-CAM), dy = gamma*invstd*dz.
 * act == NULL with mask_scale/mask_shift: the ReLU mask is recomputed as y*mask_scale + mask_shift > 0 (the forward's
 * own expression), so the activation tensor is not read. */
int qt_bn_backward(const void* dout, const void* act, const void* y, const float* mean, const float* invstd,
                   const float* gamma, const float* mask_scale, const float* mask_shift, long long m, int c, float* dgamma,
                   float* dbeta, int accumulate, int eval_mode, void* dy, void* dz_out, void* ws, size_t ws_bytes,
                   qt_stream_t stream);
/* Stem tail fused (bn1 -> relu -> maxpool 3x3/s2/p1, torchvision resnet.py:198-200): one pass forward; backward =
 * max-pool gather + ReLU mask recomputed from y + BatchNorm backward, never materialising the full-resolution
 * activated map or its gradient. The backward works on 2x2 pixel blocks and needs even h and w (returns an error
 * otherwise; callers then use qt_maxpool2d_bwd + qt_bn_backward). yarg (optional, pooled shape): the raw conv output at each
 * window's arg-max, written by the forward; with it the backward takes its statistics from the pooled-size tensors. */
int qt_bn_relu_maxpool_fwd(const void* y, const float* scale, const float* shift, void* out, void* argmax, void* yarg, int n,
                           int h, int w, int c, qt_stream_t stream);
int qt_bn_relu_maxpool_bwd(const void* dpool, const void* argmax, const void* y, const void* yarg, const float* scale,
                           const float* shift, const float* mean, const float* invstd, const float* gamma, int n, int h, int w, int c,
                           float* dgamma, float* dbeta, int eval_mode, void* dy, void* ws, size_t ws_bytes,
                           qt_stream_t stream);
int qt_relu_backward(const void* dout, const void* act, void* dz, long long n, qt_stream_t stream);
int qt_colsum(const void* x, long long m, int c, float* out, int accumulate, void* ws, size_t ws_bytes,
              qt_stream_t stream);
int qt_add_bf16(const void* a, const void* b, void* out, long long n, qt_stream_t stream);

/* ---- pooling ----------------------------------------------------------------------------------- */
/* base_cnn.maxpool (MaxPool2d(3,2,1), torchvision resnet.py:200) and its backward. */
int qt_maxpool2d_fwd(const void* x, void* out, void* argmax, int n, int h, int w, int c, int ksize, int stride,
                     int pad, qt_stream_t stream);
int qt_maxpool2d_bwd(const void* dout, const void* argmax, void* dx, int n, int h, int w, int c, int ksize,
                     int stride, int pad, qt_stream_t stream);
/* Quadtree stage (QS/models.py:277-294): MaxPool2d(2,2)+flatten of the four quadrant maps and the global
 * average pool of layer4, written at their offsets of the fused feature row [global | TL | TR | BL | BR | ...]. */
int qt_quadtree_pool_fwd(const void* q, const void* l4, void* feat, int b, int qh, int qw, int cq, int ghw, int cg,
                         int ldf, qt_stream_t stream);
int qt_quadtree_pool_bwd(const void* dfeat, const void* q, void* dq, void* dl4, int b, int qh, int qw, int cq,
                         int ghw, int cg, int ldf, qt_stream_t stream);
/* ReLU + AdaptiveAvgPool2d((1,1)) of the level-1 / level-2 region convs (QS/models.py:21-30, 64-79). */
int qt_region_avgpool_fwd(const void* x, void* out, long long regions, int p, int c, long long ldo,
                          qt_stream_t stream);
int qt_region_avgpool_bwd(const void* dout, const void* x, void* dx, long long regions, int p, int c, long long ldo,
                          int relu_mask, qt_stream_t stream);

/* Conv3d block tail (3dcnn/models.py:108-135: BatchNorm3d + ReLU + MaxPool3d((1|2,2,2))) fused: forward writes the pooled
 * activation, the raw conv output at the arg-max (yarg) and the int8 arg-max code (both NULL in inference); backward returns dy
 * w.r.t. the raw conv output, dgamma/dbeta and (optionally) the gradient of the conv bias, without the full-size activation.
 * Replaces qt_bn_apply + qt_maxpool3d_fwd and qt_maxpool3d_bwd + qt_bn_backward + qt_colsum; bit-identical to them. */
int qt_bn_relu_maxpool3d_fwd(const void* y, const float* scale, const float* shift, void* out, void* yarg, void* argmax, int n,
                             int d, int h, int w, int c, int kd, int kh, int kw, qt_stream_t stream);
int qt_bn_relu_maxpool3d_bwd(const void* dpool, const void* argmax, const void* y, const void* yarg, const float* scale,
                             const float* shift, const float* mean, const float* invstd, const float* gamma, int n, int d, int h,
                             int w, int c, int kd, int kh, int kw, float* dgamma, float* dbeta, float* dbias, int eval_mode,
                             void* dy, void* ws, size_t ws_bytes, qt_stream_t stream);
/* MaxPool3d with kernel == stride, no padding (3dcnn/models.py:111-135) on NDHWC bf16. */
int qt_maxpool3d_fwd(const void* x, void* out, void* argmax, int n, int d, int h, int w, int c, int kd, int kh, int kw,
                     qt_stream_t stream);
int qt_maxpool3d_bwd(const void* dout, const void* argmax, void* dx, int n, int d, int h, int w, int c, int kd, int kh,
                     int kw, qt_stream_t stream);
/* softmax-weighted sum of the 16 sub-quadrant vectors (attention gate, QS/models.py:86-90); fp32. */
int qt_attn_pool_fwd(const float* x, const float* scores, float* wts, float* out, int b, int r, int c, qt_stream_t stream);
int qt_attn_pool_bwd(const float* x, const float* wts, const float* dout, float* dx, float* dscores, int b, int r, int c,
                     qt_stream_t stream);

/* ---- small fp32 linears of the fusion head (numerical_mlp, classifier.3; QS/models.py:255-271) ------ */
int qt_small_linear_fwd(const void* x, int x_is_bf16, long long ldx, const float* w, const float* bias, int b, int n,
                        int k, int relu, float drop_p, unsigned long long seed, float* out, long long ldo,
                        void* out16, long long ldo16, qt_stream_t stream);
int qt_small_linear_bwd_dx(const void* dy, int dy_is_bf16, long long ldy, const float* w, int b, int n, int k,
                           const float* act, long long lda, float drop_p, unsigned long long seed, float* dx,
                           long long ldx, void* dx16, long long ldx16, qt_stream_t stream);
int qt_small_linear_bwd_dw(const void* dy, int dy_is_bf16, long long ldy, const void* x, int x_is_bf16,
                           long long ldx, int b, int n, int k, float* dw, float* db, int accumulate,
                           qt_stream_t stream);
int qt_relu_dropout(float* h, void* h16, long long n, float drop_p, unsigned long long seed, int relu,
                    qt_stream_t stream);
int qt_relu_dropout_bwd(const float* dout, const float* act, float* dz, void* dz16, long long n, float drop_p,
                        unsigned long long seed, int relu, qt_stream_t stream);


/* ---- loss and fused classifier tail (QS/Quadtree_train.py:44,64; QS/models.py:268-271,303) -------------------------- */
/* nn.CrossEntropyLoss() (mean reduction, int64 class labels) and, when dlogits != NULL, its gradient
 * (softmax - onehot) * grad_scale * (*upstream) in the same pass (grad_scale = 1/b for the mean; upstream: device scalar,
 * the gradient arriving at the loss, NULL = 1). loss_rows [b] and the zero-initialised device word `counter` (reset by
 * the kernel) implement a fixed-order mean. nc <= 32. */
int qt_cross_entropy(const float* logits, long long ld, const long long* labels, int b, int nc, float grad_scale,
                     const float* upstream, float* loss_rows, float* loss_mean, float* dlogits, unsigned int* counter,
                     qt_stream_t stream);
/* Everything behind the classifier.0 GEMM in one launch: h [b][nhid] fp32 pre-activation -> ReLU + Dropout (h updated
 * in place, optional bf16 copy h16) -> classifier.3 -> logits [b][nc]; with labels also the cross-entropy loss. */
int qt_head_tail_fwd(float* h, void* h16, int nhid, const float* w3, const float* b3, int nc, const long long* labels, int b,
                     float drop_p, unsigned long long seed, float* logits, float* loss_rows, float* loss_mean,
                     unsigned int* counter, qt_stream_t stream);
/* Backward of the tail: with labels dlogits = (softmax(logits) - onehot) * grad_scale * (*upstream) is produced (upstream:
 * device scalar, NULL = 1), without labels it is read; dh16 [b][nhid] bf16 = (dlogits . w3) through the ReLU / dropout gate
 * recorded in act (= h after qt_head_tail_fwd). */
int qt_head_tail_bwd(const float* act, int nhid, const float* w3, int nc, const float* logits, const long long* labels,
                     float grad_scale, const float* upstream, float drop_p, unsigned long long seed, float* dlogits, void* dh16,
                     int b, qt_stream_t stream);

/* ---- optimizer: optim.Adam(params, lr, weight_decay) of the scripts (QS/Quadtree_train.py:45) + clip_grad_norm_
 * (3dcnn/train_3D_Quadtree_cnn_model.py:123), every parameter in one launch -------------------------------------------- */
typedef struct qt_adam_group {
  float step_size;    /* lr / (1 - beta1^t) */
  float beta1, beta2;
  float eps;
  float weight_decay; /* L2, added to the gradient (torch.optim.Adam, not AdamW) */
  float inv_bc2_sqrt; /* 1 / sqrt(1 - beta2^t) */
  float omb1, omb2;   /* 1 - beta1, 1 - beta2, rounded from double like torch's scalar arguments */
} qt_adam_group;
typedef struct qt_adam_item {
  float* p;       /* fp32 parameter, updated in place */
  const float* g; /* gradient */
  float* m;       /* exp_avg */
  float* v;       /* exp_avg_sq */
  void* wf;       /* optional bf16 GEMM copy [cout][taps][cin] refreshed by the same pass (NULL: plain tensor) */
  void* wd;       /* optional bf16 copy [cin][taps][cout] */
  long long n;
  int cout, cin, taps;
  int co_tile, ci_tiles, first_block; /* co_tile / ci_tiles filled by qt_adam_item_plan, first_block by the caller */
  int group, pad;
} qt_adam_item;
/* returns the item's block count (or -1); first_block = running sum of the counts, table copied to device memory */
int qt_adam_item_plan(qt_adam_item* item);
/* clip_coef: device scalar multiplied into every gradient (from qt_grad_clip_coef), or NULL; grad_scale: host scalar
 * multiplied into every gradient as well (1/world after a SUM all-reduce of the gradients, else 1) */
int qt_adam_multi(const void* items_dev, int nitems, int total_blocks, int max_taps, const qt_adam_group* groups, int ngroups,
                  const float* clip_coef, float grad_scale, qt_stream_t stream);
typedef struct qt_norm_item {
  const float* g;
  long long n;
  int first_block, pad;
} qt_norm_item;
int qt_grad_norm_blocks(long long n);
/* total_norm = grad_scale * ||all gradients||_2 (fixed-order reduction), coef = min(1, max_norm / (total_norm + 1e-6));
 * partial: fp32 scratch [total_blocks]. */
int qt_grad_clip_coef(const void* items_dev, int nitems, int total_blocks, float max_norm, float grad_scale, float* partial,
                      float* total_norm, float* coef, qt_stream_t stream);


/* ---- nn.LSTM(batch_first=True) layers (3dcnn/models.py:144-158,200-203; cnn+lstm/models.py:43-49,82-85) ---------------- */
int qt_transpose_f32(const float* in, float* out, int rows, int cols, qt_stream_t stream);
/* Recurrence of one layer over the whole sequence on clusters of 8 CTAs with the recurrent weights resident in shared memory:
 * xproj [b][t][4h] fp32 = x Wih^T + b_ih for every step (one qt_small_linear_fwd over b*t rows; nn.LSTM's inter-layer dropout is
 * applied to x before that), whh_t [h][4h] (transposed torch weight, gate order i,f,g,o), bhh [4h] or NULL, zero initial state.
 * Outputs hseq / hprev / cseq [b][t][h] and the activated gates [b][t][4h]. h <= 256. */
int qt_lstm_layer_fwd(const float* xproj, const float* whh_t, const float* bhh, int b, int t, int h, float* hseq, float* hprev,
                      float* cseq, float* gates, qt_stream_t stream);
/* BPTT of one layer: dhseq [b][t][h] (gradient reaching every h_t; out_drop_p/seed: mask of the dropout that sat between
 * this layer's output and its consumer), whh [4h][h] (torch layout) -> dgates [b][t][4h] (pre-activation gate gradients).
 * dX = dgates . wih, dWih = dgates^T x_used, dWhh = dgates^T hprev, db = colsum(dgates) are plain batched products
 * (qt_small_linear_bwd_dx / _dw). */
int qt_lstm_layer_bwd(const float* dhseq, float out_drop_p, unsigned long long seed, const float* whh, const float* gates,
                      const float* cseq, int b, int t, int h, float* dgates, qt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QTCNN_H_ */
