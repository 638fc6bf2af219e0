/*
 * eeseg.h — C ABI of libeeseg_b200.so: the sm_100a kernels behind the early-exit segmentation
 * hot path of MateusGilbert/ee_semantic_segmentation (BranchyDeepLabV3).
 *
 * The reference is 100 % Python/PyTorch and has no FFI of its own; the entry
 * points below are what a binding for this path would call. Each one names the reference
 * interface (file:line under /root/reference) whose inner work it replaces. The Python mirror of
 * the reference API that calls them lives in ee_semantic_segmentation_b200/ (ctypes; see
 * INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; no allocation inside;
 *   - all launchers are asynchronous on `stream` (a cudaStream_t passed as void*) and re-entrant;
 *   - return value: 0 = ok, non-zero = error (EESEG_ERR_*); eeseg_last_error() gives the text for
 *     the calling thread;
 *   - logits/probabilities are NCHW planes ("[N][C][HW]") unless a stride is passed explicitly;
 *     dtype codes: EESEG_F32 / EESEG_BF16;
 *   - class targets are int64, any value outside [0,C) is "void" (get_seg_datasets.py:79-86 maps
 *     255 -> C).
 */
#ifndef EESEG_H_
#define EESEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EESEG_ABI_VERSION 1

enum { EESEG_F32 = 0, EESEG_BF16 = 1, EESEG_U8 = 2 };
enum { EESEG_OK = 0, EESEG_ERR_ARG = 1, EESEG_ERR_CUDA = 2, EESEG_ERR_UNSUPPORTED = 3 };

int eeseg_abi_version(void);
const char* eeseg_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench.py's
 * gpu_launches claim) */
int64_t eeseg_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Confusion-matrix histogram.
 * Replaces SegMetric._compute_basics (seg_metrics.py:13-28) as used by mIoU.forward
 * (compute_mIoU.py:16-27): cm[n][t'][p] += 1 with p = argmax_c logits (first index on ties),
 * t' = t if 0 <= t < C else C (void row). TP/FP/FN follow from cm (void rows count as FP).
 *   pred_kind 0: `pred` = logits [N][C][HW] of `dtype`         (argmax taken in the kernel)
 *   pred_kind 1: `pred` = uint8  class map [N][HW]
 *   pred_kind 2: `pred` = int64  class map [N][HW]
 * cm: int64 [N][C+1][C]; accumulate != 0 adds to the existing content, else cm is overwritten.
 * ---------------------------------------------------------------------------------------------- */
int eeseg_confusion_hist(const void* pred, int pred_kind, int dtype, const int64_t* targets,
                         int N, int C, int64_t HW, int64_t* cm, int accumulate, void* stream);

/* End-of-batch accumulation of the entropy-gated evaluator (eval_br_ent.py:57-70), one launch:
 * image n took exit e = exit_idx[n] (a still-active -1 becomes the final exit E-1, written back); the
 * argmax map amax_all[e][n] (uint8 [E][N][HW]) is histogrammed against targets and added to
 * cm_acc[e] and to the global slot cm_acc[E] (int64 [E+1][C+1][C]); counts[e] and counts[E] (int64 [E+1])
 * are incremented; pred (optional uint8 [N][HW]) receives the map of the exit taken. */
int eeseg_exit_accumulate(const uint8_t* amax_all, const int64_t* targets, int32_t* exit_idx, int E, int N,
                          int C, int64_t HW, int64_t* cm_acc, int64_t* counts, uint8_t* pred, void* stream);
/* The same with uint8 labels [N][HW] (classes 0..C-1, any value >= C is void — VOC's 255 or the reference's C;
 * C <= 255): the label map crosses PCIe at 1 byte per pixel instead of 8 and is widened inside the kernel. */
int eeseg_exit_accumulate_u8(const uint8_t* amax_all, const uint8_t* targets, int32_t* exit_idx, int E, int N,
                             int C, int64_t HW, int64_t* cm_acc, int64_t* counts, uint8_t* pred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exit gate, stage 1 (per pixel).
 * Replaces, fused: F.interpolate(bilinear, align_corners=False) (from_deepv3_new.py:149,152),
 * F.softmax(.,1) (eval_br_ent.py:58, ee_dnn_op_ne.py:80), scipy entropy base C
 * (eval_br_ent.py:29), argmax (ee_dnn_op_ne.py:82,100), and the per-pixel threshold.
 *
 *   in:  logits of `in_dtype`, element strides (in_sn, in_sc, in_sy, in_sx), spatial size h x w.
 *        When (h,w) == (H,W) no interpolation is done (standalone gate on full-res tensors).
 *   in_kind 0 = logits (softmax inside), 1 = probabilities (re-normalised p/sum p like scipy).
 *   out (each optional, NULL to skip):
 *        up_logits  [N][C][H][W] of up_dtype with image stride up_sn elements (C planes of H*W)
 *        ent        f32 [N][H][W]   normalised entropy  -sum p ln p / ln C
 *        amax       u8  [N][H][W]   argmax class (first index on ties)
 *        mask       u8  [N][H][W]   1 where ent < tau
 *        part_sum   f64 [N][eeseg_exit_gate_num_partials(H,W)]  per-block entropy sums (ordered)
 *        part_cnt   i32 [N][same]   per-block count of pixels with ent < tau
 * ---------------------------------------------------------------------------------------------- */
int eeseg_exit_gate_num_partials(int H, int W);
int eeseg_exit_gate_pixels(const void* in, int in_dtype, int in_kind,
                           int64_t in_sn, int64_t in_sc, int64_t in_sy, int64_t in_sx,
                           int N, int C, int h, int w, int H, int W, float tau,
                           void* up_logits, int up_dtype, int64_t up_sn,
                           float* ent, uint8_t* amax, uint8_t* mask,
                           double* part_sum, int32_t* part_cnt, void* stream);

/* Block max/min pooling of an entropy map followed by the mean: img_norm_entropy with s != 1
 * (eval_br_ent.py:33-35; skimage block_reduce pads the END of each axis with 0).
 * ent f32 [N][H][W] -> score f32 [N]. mode 0 = max, 1 = min. */
int eeseg_entropy_pool_mean(const float* ent, int N, int H, int W, int s, int mode,
                            float* score, void* stream);

/* Exit gate, stage 2 (per image): the decision rule of br_evaluator (eval_br_ent.py:57-64) and
 * eval_ee_deeplabv3.__call__ (ee_dnn_op_ne.py:80-87).
 *   score[n] = sum(part_sum[n][:]) / (H*W)       (or taken from score_in when part_sum == NULL)
 *   image n leaves at exit `exit_id` iff exit_idx[n] < 0 (still active) and
 *       less_than ? score < tau : score > tau
 *   exit_idx  i32 [N]  in/out (-1 = still active)       score_out f32 [N] (optional)
 *   exited_px i64 [N]  out, sum(part_cnt) (optional)
 *   active_list i32 [N] + active_count i32 [1]: compacted ascending list of images still active
 *   after this exit (optional) — the batch the next backbone section has to run. */
int eeseg_exit_gate_decide(const double* part_sum, const int32_t* part_cnt, int num_partials,
                           const float* score_in, int N, int64_t HW, float tau, int less_than,
                           int exit_id, int32_t* exit_idx, float* score_out, int64_t* exited_px,
                           int32_t* active_list, int32_t* active_count, void* stream);

/* Stage commit of the compute-skipping engine — what follows a gate for the n images still in flight, in one launch
 * (decision rule of eval_br_ent.py:57-64 / ee_dnn_op_ne.py:80-87):
 *   image j (batch position positions[j]) leaves at `exit_id` iff take_all, or less_than ? score[j] < tau : score[j] > tau;
 *   scores_row[positions[j]] = score[j] (optional);  exit_idx[positions[j]] = exit_id or -1;
 *   pred[positions[j]][:] = amax[j][:] (u8 [HW]) for the images that leave;
 *   active_list / active_count: ascending j of the survivors;  next_positions[k] = positions[active_list[k]];
 *   *exited_px_acc += sum(exited_px_in[0..n)) (both optional). */
int eeseg_exit_stage_commit(const float* score, float tau, int less_than, int exit_id, int take_all,
                            const int64_t* positions, const uint8_t* amax, int n, int64_t HW, float* scores_row,
                            int32_t* exit_idx, uint8_t* pred, const int64_t* exited_px_in, int64_t* exited_px_acc,
                            int32_t* active_list, int32_t* active_count, int64_t* next_positions, void* stream);

/* Batch compaction after a gate (the compute-skipping form of ee_dnn_op_ne.py:80-101, which always runs the tail):
 *   dst[j][:] = src[active_list[j]][:]  for j < *active_count (j < n_dst when active_count == NULL)
 * rows of row_bytes bytes (16-byte multiple, 16-byte aligned): the activations of the images still active after an
 * exit, moved to the front of the next backbone section's input. active_list / active_count as written by
 * eeseg_exit_gate_decide; list entries outside [0, n_src) are skipped. */
int eeseg_compact_rows(const void* src, void* dst, const int32_t* active_list, const int32_t* active_count,
                       int n_src, int n_dst, int64_t row_bytes, void* stream);

/* Plain bilinear up-sampling of E stacked low-res logit tensors into [.. ][C][H][W] planes
 * (the reference's forward return value, from_deepv3_new.py:149-155, without the cat copy). */
int eeseg_upsample_bilinear(const void* in, int in_dtype,
                            int64_t in_sn, int64_t in_sc, int64_t in_sy, int64_t in_sx,
                            int N, int C, int h, int w, int H, int W,
                            void* out, int out_dtype, int64_t out_sn, void* stream);
/* Adjoint of eeseg_upsample_bilinear (the backward of F.interpolate(..., mode='bilinear', align_corners=False)
 * at from_deepv3_new.py:149,152 in training): dlow[p][y][x] = sum_{Y,X} wy(Y,y) wx(X,x) dout[p][Y][X] over
 * `planes` = N*C contiguous planes. Gather form, fixed summation order (bit-reproducible; ATen scatters with
 * atomics). dout fp32 / bf16 [planes][H][W]; dlow fp32 [planes][h][w]. */
int eeseg_upsample_bilinear_bwd(const void* dout, int dtype, int64_t planes, int h, int w, int H, int W,
                                float* dlow, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-exit pixelwise cross-entropy.
 * Replaces BrXEntropyLoss.forward / _cross_entropy._compute_loss (my_pixelwise_xentropy.py:11-14,
 * 30-46): per exit e, L_e = mean over valid pixels of -log softmax(y_e)[t].
 *   logits [E][N][C][HW] (exit stride exit_stride elements), targets int64 [N][HW]
 *   coef   f32 [E] (optional): d(total)/dL_e the caller will apply; used only for dlogits
 *   per_exit f32 [E] out; valid_count i64 [1] out
 *   dlogits (optional, same dtype/layout as logits): coef[e]/valid * (softmax - onehot), 0 at void
 *   workspace: eeseg_multi_exit_ce_workspace_bytes(E,N,HW) bytes of scratch
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_multi_exit_ce_workspace_bytes(int E, int N, int64_t HW);
int eeseg_multi_exit_ce_fwd(const void* logits, int dtype, int64_t exit_stride,
                            const int64_t* targets, int E, int N, int C, int64_t HW,
                            int64_t ignore_index, const float* coef, float* per_exit,
                            int64_t* valid_count, void* dlogits, void* workspace, void* stream);
/* Unfused backward (reads logits again): dlogits = g[e]/valid * (softmax - onehot). */
int eeseg_multi_exit_ce_bwd(const void* logits, int dtype, int64_t exit_stride,
                            const int64_t* targets, int E, int N, int C, int64_t HW,
                            int64_t ignore_index, const float* g, const int64_t* valid_count,
                            void* dlogits, void* stream);
/* dlogits[e] *= g[e]/coef[e], skipped on the device when every ratio is exactly 1. */
int eeseg_scale_exits(void* dlogits, int dtype, int64_t exit_stride, int E, int64_t per_exit_elems,
                      const float* g, const float* coef, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Lovasz-softmax (multi-exit).
 * Replaces lovasz_softmax / lovasz_softmax_flat / flatten_probas / lovasz_grad
 * (lovaszsoftmax.py:154-219, 19-31) as looped by BSL.LovaszSoftmax.forward
 * (branchy_seg_losses.py:151-159).
 *   probas [E][N][C][HW] (raw logits on the reference path), labels int64 [N][HW]
 *   ignore: label value to drop (has_ignore != 0), classes_mode 0 = 'present', 1 = 'all'
 *   per_image != 0: loss per image then mean over images
 *   per_exit f32 [E] out
 *   dprobas (optional): d per_exit[e] / d probas[e]  (caller scales by its weights)
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_lovasz_workspace_bytes(int E, int N, int C, int64_t HW);
int eeseg_lovasz_fwd_bwd(const void* probas, int dtype, int64_t exit_stride, const int64_t* labels,
                         int E, int N, int C, int64_t HW, int has_ignore, int64_t ignore,
                         int classes_mode, int per_image, float* per_exit, void* dprobas,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (tcgen05 / TMEM / TMA), inference form with folded BatchNorm.
 * Replaces the dense contraction of torchvision DeepLabHead / ASPP / ASPPConv as called from
 * from_deepv3_new.py:34,131,147,151 (branches[i](X), classifier(X)).
 * The same kernel runs the ResNet Bottleneck convolutions of the backbone sections
 * (base_model[i], from_deepv3_new.py:146,151) with the residual add fused.
 *   x   bf16 NHWC [N][hin][win][Cin]       (Cin % 64 == 0)
 *   wt  bf16 [Cout][R][S][Cin]             (Cout % 16 == 0; tiled by <= 256 output channels)
 *   y = act( scale[co] * conv(x, wt; dilation, stride, pad) + shift[n?][co] + residual )
 *       tap (r,s) of output (y,x) reads input (y*stride + r*dilation - pad, x*stride + s*dilation - pad),
 *       zero outside the image (the pad applies along each kernel dimension that has more than one tap);
 *       pad < 0 = 'same' padding dilation*(R/2) of an odd square kernel;
 *       output spatial size (hin-1)/stride+1 x (win-1)/stride+1
 *   shift_sn: image stride of `shift` in elements (0 = shared by all images)
 *   residual: optional bf16 NHWC tensor of the OUTPUT shape (pixel stride ldr), or NULL. It is added
 *   inside the accumulator (identity K blocks on the tensor core), i.e. BEFORE the scale: pass
 *   weights with the scale folded in and scale == 1 (Cout %% 64 == 0)
 *   scale: NULL = 1 for every channel (weights with the scale folded in): the epilogue then only adds the shift and
 *   reads half as much shared memory per output column (a group passes NULL for all of its problems or for none)
 *   out NHWC with pixel stride ldo elements, written at channel offset already applied to `out`
 *   out_dtype EESEG_BF16 or EESEG_F32; relu != 0 applies max(.,0) last
 * ---------------------------------------------------------------------------------------------- */
int eeseg_conv_igemm_fwd(const void* x, const void* wt, const float* scale, const float* shift,
                         int64_t shift_sn, int N, int hin, int win, int Cin, int Cout, int R, int S,
                         int dilation, int stride, int pad, int relu, const void* residual, int64_t ldr,
                         void* out, int out_dtype, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training: gradients of the stride-1 'same' convolutions of the exit heads (the backward of the
 * DeepLabHead / ASPP convolutions under train_epoch, train_funcs.py:22-27; the reference gets them from
 * autograd through cuDNN).
 *
 * eeseg_conv_igemm_wgrad:  dW[co][r][s][ci] = sum_{n,y,x} dY[n,y,x,co] * X[n, y+r*dil-pad, x+s*dil-pad, ci]
 *   tcgen05 implicit GEMM with the pixels as the contraction dimension (both operands MN-major, TMA boxes of
 *   NHWC tensors; the tap is a coordinate offset, zero fill = padding). pad = dil * (R / 2) along every kernel
 *   dimension with more than one tap ('same' for odd kernels; R = 4, S = 1 is the space-to-depth'ed ResNet stem,
 *   whose forward runs with pad = 2).
 *   x  bf16 NHWC [N][h][w][Cin] (Cin % 64 == 0);  dy bf16 NHWC, pixel stride ldy, dy_channels channels in
 *   total, this convolution's Cout (% 64 == 0) channels starting at co_off;  dw fp32 [Cout][R][S][Cin]
 *   (overwritten). When the pixel dimension is split over CTAs to fill the GPU every split writes its own partial
 *   dW into `workspace` (eeseg_conv_igemm_wgrad_workspace_bytes bytes, 16 B aligned) and a second kernel sums them
 *   in a fixed order: bit-reproducible, no atomics.
 *
 * eeseg_conv_igemm_dgrad:  dX = conv(dY, W'), W'[ci][r][s][co] = W[co][R-1-r][S-1-s][ci], on the forward
 *   kernel. dy bf16 NHWC contiguous [N][h][w][Cout] (Cout % 64 == 0), wt bf16 [Cout][R][S][Cin]
 *   (Cin % 16 == 0), dx bf16 / fp32 NHWC with pixel stride lddx; workspace of
 *   eeseg_conv_igemm_dgrad_workspace_bytes bytes (the transformed weights), 256-byte aligned.
 * eeseg_conv_weight_rot180_t: the weight transform alone ([Cout][R][S][Cin] -> [Cin][R][S][Cout], taps
 *   rotated by 180 degrees), for callers that cache it.
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_conv_igemm_wgrad_workspace_bytes(int N, int h, int w, int Cin, int Cout, int R, int S);
int eeseg_conv_igemm_wgrad(const void* x, const void* dy, int64_t ldy, int dy_channels, int co_off, int N, int h,
                           int w, int Cin, int Cout, int R, int S, int dilation, float* dw, void* workspace,
                           void* stream);
/* The weight gradient written straight into the PARAMETER's gradient tensor: grad fp32 [Cout][Cin][R][S] (the layout of
 * nn.Conv2d.weight.grad), overwritten or, with accumulate != 0, added to — what autograd's AccumulateGrad does with one
 * more strided read-modify-write launch per layer (`grad.add_(dW.permute(0, 3, 1, 2))`). The partial sums always go
 * through `workspace` (eeseg_conv_igemm_wgrad_to_param_workspace_bytes bytes) and are reduced in a fixed order. */
size_t eeseg_conv_igemm_wgrad_to_param_workspace_bytes(int N, int h, int w, int Cin, int Cout, int R, int S);
int eeseg_conv_igemm_wgrad_to_param(const void* x, const void* dy, int64_t ldy, int dy_channels, int co_off, int N, int h,
                                    int w, int Cin, int Cout, int R, int S, int dilation, float* grad, int accumulate,
                                    void* workspace, void* stream);
size_t eeseg_conv_igemm_dgrad_workspace_bytes(int Cin, int Cout, int R, int S);
int eeseg_conv_igemm_dgrad(const void* dy, const void* wt, int N, int h, int w, int Cin, int Cout, int R, int S,
                           int dilation, void* dx, int dx_dtype, int64_t lddx, void* workspace, void* stream);
int eeseg_conv_weight_rot180_t(const void* w, int Cout, int R, int S, int Cin, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Soft-overlap sums of the branchy Dice / Jaccard losses (BSL.DiceLoss / BSL.JaccardLoss._compute_loss,
 * branchy_seg_losses.py:40-77, under BrSegLoss.forward :24-38) and their gradient.
 *   logits [E][N][C][HW] (f32 / bf16, exit stride in elements), targets int64 [N][HW] (values outside [0,C) match no class)
 *   fwd: sums f32 [E][N][3][C] = { sum_px p_c [t==c],  sum_px p_c,  sum_px [t==c] },  p = softmax over C
 *   bwd: dlogits[e,n,c,px] = p_c * (g_c - sum_k p_k g_k),  g_c = dsum_p[e,n,c] + dsum_pt[e,n,c] [t==c]
 *        (dsum_pt, dsum_p f32 [E][N][C]: the loss formula's derivatives with respect to the first two sums)
 *   workspace: eeseg_soft_overlap_workspace_bytes bytes. Fixed summation order (bit-reproducible).
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_soft_overlap_workspace_bytes(int E, int N, int C, int64_t HW);
int eeseg_soft_overlap_fwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N,
                           int C, int64_t HW, float* sums, void* workspace, void* stream);
int eeseg_soft_overlap_bwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N,
                           int C, int64_t HW, const float* dsum_pt, const float* dsum_p, void* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Branchy focal loss (BSL.FocalLoss._compute_loss, branchy_seg_losses.py:122-131, under BrSegLoss.forward :24-38):
 *   loss[e,n,px] = -alpha[t] * w[e,n,px] * (1 - p_t)^gamma * log p_t,  p = softmax(logits[e,n,:,px]),  t = targets[n,px] in [0,C)
 *   w = pixel_weight[e*pw_exit_stride + n*pw_image_stride + px] (optional f32; stride 0 = broadcast): the reference's
 *   broadcast `loss * alpha[targets]` product (:128-129) and the upstream gradient map of reduction='none';
 *   per_exit_sum f32 [E] = sum over (n,px) of loss (ordered fp64 partials);  loss_map (optional) f32 [E][N][HW];
 *   dlogits (optional, logits' dtype) = coef[e] * d loss / d logits  (coef folds the reduction and the exit weight);
 *   alpha (optional) f32 [C];  workspace: eeseg_focal_workspace_bytes bytes.
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_focal_workspace_bytes(int E, int N, int64_t HW);
int eeseg_focal_fwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N, int C,
                    int64_t HW, float gamma, const float* alpha, const float* pixel_weight, int64_t pw_exit_stride,
                    int64_t pw_image_stride, const float* coef, float* per_exit_sum, float* loss_map, void* dlogits,
                    void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-mode BatchNorm2d (+ residual add) (+ ReLU) on NHWC bf16 activations — the nn.BatchNorm2d / ReLU /
 * `out += identity` modules of DeepLabHead, ASPP and the ResNet Bottlenecks as they run under net.train() in
 * train_epoch (train_funcs.py:12-33).  x, y, residual, dy, dx, dres: bf16 [P = N*h*w][C], C % 64 == 0.
 *   fwd:  mean/var over the P pixels per channel (biased variance inside the normalisation),
 *         y = act(gamma*(x-mean)*invstd + beta (+ residual)), act = ReLU if relu else identity;
 *         running_mean/var (optional) <- (1-momentum)*running + momentum*(mean / unbiased variance);
 *         save_mean, save_invstd fp32 [C] for the backward.
 *   bwd:  dy' = dy * [y > 0] (relu; needs y),  dbeta = sum dy',  dgamma = sum dy'*xhat,
 *         dx = gamma*invstd*(dy' - dbeta/P - xhat*dgamma/P),  dres (optional) = dy'.
 *   Fixed summation order (bit-reproducible). workspace: eeseg_bn_train_workspace_bytes(C) bytes, 16 B aligned.
 * ---------------------------------------------------------------------------------------------- */
size_t eeseg_bn_train_workspace_bytes(int C);
int eeseg_bn_train_fwd(const void* x, int64_t P, int C, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int relu, const void* residual, void* y,
                       float* save_mean, float* save_invstd, void* workspace, void* stream);
int eeseg_bn_train_bwd(const void* dy, const void* x, const void* y, int64_t P, int C, const float* gamma,
                       const float* save_mean, const float* save_invstd, int relu, void* dx, void* dres,
                       float* dgamma, float* dbeta, void* workspace, void* stream);
/* The same with dgamma / dbeta ADDED to gamma_grad / beta_grad, the parameters' own .grad tensors (no separate
 * accumulation launch per BatchNorm). */
int eeseg_bn_train_bwd_acc(const void* dy, const void* x, const void* y, int64_t P, int C, const float* gamma,
                           const float* save_mean, const float* save_invstd, int relu, void* dx, void* dres,
                           float* gamma_grad, float* beta_grad, void* workspace, void* stream);

/* ResNet stem helpers (base_model[0][0:4], torchvision resnet.py conv1/bn1/relu/maxpool):
 * space-to-depth (2x2) plus horizontal tap unrolling of the fp32 NCHW image, so that the 7x7 / stride-2 /
 * pad-3 convolution becomes a 4x1 / stride-1 / pad-2 implicit GEMM with K = 4 taps x 64 channels:
 * out bf16 NHWC [N][(H+1)/2][(W+1)/2][64], channel u*12 + (a*2+b)*3 + c = x[n][c][2Y+a][2(X+u-2)+b]
 * for u in 0..3 (zero outside the image and for channels >= 48). */
int eeseg_stem_space_to_depth(const float* x, int N, int H, int W, void* out, void* stream);
/* The same for an image tensor of dtype x_kind: EESEG_F32, EESEG_BF16 (what a bf16 inference stream uploads: the
 * kernel rounds fp32 images to bf16 as its first step anyway, so the result is bit-identical at half the host->device
 * bytes) or EESEG_U8 (raw 0..255 pixels, a quarter of the bytes): for uint8 the kernel applies what the reference's loader does
 * on the CPU (get_seg_datasets.py:62-70, ToTensor + Normalize): v = (u / 255 - mean[c]) / std[c], mean / std HOST
 * arrays of 3 floats (NULL: mean 0, std 1). */
int eeseg_stem_space_to_depth_any(const void* x, int x_kind, const float* mean, const float* std_, int N, int H, int W,
                                  void* out, void* stream);
/* 3x3 / stride-2 / pad-1 max pooling of a bf16 NHWC tensor (C % 8 == 0). */
int eeseg_maxpool3x3s2_nhwc(const void* x, int N, int h, int w, int C, void* out, void* stream);

/* Training max-pool 3x3 / stride 2 / pad 1 (torchvision resnet `maxpool` under net.train(), from_deepv3_new.py:146):
 * forward also writes idx u8 [N][ho][wo][C], the winning window tap 0..8 (row-major; first maximum, NaN wins — ATen's
 * rule); backward gathers dx[n][y][x][c] = sum of dout over the <= 2x2 windows whose idx points at (y,x): fixed
 * order, no atomics. x / dx bf16 NHWC [N][h][w][C], out / dout [N][ho][wo][C], C % 8 == 0. */
int eeseg_maxpool3x3s2_nhwc_train(const void* x, int N, int h, int w, int C, void* out, void* idx, void* stream);
int eeseg_maxpool3x3s2_nhwc_bwd(const void* dout, const void* idx, int N, int h, int w, int C, void* dx, void* stream);

/* Grouped form: `nprob` (<= 4) stride-1 'same' convolutions of the SAME input (the ASPP branches:
 * 1x1 and the three atrous 3x3) as ONE persistent launch. Their tiles differ a lot in cost — a tile near
 * the border of a 65x65 map keeps 4 of the 9 taps of a d=24 conv, an interior one all 9 — so separate
 * launches (one wave of 144 tiles each) wait for their slowest tile; the grouped launch walks one
 * work list over all problems, longest items first.
 *   wt/scale/shift/ksize/dilation/ch_off: HOST arrays of length nprob (device pointers / ints);
 *   problem g writes Cout channels at channel ch_off[g] of `out` (bf16 NHWC, pixel stride ldo,
 *   out_channels channels in total);
 *   schedule: DEVICE int32[n_items], item = g << 24 | tile (tile = m_tile * (Cout/BN) + n_tile in the
 *   geometry eeseg_conv_group_tiles reports), every (g, tile) exactly once, ordered by decreasing cost;
 *   cta_pairs != 0: the list is a PAIR list — entries 2i and 2i+1 are the two spatial tiles (same g, same channel
 *   tile) that one cluster of two CTAs runs as a single 256-row tcgen05.mma.cta_group::2 tile, each CTA staging its own
 *   activation tile and half of the weight tile; n_items counts entries (even). eeseg_conv_pair_clusters() = number of
 *   such clusters resident at once (the round size of a pair list). */
int eeseg_conv_group_tiles(int hin, int win, int Cout, int* tiles_x, int* tiles_y, int* bw, int* bh, int* bn);
int eeseg_conv_igemm_grouped(const void* x, int nprob, const void* const* wt, const float* const* scale,
                             const float* const* shift, const int* ksize, const int* dilation,
                             const int* ch_off, int N, int hin, int win, int Cin, int Cout, int relu,
                             void* out, int64_t ldo, int out_channels, const int32_t* schedule, int n_items,
                             int cta_pairs, void* stream);
int eeseg_conv_pair_clusters(void);

/* Programmatic dependent launch (the prologue of conv launch i+1 overlaps the tail of launch i; the kernel executes
 * griddepcontrol.wait before touching its inputs) is on unless the environment holds EESEG_CONV_PDL=0 when the library
 * is first used. The library keeps no mutable process-wide state: every entry point is re-entrant per stream; the
 * cycle-counter / %globaltimer instrumentation used while tuning exists only in -DEESEG_TUNING builds
 * (include/eeseg_tuning.h). */

/* Small fp32 dense layer y[n][o] = act((x[n] . W[o]) * scale[o] + shift[o]) (scale/shift optional) for the
 * ASPP pooled branch (torchvision deeplabv3.py:70-83) and its share of the ASPP projection. */
int eeseg_dense_bn_act(const float* x, const float* W, const float* scale, const float* shift, int N, int K,
                       int O, int relu, float* y, void* stream);

/* Global average pool of an NHWC bf16 tensor: [N][h][w][C] -> f32 [N][C] (ASPPPooling's
 * AdaptiveAvgPool2d(1), torchvision deeplabv3.py:70-83). Two ordered stages (deterministic);
 * workspace: eeseg_global_avgpool_workspace_bytes(N, C) bytes of scratch. */
size_t eeseg_global_avgpool_workspace_bytes(int N, int C);
int eeseg_global_avgpool_nhwc(const void* x, int N, int64_t hw, int C, float* out, void* workspace,
                              void* stream);

/* ---- training-step helpers (csrc/train_misc.cu) ---------------------------------------------------------------
 * Dropout of ASPP.project (torchvision deeplabv3.py:101, active under net.train(): from_deepv3_new.py:147 inside
 * train_funcs.py:22): y = x * keep / (1 - p) on bf16 tensors of n elements (n % 8 == 0), keep ~ Bernoulli(1 - p) from a
 * counter-based generator; rng_state: DEVICE uint64[2] = {seed, offset}, the offset advances by n/8 per call ON THE
 * DEVICE, so a captured CUDA graph draws a fresh mask on every replay. mask: uint8[n/8] (bit k = element 8i+k kept),
 * consumed by the backward dx = dy * keep / (1 - p). p is realised in 1/256 steps (0.5 exactly). */
int eeseg_dropout_fwd(const void* x, int64_t n, float p, void* rng_state, void* y, void* mask, void* stream);
int eeseg_dropout_bwd(const void* dy, const void* mask, int64_t n, float p, void* dx, void* stream);

/* torch.optim.SGD's update (deepv3_funcs.py:74-101: momentum 0.9, weight decay 5e-4, one learning rate per parameter
 * group) for MANY tensors in one launch: d = g + wd * p; buf = momentum * buf + d; p -= lr[group] * buf. `chunks`: DEVICE
 * array of n_chunks records {float* p; const float* g; float* buf; int32 count; int32 group} (eeseg_sgd_chunk_bytes()
 * bytes each; one thread block per record); lrs: DEVICE float array indexed by group — read at run time, so a
 * learning-rate schedule does not invalidate a captured graph. Zero-initialised momentum buffers give torch's first step. */
size_t eeseg_sgd_chunk_bytes(void);
int eeseg_sgd_multi(const void* chunks, int n_chunks, const float* lrs, float momentum, float weight_decay, void* stream);

/* Backward of the small dense layer y[n][o] = x[n] . W[o] (eeseg_dense_bn_act without scale/shift; the pooled ASPP
 * branch's 1x1 convolution on one pixel per image): dW[o][k] = sum_n dy[n][o] x[n][k], dx[n][k] = sum_o dy[n][o] W[o][k]
 * (either output may be NULL). fp32, fixed summation order. */
int eeseg_dense_bwd(const float* dy, const float* x, const float* W, int N, int K, int O, float* dW, float* dx, void* stream);

/* Per-step weight preparation of MANY convolutions in one launch: for every table record {const float* src
 * [Cout][Cin][RS] (the nn.Conv2d parameter); bf16* krsc [Cout][RS][Cin]; bf16* rot [Cin][RS][Cout] with the taps reversed;
 * int32 Cout, Cin, RS, co0, ci0, pad} (eeseg_weight_prep_tile_bytes() bytes, DEVICE array) one thread block converts the
 * 32 x 32 (co, ci) tile at (co0, ci0): what the forward / weight-gradient kernels (krsc) and the input-gradient kernel
 * (rot, see eeseg_conv_weight_rot180_t) read. Cout, Cin multiples of 32; max_rs = the largest RS in the table. */
size_t eeseg_weight_prep_tile_bytes(void);
int eeseg_weight_prep_multi(const void* tiles, int n_tiles, int max_rs, void* stream);

/* BatchNorm (batch statistics over the N rows, biased variance for the normalisation, unbiased for the running estimate,
 * as nn.BatchNorm2d on an [N,C,1,1] tensor) + optional ReLU on fp32 row vectors [N][C] — the pooled ASPP branch
 * (deeplabv3.py:70-83); forward saves mean / invstd for the backward, which returns dx, dgamma, dbeta. */
int eeseg_bn_rows_fwd(const float* x, int N, int C, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, int relu, float* y, float* save_mean,
                      float* save_invstd, void* stream);
int eeseg_bn_rows_bwd(const float* dy, const float* x, const float* y, int N, int C, const float* gamma,
                      const float* save_mean, const float* save_invstd, int relu, float* dx, float* dgamma, float* dbeta,
                      void* stream);

/* out[n][p][c] = bf16(v[n][c] * scale) for p < hw: a per-image vector broadcast over the map (ASPPPooling's up-sampling of
 * a 1x1 map, deeplabv3.py:83; with scale = 1/hw the backward of the global average pool). C % 8 == 0. */
int eeseg_broadcast_rows_nhwc(const float* v, int N, int64_t hw, int C, float scale, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EESEG_H_ */
