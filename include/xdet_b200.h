/* xdet_b200.h -- C-ABI of the B200-native Light-Head R-CNN hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, no C++/torch/TF
 * types.  Every entry point names the reference interface it replaces (paths relative to
 * HiKapok/X-Detector @ 1b19e15).  All `d_` pointers are DEVICE pointers on the current CUDA
 * device; all calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 * legacy default stream) unless the name ends in `_host`.  The library never frees or keeps a
 * caller pointer.  Return value: 0 on success, negative XDET_E* on error, with a message
 * retrievable (per thread) from xdet_last_error().  Nothing here ever calls exit()
 * (the reference does on a launch failure: cpp/PSROIPooling/ps_roi_align_op.cu:150-155).
 */
#ifndef XDET_B200_H_
#define XDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XDET_OK 0
#define XDET_EINVAL (-1) /* shape / attribute violation (the reference's errors::InvalidArgument) */
#define XDET_ECUDA (-2)  /* CUDA runtime / launch error */
#define XDET_ENOMEM (-3)

/* Message of the last failing call made by this thread ("" if none). */
const char* xdet_last_error(void);
/* "xdet_b200 <version> sm_100a" */
const char* xdet_version(void);
/* Number of kernels launched by this library in this process (bench.py's gpu_launches). */
long long xdet_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * PsRoiAlign forward.
 * Replaces: TF op "PsRoiAlign" -- REGISTER_OP cpp/PSROIPooling/ps_roi_align_op.cc:38-76,
 *           PSROIAlignOp::Compute :220-244, functor seam PSROIAlignFunctor<Device,T>::operator()
 *           cpp/PSROIPooling/ps_roi_align_op.h:40-57 (same flat pointers + dims tuple),
 *           CUDA kernel cpp/PSROIPooling/ps_roi_align_op.cu:37-156.
 *   d_inputs  [N,C,H,W] f32 (NCHW, as the reference requires)
 *   d_rois    [N,R,4]   f32 (center_y, center_x, h, w), normalised to [0,1]
 *   d_pooled  [N,R,gw*gh,C/(gw*gh)] f32   (out)
 *   d_index   [N,R,gw*gh,C/(gw*gh)] i32   (out; arg-max sample id for 'max', 0 for 'mean')
 *   use_max   1 = pool_method contains "max", 0 = "mean"
 * Errors (XDET_EINVAL): gw,gh <= 0, C % (gw*gh) != 0 (shape fn :68-70), negative sizes.
 * Results are bit-identical to the reference's CPU functor (:94-193); the two documented
 * deviations (degenerate RoIs write index 0; out-of-contract samples are clamped) are listed
 * in DESIGN.md.
 */
int xdet_psroi_align_fwd(const float* d_inputs, const float* d_rois, float* d_pooled, int32_t* d_index,
                         int N, int C, int H, int W, int R, int gw, int gh, int use_max, void* stream);

/* Same, with an explicit kernel variant (tests and benchmarks):
 *   XDET_PSROI_AUTO   pick by shape
 *   XDET_PSROI_GATHER one thread per output element, taps gathered from global memory (any shape)
 *   XDET_PSROI_PLANES channel-slice planes staged in shared memory, RoIs streamed past them
 *                     (needs (C/(gw*gh))*H*W*4 B <= ~200 KB)
 *   XDET_PSROI_SELECT PLANES staging; max pooling only: an fp32 pass selects the arg-max sample when that is provable
 *                     (else the exact loop runs), the fp64 blend is evaluated for the selected sample only */
#define XDET_PSROI_AUTO 0
#define XDET_PSROI_GATHER 1
#define XDET_PSROI_PLANES 2
#define XDET_PSROI_SELECT 3
int xdet_psroi_align_fwd_ex(const float* d_inputs, const float* d_rois, float* d_pooled, int32_t* d_index,
                            int N, int C, int H, int W, int R, int gw, int gh, int use_max, int variant,
                            void* stream);

/* PsRoiAlign backward.
 * Replaces: TF op "PsRoiAlignGrad" -- cpp/PSROIPooling/ps_roi_align_grad_op.cc:39-57,326-373,
 *           CPU functor :187-322, CUDA kernel ps_roi_align_grad_op.cu:37-165; registered as the
 *           gradient of PsRoiAlign at light_head_rfcn_train.py:201-213.
 *   d_rois [N,R,4], d_pooled_grad [N,R,G,C/G] f32, d_index [N,R,G,C/G] i32 -> d_grad [N,C,H,W] f32
 * d_grad is fully overwritten (the reference zero-fills first, :197).  Deterministic: each map
 * cell accumulates its contributions in the reference CPU functor's order (roi-major), so the
 * result is bit-identical to it and run-to-run reproducible (the reference's GPU kernel uses
 * float atomics and is not).
 */
int xdet_psroi_align_bwd(const float* d_rois, const float* d_pooled_grad, const int32_t* d_index, float* d_grad,
                         int N, int C, int H, int W, int R, int gw, int gh, int use_max, void* stream);

/* Host-buffer convenience forms (what a framework without device tensors would bind):
 * copy in, run, copy out, synchronise.  h_ pointers are host memory (pinned or pageable). */
int xdet_psroi_align_fwd_host(const float* h_inputs, const float* h_rois, float* h_pooled, int32_t* h_index,
                              int N, int C, int H, int W, int R, int gw, int gh, int use_max);
int xdet_psroi_align_bwd_host(const float* h_rois, const float* h_pooled_grad, const int32_t* h_index,
                              float* h_grad, int N, int C, int H, int W, int R, int gw, int gh, int use_max);


/* ---------------------------------------------------------------------------------------
 * Convolution (stride 1 or 2, dilated) / GEMM on the tcgen05 tensor cores (bf16 operands, fp32 accumulate).
 * Replaces: the cuDNN / cuBLAS kernels TensorFlow runs for tf.layers.conv2d / tf.layers.dense on
 *   the path -- conv2d_fixed_padding net/resnet_v2.py:89-100 and the 7x7/s2 stem :320-325, dilate_conv2d
 *   net/xdet_body.py:28-37, get_rpn net/xception_body.py:381-400, large_sep_kernel :450-475,
 *   the pointwise half of tf.layers.separable_conv2d :224-233, get_head's dense layers :540-558.
 * Input  : NHWC bf16, pixel (n,y,x) at d_in + ((n*H + y)*W + x)*in_cs, in_cs % 8 == 0.
 * Weights: [Cout][KH*KW][ceil(Cin/64)*64] bf16, tap-major, input channels zero-padded to 64.
 * Output : out(n,y,x,c) at out + n*out_sn + y*out_sy + x*out_sx + c*out_sc (elements), bf16 or fp32:
 *            v    = acc*scale[c] + bias[c]  (+ residual(n,y,x,c), bf16, laid out like `out`)  (ReLU if relu)
 *            out  = v ;  out2 = ReLU(v*scale2[c] + bias2[c])  (optional, bf16, laid out like `out`)
 *          with acc(n,y,x,c) = sum_{kh,kw,ci} in(n, y*stride_h + kh*dil_h - pad_top, x*stride_w + kw*dil_w - pad_left, ci) * w
 *          (reads outside the image are zero: TF 'SAME' padding = pad_top/left = floor(total/2); explicit
 *          fixed_padding = pad_top/left = (k-1)/2).
 * residual / out2 require a bf16 `out` with unit channel stride and pixel strides that are multiples of 8.
 * A plain GEMM D[M,Nc] = A[M,K] * W[Nc,K]^T is the case N=1, H=1, W=M, Cin=K, KH=KW=1.
 * fold_w != 0 (few-channel inputs such as the 3-channel image of the stem): the KW taps of a filter row and the
 *   in_cs (= 8) padded channels form ONE K chunk of KW*in_cs <= 64 contiguous elements.  The caller pads the rows
 *   horizontally: pixel (n,y,x) lives at d_in + ((n*H + y)*in_wp + x + pad_left)*in_cs with zeros in the padding
 *   and in_wp >= (Wout-1)*stride_w + 64/in_cs.  Weights are then [Cout][KH][64] with element kw*in_cs + ci.
 */
typedef struct {
  int N, H, W, Cin, in_cs;
  int Cout, KH, KW, dil_h, dil_w, pad_top, pad_left;
  int Hout, Wout;
  int stride_h, stride_w; /* 1 or 2 (0 = 1) */
  int fold_w, in_wp;      /* see above; in_wp is only read when fold_w != 0 */
  const void* weights;
  const float* scale; /* per-Cout multiplier (folded batch-norm), NULL = 1 */
  const float* bias;  /* per-Cout addend (conv bias / folded batch-norm), NULL = 0 */
  int relu;
  const void* residual; /* bf16 or NULL */
  void* out; /* may be NULL when out2 is given: only the second output is stored */
  int out_fp32; /* 0: bf16, 1: fp32 */
  long long out_sn, out_sy, out_sx, out_sc;
  void* out2; /* bf16 or NULL */
  const float* scale2;
  const float* bias2;
  int block_n;  /* 0 = auto; otherwise the N tile (multiple of 16, <= 256; of 64 for bf16 NHWC outputs) */
  int epi_groups; /* 0 = auto; 1 or 2 epilogue warpgroups (tuning / tests) */
  int max_ctas; /* 0 = one persistent CTA per SM; otherwise an upper bound (leaves SMs to a concurrent stream) */
  float* stats; /* NULL, or [2][Cout] fp32 the kernel ADDS the column sums and sums of squares of the stored bf16 `out`
                   to (training: the statistics of the batch-norm that reads `out`; Cout % 8 == 0, 16-byte aligned) */
} xdet_conv_desc;
int xdet_conv2d_bf16(const void* d_in, const xdet_conv_desc* desc, void* stream);
/* Convolutions are launched with programmatic stream serialization (their prologue overlaps the predecessor's tail;
 * the kernel executes griddepcontrol.wait before touching tensor data).  0 turns that off (debugging). */
void xdet_set_conv_pdl(int enabled);

/* Weight gradient of the same convolutions (training; replaces TensorFlow's Conv2DBackpropFilter behind
 * `optimizer.minimize`, light_head_rfcn_train.py:426-441):
 *   dw[co][kh*KW+kw][ci] += sum_{n,y,x} dy(n,y,x,co) * x(n, y*stride_h + kh*dil_h - pad_top, x*stride_w + kw*dil_w - pad_left, ci)
 * d_x  [N,H,W,in_cs] bf16 (the forward input), d_dy [N,Hout,Wout,dy_cs] bf16 (gradient of the forward output),
 * dw   [Cout][KH*KW][ceil(Cin/64)*64] fp32 -- the packed weight layout of xdet_conv2d_bf16 -- ACCUMULATED into
 *      (the caller zero-fills it; pixel splits of one tile meet through fp32 atomics, so the summation order and
 *      hence the last bits are not run-to-run reproducible).
 * The input gradient needs no entry point of its own: it is xdet_conv2d_bf16 on dy with the filter flipped and
 * its channel axes swapped (ops/conv.py:pack_dgrad_weight), on a zero-stuffed dy for stride-2 layers. */
typedef struct {
  int N, H, W, Cin, in_cs;
  int Cout, KH, KW, dil_h, dil_w, pad_top, pad_left, stride_h, stride_w;
  int Hout, Wout, dy_cs;
  float* dw;
  int splits; /* 0 = auto: enough pixel splits to fill the GPU */
  int fold_w, in_wp; /* as in xdet_conv_desc (the stem); dw is then [Cout][KH][64] with element kw*in_cs + ci */
} xdet_wgrad_desc;
int xdet_conv2d_wgrad_bf16(const void* d_x, const void* d_dy, const xdet_wgrad_desc* desc, void* stream);
/* The same weight gradient in the fp32-accurate "f16x2" precision (the training parity mode): x and dy are f16x2 planes
 * ([2][N,H,W,cs] fp16, `*_plane` elements apart, see xdet_conv2d_f16x2); dw += (sum dy*x) * out_scale, where out_scale
 * undoes a power-of-two scaling the caller applied to dy (loss scaling keeps gradients inside fp16's range). */
int xdet_conv2d_wgrad_f16x2(const void* d_x_pair, long long x_plane, const void* d_dy_pair, long long dy_plane,
                            const xdet_wgrad_desc* desc, float out_scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Bandwidth helpers around the tensor-core convolutions (bf16 NHWC tensors).
 * xdet_im2col_bf16      patch gather that turns the STRIDED convolutions of the ResNet-v2 stem / stage heads
 *                       (conv2d_fixed_padding with strides > 1: net/resnet_v2.py:62-100) into GEMMs:
 *                       dst[n,yo,xo, (kh*KW+kw)*C + c] = src(n, yo*stride+kh-pad_top, xo*stride+kw-pad_left, c)
 *                       (0 outside); src is NHWC bf16 (pixel pitch in_cs) or NCHW fp32 (the input image);
 *                       dst pixel pitch out_cs >= KH*KW*C, the tail is zero-filled.
 * xdet_maxpool3x3s2_bf16  tf.layers.max_pooling2d(3, 2, 'SAME') (net/resnet_v2.py:326-328); optional second
 *                       output ReLU(pooled*scale2 + bias2) = the first block's pre-activation (resnet_v2.py:163-164).
 * xdet_affine_relu_bf16 inference batch_norm (+ReLU): y = x*scale[c] + bias[c] (net/resnet_v2.py:41-50).
 * xdet_f32_to_bf16_rows [rows, cols] fp32 -> [rows, dst_pitch] bf16, zero tail (PsRoIAlign output -> dense operand).
 * xdet_image_to_nhwc8_bf16  the input image [N,C<=8,H,W] fp32 NCHW (what light_head_preprocess_for_eval delivers,
 *                       transposed by the model_fn) -> [N,H,Wp,8] bf16 with `pad_left` zero pixels in front of every
 *                       row, zero channels C..7 and zero pixels up to Wp: the layout xdet_conv2d_bf16's fold_w mode reads
 *                       (the 7x7/s2 stem, net/resnet_v2.py:320-325; block1_conv1 of XceptionBody, xception_body.py:243).
 */
int xdet_im2col_bf16(const void* d_src, int src_is_nchw_f32, void* d_dst, int N, int H, int W, int C, int in_cs,
                     int KH, int KW, int stride, int pad_top, int pad_left, int Ho, int Wo, int out_cs, void* stream);
int xdet_maxpool3x3s2_bf16(const void* d_src, void* d_dst, void* d_dst2, const float* d_scale2, const float* d_bias2,
                           int N, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left, void* stream);
/* same as xdet_maxpool3x3s2_bf16 plus tf.add(pooled, residual) (Xception entry flow, net/xception_body.py:283-289);
 * d_residual [N,Ho,Wo,C] bf16 or NULL. */
int xdet_maxpool3x3s2_add_bf16(const void* d_src, void* d_dst, void* d_dst2, const float* d_scale2, const float* d_bias2,
                               const void* d_residual, int N, int H, int W, int C, int Ho, int Wo, int pad_top,
                               int pad_left, void* stream);
/* Depthwise 3x3 'SAME' stride-1 convolution, depth multiplier 1, dilation 1 or 2: the depthwise half of
 * tf.layers.separable_conv2d (net/xception_body.py:224-233,264-272,351-376; its pointwise half is xdet_conv2d_bf16).
 * d_src/d_dst [N,H,W,C] bf16 NHWC (C % 8 == 0), d_weights [3*3][C] fp32 (TF depthwise_kernel [3,3,C,1] flattened),
 * relu_in != 0 applies the tf.nn.relu that precedes the layer in relu_separable_bn_block (:223) while loading. */
int xdet_depthwise3x3_bf16(const void* d_src, const float* d_weights, void* d_dst, int N, int H, int W, int C,
                           int dilation, int relu_in, void* stream);
/* dilation 1 runs the rolling-rows kernel (a thread walks down a strip of rows with the 3-row window in registers);
 * 0 selects the older one-row-per-thread kernel (A/B timing, tools/dw_one.py). */
void xdet_set_depthwise_rows(int enabled);
/* training-mode forward of the same pooling: also writes d_argmax [N,Ho,Wo,C] uint8 (see xdet_maxpool3x3s2_bwd_bf16) */
int xdet_maxpool3x3s2_argmax_bf16(const void* d_src, void* d_dst, void* d_argmax, int N, int H, int W, int C, int Ho,
                                  int Wo, int pad_top, int pad_left, void* stream);
int xdet_affine_relu_bf16(const void* d_src, void* d_dst, const float* d_scale, const float* d_bias, long long pixels,
                          int C, int relu, void* stream);
int xdet_f32_to_bf16_rows(const float* d_src, void* d_dst, long long rows, int cols, int dst_pitch, void* stream);
int xdet_image_to_nhwc8_bf16(const float* d_src, void* d_dst, int N, int C, int H, int W, int Wp, int pad_left,
                             void* stream);

/* ---------------------------------------------------------------------------------------
 * Detection post-processing.
 * Replaces: the '/device:CPU:0' block of bboxes_eval light_head_rfcn_eval.py:263-290 up to bboxes_nms_batch =
 *   eval_helper.tf_bboxes_select utility/eval_helper.py:556-625 (per class c >= 1: score and box times the 0/1 mask of
 *   score > select_threshold) -> bboxes_clip :365-404 (against bbox_img) -> filter_boxes :278-317 (both sides >
 *   min_size, centre inside (0,1)) -> bboxes_resize :423-447 -> bboxes_sort :333-361 (tf.nn.top_k, top_k = 2*nms_topk)
 *   -> bboxes_nms_batch :449-506 (tf.image.non_max_suppression, zero padded to keep_top_k).
 * The reference handles one image per call (its comment at :261); here N images and all num_classes-1 classes are
 * one batch of N*(num_classes-1) independent selections.
 *   d_probs [N,R,num_classes] softmax scores, d_boxes [N,R,4] decoded boxes (ymin,xmin,ymax,xmax),
 *   d_bbox_img [N,4], d_min_size [N] (= max(0.0001, min_size_ratio * sqrt(image_h*image_w / net_h*net_w)), :295)
 *   -> d_out_scores [N,num_classes-1,keep_top_k], d_out_boxes [N,num_classes-1,keep_top_k,4], class c at index c-1,
 *   descending score, zero padded.  Selection is bit-identical to the CPU restatement (oracle/detections.py).
 */
size_t xdet_det_postprocess_workspace_bytes(int N, int R, int num_classes, int top_k);
int xdet_det_postprocess(const float* d_probs, const float* d_boxes, const float* d_bbox_img, const float* d_min_size,
                         int N, int R, int num_classes, float select_threshold, int top_k, int keep_top_k,
                         float nms_threshold, float* d_out_scores, float* d_out_boxes, void* d_workspace,
                         size_t workspace_bytes, void* stream);

/* TP / FP matching of the per-class detections against ground truth (GPU half).
 * Replaces: eval_helper.bboxes_matching_batch utility/eval_helper.py:790-840 = bboxes_matching :700-788 (one
 *   tf.while_loop per class over the detections in score order) with bboxes_jaccard :671-699.
 *   d_det_boxes [N,num_classes-1,Kd,4] (xdet_det_postprocess's d_out_boxes), d_glabels [N,G] int32 (0 = padding),
 *   d_gbboxes [N,G,4], d_gdifficult [N,G] int32 -> d_tp, d_fp [N,num_classes-1,Kd] uint8, d_n_gbboxes
 *   [N,num_classes-1] int32 (ground-truth boxes of the class that are not 'difficult').
 * Zero-padded detections are matched like any other (the reference does the same and drops them later by score,
 * utility/metrics.py:170-176).  G == 0 (the reference would fail in tf.argmax): every detection is a false positive. */
int xdet_det_match(const float* d_det_boxes, const int* d_glabels, const float* d_gbboxes, const int* d_gdifficult,
                   int N, int num_classes, int Kd, int G, float matching_threshold, unsigned char* d_tp,
                   unsigned char* d_fp, int* d_n_gbboxes, void* stream);

/* ---------------------------------------------------------------------------------------
 * fp32-accurate PARITY MODE ("fp32x3", csrc/parity_ops.cu; not on the throughput path).
 * The reference computes every convolution / dense layer in fp32 (tf.layers.conv2d/dense, net/resnet_v2.py:89-100,
 * net/xception_body.py:224-233,381-400,450-475,540-558); north_star asks for box/score deltas within 1e-4 of it.
 * xdet_conv2d_bf16 reaches fp32-level error when both operands are split into three bf16 pieces and the six
 * significant cross products are summed in its fp32 accumulator: activations [mid|lo|hi|mid|hi|hi] (blocks of C
 * channels, written by xdet_split3_bf16) against weights [mid|hi|lo|hi|mid|hi] (packed by the host).
 * xdet_split3_bf16      d_src fp32, element (n,y,x,c) at n*sn + y*sy + x*sx + c*sc -> d_dst [N,H,W,out_cs] bf16,
 *                       out_cs >= 6*C, out_cs % 8 == 0, zero tail.
 * xdet_f32_post         v = x (+ residual) (ReLU if relu); out = v (NULL = not stored); out2 = v*scale2[c]+bias2[c]
 *                       (ReLU if relu2) (NULL = none): the fp32 form of the conv epilogue's residual / second-output
 *                       stages and of batch_norm_relu (net/resnet_v2.py:41-50).  [rows, C] fp32.
 * xdet_maxpool3x3s2_f32 tf.layers.max_pooling2d(3,2,'SAME') (+ residual, + second output) on NHWC fp32.
 * xdet_depthwise3x3_f32 depthwise half of tf.layers.separable_conv2d (net/xception_body.py:224-233) on NHWC fp32.
 */
int xdet_split3_bf16(const float* d_src, long long sn, long long sy, long long sx, long long sc, int N, int H, int W,
                     int C, void* d_dst, int out_cs, void* stream);
int xdet_f32_post(const float* d_x, const float* d_residual, int relu, float* d_out, const float* d_scale2,
                  const float* d_bias2, int relu2, float* d_out2, long long rows, int C, void* stream);
int xdet_maxpool3x3s2_f32(const float* d_src, float* d_dst, float* d_dst2, const float* d_scale2, const float* d_bias2,
                          const float* d_residual, int N, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left,
                          void* stream);
int xdet_depthwise3x3_f32(const float* d_src, const float* d_weights, float* d_dst, int N, int H, int W, int C,
                          int dilation, int relu_in, void* stream);

/* ---------------------------------------------------------------------------------------
 * fp32-ACCURATE convolution / GEMM on the tcgen05 tensor cores ("f16x2" precision, csrc/conv_gemm_f16x2.cu): the
 * precision the parity claim (box / score deltas within 1e-4 of the reference's fp32 graph) is benchmarked at.
 * Replaces the same TensorFlow layers as xdet_conv2d_bf16 (tf.layers.conv2d / dense: net/resnet_v2.py:89-100,
 * 320-325, net/xception_body.py:224-233,381-400,450-475,540-558), computed to fp32-level error.
 * Operands are "f16x2 planes": every fp32 value v is two fp16 values hi = fp16(v), lo = fp16((v - hi) * 2^11), stored
 * as two planes of the same layout (`*_plane` = elements between the planes).  |v| must stay below 65504 (fp16's
 * range; larger values saturate) -- weights are pre-scaled per output channel by a power of two by the host, the
 * inverse scale goes into `scale`.
 * Input  : planes [2][N,H,W,in_cs] fp16 (NHWC, in_cs % 8 == 0), written by xdet_split2_f16 or by the previous
 *          convolution's epilogue (out_pair / out2_pair).
 * Weights: planes [2][Cout][KH*KW][ceil(Cin/64)*64] fp16 (tap-major, as xdet_conv2d_bf16).
 * Output : v = acc*scale[c] + bias[c] (+ residual, fp32, laid out like `out`) (ReLU if relu), where acc is the fp32
 *          sum over the reduction, accumulated in chunks of <= chunk_kb*64 elements (the tensor core's fp32 accumulator
 *          truncates; chunk sums are added with round-to-nearest in registers);
 *          out (fp32, element strides out_s*, NULL = not stored), out_pair (f16x2 planes [2][N*Hout*Wout][pair_cs] of v);
 *          out2 = ReLU(v*scale2[c] + bias2[c]) (fp32, laid out like `out`) and out2_pair, all optional.
 * fold_w / in_wp as in xdet_conv_desc (the few-channel stem on a row-padded NHWC8 image). */
typedef struct {
  int N, H, W, Cin, in_cs;
  int Cout, KH, KW, dil_h, dil_w, pad_top, pad_left;
  int Hout, Wout;
  int stride_h, stride_w;
  int fold_w, in_wp;
  long long in_plane; /* elements between the hi and lo planes of the input */
  const void* weights;
  long long w_plane;  /* elements between the hi and lo planes of the weights */
  const float* scale;
  const float* bias;
  int relu;
  const float* residual;
  float* out;
  long long out_sn, out_sy, out_sx, out_sc;
  void* out_pair;
  long long pair_plane;
  int pair_cs;
  float* out2;
  const float* scale2;
  const float* bias2;
  void* out2_pair;
  int block_n;  /* 0 = auto; 32, 64 or 128 */
  int chunk_kb; /* 0 = 2: k-blocks (of 64 reduction elements) per accumulator flush */
  int max_ctas;
  int cluster;  /* 2 = clusters of two CTAs work on two M tiles of one N tile and multicast each other half of the B
                   tiles (a quarter less operand traffic L2 -> SM); 0 / 1 = independent CTAs */
} xdet_conv_f16x2_desc;
int xdet_conv2d_f16x2(const void* d_in_pair, const xdet_conv_f16x2_desc* desc, void* stream);
/* fp32 (element (n,y,x,c) at n*sn + y*sy + x*sx + c*sc) -> f16x2 planes [2][N,H,Wp,cs]: pixel (n,y,x) is written at
 * ((n*H + y)*Wp + x + x_off)*cs, channels C..cs-1 as zeros; only those pixels are written (a caller that pads rows,
 * Wp > W, zero-fills the buffer once).  relu != 0 applies ReLU first. */
int xdet_split2_f16(const float* d_src, long long sn, long long sy, long long sx, long long sc, int N, int H, int W,
                    int C, void* d_dst, int cs, int Wp, int x_off, long long plane, int relu, void* stream);

/* tf.layers.max_pooling2d(3, 2, 'SAME') on NHWC fp32 (C % 8 == 0) for the f16x2 precision: v = window max (+ residual);
 * any of dst (fp32), dst_pair (f16x2 planes of v), dst2 = ReLU(v*scale2 + bias2) (fp32) and dst2_pair may be NULL.
 * Replaces tf.layers.max_pooling2d at net/resnet_v2.py:326-328 and net/xception_body.py:272,302,322 together with the
 * batch_norm_relu / split pass that follows it. */
int xdet_maxpool3x3s2_f32x(const float* d_src, float* d_dst, void* d_dst_pair, float* d_dst2, void* d_dst2_pair,
                           const float* d_scale2, const float* d_bias2, const float* d_residual, long long pair_plane,
                           int N, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left, void* stream);

/* Depthwise 3x3 'SAME' stride-1 convolution on NHWC fp32 (C % 8 == 0, dilation 1 or 2, optional ReLU on load) for the
 * f16x2 precision: the depthwise half of tf.layers.separable_conv2d (net/xception_body.py:224-233).  d_dst (fp32)
 * and / or d_dst_pair (the f16x2 planes the pointwise convolution reads) may be NULL. */
int xdet_depthwise3x3_f32x(const float* d_src, const float* d_weights, float* d_dst, void* d_dst_pair, long long pair_plane,
                           int N, int H, int W, int C, int dilation, int relu_in, void* stream);
/* xdet_depthwise3x3_f32x has two implementations with identical results: TMA-staged tiles (default) and a register-window
 * kernel; 0 selects the second (tests compare the two, tools time them). */
void xdet_set_depthwise_f32_tma(int enabled);

/* Input pipeline of the eval / test scripts.
 * Replaces: light_head_preprocess_for_eval / _for_test preprocessing/common_preprocessing.py:383-458 with
 *   resize = WARP_RESIZE (the scripts' default): convert_image_dtype(uint8 -> float32) * 2, tf_image_whitened
 *   (:136-152) with the means / 127.5 of :36-38, tf.image.resize_images(BILINEAR, align_corners=False)
 *   (preprocessing/tf_image.py:307-319), transpose to NCHW.
 *   d_image [H,W,3] uint8 (RGB) -> d_out [3,Ho,Wo] fp32 (one image of the NCHW batch); h_means3 = host pointer to the
 *   three per-channel means already divided by 127.5.  fp32 arithmetic with one rounding per operation. */
int xdet_preprocess_eval_u8(const unsigned char* d_image, int H, int W, int Ho, int Wo, const float* h_means3,
                            float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * RPN proposals.
 * xdet_rpn_decode  Replaces: score/loc reshaping + softmax light_head_rfcn_eval.py:389-397 (train :295-305) and
 *   AnchorEncoder.decode_all_anchors preprocessing/anchor_manipulator.py:641-669 (prior_scaling = 1).
 *   d_rpn_out [N,fh,fw,ch_stride] fp32: class logits at channel cls_off + a*2 + {0,1}, box deltas at
 *   box_off + a*4 + {cy,cx,h,w}; anchors yref/xref [fh*fw], href/wref [A] (AnchorCreator, :698-743).
 *   -> d_scores [N, fh*fw*A], d_boxes [N, fh*fw*A, 4] (ymin,xmin,ymax,xmax); anchor order (y, x, a).
 * xdet_rpn_select  Replaces: get_proposals (inference branch) net/xception_body.py:402-444 =
 *   _bboxes_clip :173-194 -> _filter_and_sort_boxes :133-158 (tf.nn.top_k) -> _bboxes_nms :57-67
 *   (tf.image.non_max_suppression) -> _upsample_rois :196-213, plus _point2center :215-218.
 *   d_shuffle_keys [N, post_nms_top_n] fp32 or NULL: stands in for tf.random_shuffle (stable argsort of the
 *   first n keys; NULL = identity).  Outputs: d_rois [N,post,4] (ymin,xmin,ymax,xmax), d_rois_yxhw [N,post,4]
 *   (cy,cx,h,w; may be NULL), d_roi_scores [N,post] (may be NULL), d_nms_keep_idx [N,post] positions of the NMS
 *   survivors in the sorted top-k list, -1 padded (may be NULL; for tests).
 */
int xdet_rpn_decode(const float* d_rpn_out, int ch_stride, int cls_off, int box_off, const float* d_yref,
                    const float* d_xref, const float* d_href, const float* d_wref, int N, int fh, int fw, int A,
                    float* d_scores, float* d_boxes, void* stream);
size_t xdet_rpn_select_workspace_bytes(int N, int A_tot, int pre_nms_top_n);
int xdet_rpn_select(const float* d_scores, const float* d_boxes, int N, int A_tot, int pre_nms_top_n,
                    int post_nms_top_n, float nms_threshold, float min_size, const float* d_shuffle_keys,
                    float* d_rois, float* d_rois_yxhw, float* d_roi_scores, int* d_nms_keep_idx, void* d_workspace,
                    size_t workspace_bytes, void* stream);

/* Head post-processing.  Replaces: tf.nn.softmax(cls_score) and AnchorEncoder.ext_decode_rois
 * (light_head_rfcn_eval.py:406-410, preprocessing/anchor_manipulator.py:671-683; head_prior_scaling = 1).
 * d_rois [M,4] (ymin,xmin,ymax,xmax); d_head_out [M,ch_stride] fp32 with class scores at cls_off.. and the
 * 4 box deltas (cy,cx,h,w) at loc_off..  ->  d_probs [M,num_classes], d_boxes [M,4]. */
int xdet_head_decode(const float* d_rois, const float* d_head_out, int ch_stride, int cls_off, int num_classes,
                     int loc_off, long long M, float* d_probs, float* d_boxes, void* stream);
/* same, plus the 'classes' / 'probabilities' entries of the predictions dict (tf.argmax / tf.reduce_max over the class
 * axis, light_head_rfcn_eval.py:413-416): d_classes [M] int64 and d_best_prob [M] fp32, either may be NULL. */
int xdet_head_decode_ex(const float* d_rois, const float* d_head_out, int ch_stride, int cls_off, int num_classes,
                        int loc_off, long long M, float* d_probs, float* d_boxes, long long* d_classes,
                        float* d_best_prob, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training-step kernels (csrc/train_ops.cu).  Replace the TF ops/gradients of the training graph
 * (light_head_rfcn_train.py:277-451) that are not convolutions.
 *
 * Batch norm, training mode (tf.layers.batch_normalization(training=True, fused=True), net/resnet_v2.py:41-50):
 *   xdet_col_stats_bf16   d_sums[0..C) += column sums of x [rows, cs] bf16 (and d_sums[C..2C) += sums of squares);
 *                         also the bias gradient of a convolution (column sums of dy).
 *   xdet_bn_finalize      batch mean / biased variance -> scale = gamma/sqrt(var+eps), shift = beta - mean*scale
 *                         (what xdet_affine_relu_bf16 / the conv epilogue apply), mean, invstd (saved for backward);
 *                         moving_mean/var <- decay*moving + (1-decay)*batch (unbiased variance), NULL = no update.
 *   xdet_bn_relu_bwd_bf16 gradient of y = relu(x*scale+shift) w.r.t. x (two passes: reduce, apply), + d_add_in
 *                         (a gradient arriving over the identity shortcut); d_sums[0..C) = dbeta, [C..2C) = dgamma.
 * xdet_maxpool3x3s2_bwd_bf16  gradient of tf.layers.max_pooling2d(3,2,'SAME'): dy goes to the first maximum of each
 *                     window, whose position (kh*3+kw, one byte per output element) xdet_maxpool3x3s2_argmax_bf16 recorded.
 * xdet_nchw_f32_to_nhwc_bf16 / xdet_affine_relu_to_nchw_f32  repacks around the fp32 NCHW thin feature map; the
 *                     NHWC side has a channel pitch (490 channels live in rows of 496, zero tail).
 * xdet_softmax_ce     tf.nn.sparse_softmax_cross_entropy_with_logits: loss_row[r] and
 *                     dlogits[r,c] = w_all * row_w[r] * (softmax - onehot) (light_head_rfcn_train.py:361,385).
 * xdet_smooth_l1      modified_smooth_l1 (:257-275, sigma 1) summed over the 4 coordinates, times row_w; gradient.
 * xdet_sgd_momentum_* tf.train.MomentumOptimizer (:436-441) on fp32 masters in TF layout with the L2 term of :420
 *                     (grad + wd*w); the conv form reads dw in the packed layout of xdet_conv2d_wgrad_bf16 and
 *                     rewrites the bf16 forward / input-gradient packs (slices of fused packs via the offsets).
 * xdet_match_encode   iou_matrix + do_dual_max_match + target encoding (preprocessing/anchor_manipulator.py:40-94,
 *                     118-171 for anchors: box_img_stride 0 + d_ref_yxhw; :345-392 for RoIs): labels (class, 0 bg,
 *                     -1 ignore), targets [.,4], matched overlap scores.  gt labels <= 0 are padding.
 * xdet_sample_fg_bg   fg/bg sampling with up-sampling (:394-432; light_head_rfcn_train.py:321-358); the three
 *                     tf.random_shuffle calls are stable argsorts of the injected key arrays (keys >= 0).
 */
int xdet_col_stats_bf16(const void* d_x, long long rows, int C, int cs, int with_squares, float* d_sums, void* stream);
int xdet_bn_finalize(const float* d_sums, const float* d_gamma, const float* d_beta, long long rows, int C, float eps,
                     float decay, float* d_moving_mean, float* d_moving_var, float* d_scale, float* d_shift,
                     float* d_mean, float* d_invstd, void* stream);
/* xdet_col_stats_bf16 (with squares) + xdet_bn_finalize in ONE launch.  d_scratch: xdet_bn_train_scratch_bytes(C) bytes
 * that are ZERO when the call is issued; the kernel leaves them zero again, so one buffer (sized for the widest layer)
 * serves every batch-norm issued on the same stream. */
size_t xdet_bn_train_scratch_bytes(int C);
int xdet_bn_train_stats_bf16(const void* d_x, long long rows, int C, int cs, const float* d_gamma, const float* d_beta,
                             float eps, float decay, float* d_moving_mean, float* d_moving_var, float* d_scale,
                             float* d_shift, float* d_mean, float* d_invstd, void* d_scratch, void* stream);
/* bn_finalize + affine (+ReLU) in one launch, from statistics some producer already accumulated (xdet_conv_desc.stats):
 * every block derives scale/shift of its channels from d_sums ([0,C) sums, [C,2C) sums of squares over `rows` rows);
 * the first row of blocks also writes d_scale/d_shift/d_mean/d_invstd (what the backward reads) and the moving averages. */
int xdet_bn_train_apply_bf16(const void* d_x, void* d_y, long long rows, int C, const float* d_sums, const float* d_gamma,
                             const float* d_beta, float eps, float decay, float* d_moving_mean, float* d_moving_var,
                             float* d_scale, float* d_shift, float* d_mean, float* d_invstd, int relu, void* stream);
/* sums_zeroed != 0: d_sums (2*C floats: dbeta, dgamma) already holds zeros (e.g. a slice of a gradient buffer cleared at
 * the start of the step) -- the call skips its own memset. */
int xdet_bn_relu_bwd_bf16(const void* d_dy, const void* d_x, const float* d_scale, const float* d_shift,
                          const float* d_mean, const float* d_invstd, long long rows, int C, int relu,
                          const void* d_add_in, float* d_sums, void* d_dx, int sums_zeroed, void* stream);
/* dx = dy where y > 0 else 0 (gradient of a ReLU fused into a convolution epilogue); bf16, n elements (n % 8 == 0) */
int xdet_relu_bwd_bf16(const void* d_dy, const void* d_y, void* d_dx, long long n, void* stream);
/* fp32 forms of the kernels above for the fp32-ACCURATE training mode ("f16x2" precision; csrc/train_ops_f32.cu):
 * deterministic (one owner per channel, fp64 accumulation, no atomics).  xdet_col_stats_f32 overwrites d_sums unless
 * accumulate != 0; xdet_bn_relu_bwd_f32 overwrites d_sums ([0,C) = sum g, [C,2C) = sum g*xhat). */
int xdet_col_stats_f32(const float* d_x, long long rows, int C, int cs, int with_squares, int accumulate, float* d_sums,
                       void* stream);
int xdet_bn_relu_bwd_f32(const float* d_dy, const float* d_x, const float* d_scale, const float* d_shift,
                         const float* d_mean, const float* d_invstd, long long rows, int C, int relu,
                         const float* d_add_in, float* d_sums, float* d_dx, void* stream);
int xdet_relu_bwd_f32(const float* d_dy, const float* d_y, float* d_dx, long long n, void* stream);
/* dw [9, C] += depthwise 3x3 weight gradient on NHWC fp32 tensors (see xdet_depthwise3x3_wgrad_bf16) */
int xdet_depthwise3x3_wgrad_f32(const float* d_x, const float* d_dy, float* d_dw, int N, int H, int W, int C,
                                int dilation, int relu_in, void* stream);
int xdet_maxpool3x3s2_argmax_f32(const float* d_src, float* d_dst, unsigned char* d_argmax, int N, int H, int W, int C,
                                 int Ho, int Wo, int pad_top, int pad_left, void* stream);
int xdet_maxpool3x3s2_bwd_f32(const unsigned char* d_argmax, const float* d_dy, float* d_dx, int N, int H, int W, int C,
                              int Ho, int Wo, int pad_top, int pad_left, void* stream);
/* Weight gradient of the depthwise 3x3 'SAME' stride-1 convolution (depth multiplier 1, dilation 1 or 2) -- the
 * depthwise half of tf.layers.separable_conv2d in XceptionBody (net/xception_body.py:224-233), training mode:
 *   dw[kh*3+kw, c] += sum_{n,y,x} act(x[n, y+(kh-1)*dil, x+(kw-1)*dil, c]) * dy[n,y,x,c],  act = ReLU if relu_in.
 * x, dy NHWC bf16 (C % 8 == 0); dw [9, C] fp32 is ACCUMULATED into.  The input gradient needs no entry point:
 * it is xdet_depthwise3x3_bf16 on dy with the taps flipped. */
int xdet_depthwise3x3_wgrad_bf16(const void* d_x, const void* d_dy, float* d_dw, int N, int H, int W, int C,
                                 int dilation, int relu_in, void* stream);
/* fp32 [rows, cols] -> bf16 [rows, dst_pitch] is xdet_f32_to_bf16_rows above */
int xdet_maxpool3x3s2_bwd_bf16(const void* d_argmax, const void* d_dy, void* d_dx, int N, int H, int W, int C, int Ho,
                               int Wo, int pad_top, int pad_left, void* stream);
int xdet_nchw_f32_to_nhwc_bf16(const float* d_src, void* d_dst, int N, int C, int dst_cs, int HW, void* stream);
int xdet_affine_relu_to_nchw_f32(const void* d_src, const float* d_scale, const float* d_shift, float* d_dst, int N,
                                 int C, int src_cs, int HW, int relu, void* stream);
int xdet_softmax_ce(const float* d_logits, int ld, int C, const int* d_labels, const float* d_row_w, float w_all,
                    long long M, float* d_loss_row, float* d_dlogits, int dld, void* stream);
int xdet_smooth_l1(const float* d_pred, int ld, const float* d_target, const float* d_row_w, float w_all, long long M,
                   float* d_loss_row, float* d_dpred, int dld, void* stream);
int xdet_sgd_momentum_conv(const float* d_dw, float* d_w, float* d_mom, void* d_w_pack, void* d_w_dgrad_pack, int Cout,
                           int KH, int KW, int Cin, int pack_co_off, int pack_ci_off, int pack_cin_pad,
                           int pack_cout_pad, int fold, float lr, float momentum, float wd, float grad_scale,
                           void* stream);
int xdet_sgd_momentum_vec(const float* d_g, float* d_w, float* d_mom, long long n, float lr, float momentum, float wd,
                          float grad_scale, void* stream);
/* The whole apply_gradients (:436-441) in one launch: a device-resident table of items, one per variable.  A
 * convolution item is what xdet_sgd_momentum_conv takes (dw / w_pack / w_dgrad_pack already offset to the variable's slice
 * of a fused pack, regular layout only); a vector item (bias, beta, gamma) has taps == 0 and Cout = its length.  The
 * caller fills first_block with the running sum of the items' block counts -- tiles_ci * tiles_co * taps 32x32 tiles for
 * a convolution (tiles_* = ceil(C* / 32)), ceil(Cout / 256) for a vector -- and passes the total. */
typedef struct xdet_sgd_item {
  const float* dw;
  float* w;
  float* mom;
  void* w_pack;
  void* w_dgrad_pack; /* NULL: no input-gradient pack */
  int Cout, taps, Cin, cin_pad, cout_pad;
  int tiles_ci, tiles_co;
  int first_block;
  float wd;
  int reserved;
} xdet_sgd_item;
int xdet_sgd_momentum_multi(const xdet_sgd_item* d_items, int n_items, int total_blocks, float lr, float momentum,
                            float grad_scale, void* stream);
size_t xdet_match_workspace_bytes(int N, int G);
int xdet_match_encode(const float* d_boxes, long long box_img_stride, const float* d_ref_yxhw, const float* d_gt,
                      const int* d_gt_labels, int N, int A, int G, float allowed_border, float high_thres,
                      float low_thres, const float* prior_scaling4, int* d_labels, float* d_targets, float* d_scores,
                      void* d_workspace, void* stream);
int xdet_sample_fg_bg(const int* d_labels, const float* d_scores, float bg_low, int groups, int n, int exp_fg, int total,
                      const float* d_keys_fg, const float* d_keys_bg, const float* d_keys_up, int* d_workspace,
                      int* d_out, int* d_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XDET_B200_H_ */
