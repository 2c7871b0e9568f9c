// TEST INFRASTRUCTURE ONLY (oracle). Drives the reference's UNMODIFIED CPU op
// `PSROIAlignGradOp<CPUDevice,float>::Compute` (/root/reference/cpp/PSROIPooling/
// ps_roi_align_grad_op.cc:326-373; live functor :187-322) through ref_shim/tf_shim.h.
#include "ps_roi_align_grad_op.cc"  // found via -I/root/reference/cpp/PSROIPooling

#include <string>

static std::string g_ref_bwd_err;

extern "C" const char* ref_psroi_last_error_bwd() { return g_ref_bwd_err.c_str(); }

namespace tensorflow {
// work_sharder.h:47-48 declares it; TF's pool is replaced by std::thread.
void Shard(int max_parallelism, thread::ThreadPool*, int64 total, int64, std::function<void(int64, int64)> work) {
  ShardImpl(max_parallelism, total, work);
}
}  // namespace tensorflow

// inputs[N,C,H,W] (shape only), rois[N,R,4], pooled_grad[N,R,G,C/G], index -> grad[N,C,H,W].
extern "C" int ref_psroi_align_bwd(const float* inputs, const float* rois, const float* pooled_grad,
                                   const int32_t* index, float* grad, int N, int C, int H, int W, int R,
                                   int gw, int gh, int use_max, int threads) {
  tensorflow::OpKernelConstruction cons(gw, gh, use_max ? "max" : "mean");
  PSROIAlignGradOp<CPUDevice, float> op(&cons);
  if (!cons.status().ok()) { g_ref_bwd_err = cons.status().error_message(); return -1; }
  tensorflow::OpKernelContext ctx(threads);
  int G = gw * gh;
  ctx.add_input(tensorflow::TensorShape({N, C, H, W}), inputs);
  ctx.add_input(tensorflow::TensorShape({N, R, 4}), rois);
  ctx.add_input(tensorflow::TensorShape({N, R, G, G ? C / G : 0}), pooled_grad);
  ctx.add_input(tensorflow::TensorShape({N, R, G, G ? C / G : 0}), index);
  ctx.add_output_buffer(grad);
  op.Compute(&ctx);
  if (!ctx.status().ok()) { g_ref_bwd_err = ctx.status().error_message(); return -2; }
  return 0;
}
