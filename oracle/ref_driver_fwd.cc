// TEST INFRASTRUCTURE ONLY (oracle). Drives the reference's UNMODIFIED CPU op
// `PSROIAlignOp<CPUDevice,float>::Compute` (/root/reference/cpp/PSROIPooling/
// ps_roi_align_op.cc:205-250) through the TF stand-in in ref_shim/tf_shim.h.
// The reference source is #included where it lies; nothing is copied.
#include "ps_roi_align_op.cc"  // found via -I/root/reference/cpp/PSROIPooling

#include <string>

static std::string g_ref_fwd_err;

extern "C" const char* ref_psroi_last_error_fwd() { return g_ref_fwd_err.c_str(); }

// inputs[N,C,H,W], rois[N,R,4] -> pooled[N,R,G,C/G] f32, index[...] i32. Returns 0 on success.
extern "C" int ref_psroi_align_fwd(const float* inputs, const float* rois, float* pooled, int32_t* index,
                                   int N, int C, int H, int W, int R, int gw, int gh, int use_max,
                                   int threads) {
  tensorflow::OpKernelConstruction cons(gw, gh, use_max ? "max" : "mean");
  PSROIAlignOp<CPUDevice, float> op(&cons);
  if (!cons.status().ok()) { g_ref_fwd_err = cons.status().error_message(); return -1; }
  tensorflow::OpKernelContext ctx(threads);
  ctx.add_input(tensorflow::TensorShape({N, C, H, W}), inputs);
  ctx.add_input(tensorflow::TensorShape({N, R, 4}), rois);
  ctx.add_output_buffer(pooled);
  ctx.add_output_buffer(index);
  op.Compute(&ctx);
  if (!ctx.status().ok()) { g_ref_fwd_err = ctx.status().error_message(); return -2; }
  return 0;
}
