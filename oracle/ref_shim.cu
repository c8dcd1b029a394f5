// ref_shim.cu -- extern "C" doorway onto the REFERENCE's own CUDA launchers, which are compiled
// unmodified from /root/reference by oracle/Makefile into oracle/_ref/libupsnet_ref.so.
// TEST INFRASTRUCTURE ONLY (second oracle / "kernel to beat"); never loaded by upsnet_b200/.
//
// Prototypes restated from:
//   upsnet/operators/src/roi_align_kernel.cu:351        roi_align_forward_gpu_kernel_launcher
//   upsnet/operators/src/deform_conv_kernel.cu:264      deformable_im2col_gpu_kernel_launcher
//   upsnet/operators/src/mod_deform_conv_kernel.cu:383  modulated_deformable_im2col_gpu_kernel_launcher
//   upsnet/nms/gpu_nms.hpp:14                           _nms
#include <cuda_runtime.h>

int roi_align_forward_gpu_kernel_launcher(cudaStream_t stream, const float* bottom_data,
                                          const float spatial_scale, const int num_rois,
                                          const int height, const int width, const int channels,
                                          const int pooled_height, const int pooled_width,
                                          const int sampling_ratio, const float* bottom_rois,
                                          float* top_data);
void deformable_im2col_gpu_kernel_launcher(cudaStream_t stream, const float* data_im,
                                           const float* data_offset, const int channels,
                                           const int height, const int width, const int ksize_h,
                                           const int ksize_w, const int pad_h, const int pad_w,
                                           const int stride_h, const int stride_w,
                                           const int dilation_h, const int dilation_w,
                                           const int parallel_imgs, const int deformable_group,
                                           float* data_col);
void modulated_deformable_im2col_gpu_kernel_launcher(
    cudaStream_t stream, const float* data_im, const float* data_offset, const float* data_mask,
    const int batch_size, const int channels, const int height_im, const int width_im,
    const int height_col, const int width_col, const int kernel_h, const int kenerl_w,
    const int pad_h, const int pad_w, const int stride_h, const int stride_w, const int dilation_h,
    const int dilation_w, const int deformable_group, float* data_col);
void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
          float nms_overlap_thresh, int device_id);

extern "C" {

int ref_roi_align_forward(const float* feat, float spatial_scale, int num_rois, int H, int W, int C,
                          int PH, int PW, int sampling_ratio, const float* rois, float* out,
                          void* stream) {
  roi_align_forward_gpu_kernel_launcher((cudaStream_t)stream, feat, spatial_scale, num_rois, H, W,
                                        C, PH, PW, sampling_ratio, rois, out);
  return (int)cudaGetLastError();
}

// one image: x [Cin,H,W], offset [dg*2*kh*kw,Ho,Wo] -> col [Cin*kh*kw, Ho*Wo]
int ref_deform_im2col(const float* x, const float* offset, int Cin, int H, int W, int kh, int kw,
                      int ph, int pw, int sh, int sw, int dh, int dw, int dg, float* col,
                      void* stream) {
  deformable_im2col_gpu_kernel_launcher((cudaStream_t)stream, x, offset, Cin, H, W, kh, kw, ph, pw,
                                        sh, sw, dh, dw, 1, dg, col);
  return (int)cudaGetLastError();
}

int ref_mod_deform_im2col(const float* x, const float* offset, const float* mask, int Cin, int H,
                          int W, int Ho, int Wo, int kh, int kw, int ph, int pw, int sh, int sw,
                          int dh, int dw, int dg, float* col, void* stream) {
  modulated_deformable_im2col_gpu_kernel_launcher((cudaStream_t)stream, x, offset, mask, 1, Cin, H,
                                                  W, Ho, Wo, kh, kw, ph, pw, sh, sw, dh, dw, dg,
                                                  col);
  return (int)cudaGetLastError();
}

// host boxes [N,5] ALREADY sorted by score desc (what gpu_nms.pyx:32-35 passes); keep_out = sorted positions
void ref_nms(int* keep_out, int* num_out, const float* boxes_host, int N, float thresh,
             int device_id) {
  _nms(keep_out, num_out, boxes_host, N, 5, thresh, device_id);
}

}  // extern "C"
