/* C ABI of the tensor-core convolution used by the update operator.
 * Replaces the cuDNN convolutions + torch.cat + elementwise gating kernels behind
 * UpdateModule.forward (networks/droid_net.py:118-150), ConvGRU.forward
 * (networks/modules/gru.py:19-32) and GraphAgg.forward (networks/droid_net.py:59-75).
 * See nerf_slam_b200/csrc/conv_igemm.cu for the epilogue modes. */
#ifndef NSLAM_NN_H_
#define NSLAM_NN_H_
#ifdef __cplusplus
extern "C" {
#endif

/* srcs: HOST array of n_src (<=4) DEVICE pointers to NHWC fp16 tensors [B,H,W,src_channels[i]]
 * (src_channels: HOST ints, each a multiple of 8); wpacked: DEVICE, packed by
 * nerf_slam_b200/conv.py::pack_weights; bias/gctx/gsum fp32 DEVICE; out0/out1 NHWC fp16 DEVICE
 * (out0 may be a channel slice: pass the pointer to its first channel and the FULL channel count
 * of the underlying tensor as out0_channels).  N in {16,64,128,256}. */
int nslam_conv_igemm(const void* const* srcs, const int* src_channels, int n_src, int B, int H, int W,
                     int KH, int KW, int pad, int N, const void* wpacked, const float* bias, int mode,
                     int act, const float* gctx, const void* net, const void* zbuf, float* gsum,
                     void* out0, int out0_channels, void* out1, int num_sms, void* stream);

#ifdef __cplusplus
}
#endif
#endif
