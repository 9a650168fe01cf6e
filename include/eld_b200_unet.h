/* eld_b200_unet.h - U-Net part of the C ABI (included by eld_b200.h).
 *
 * Replaces what PyTorch-eager dispatches to cuDNN/ATen for UNetSeeInDark (reference
 * models/arch/Unet.py:6-91) and ELDModel.optimize_parameters (models/ELD_model.py:469-475).
 * Activations are NHWC bf16 with an explicit channel pitch (so a concat buffer is just a wider
 * pitch: torch.cat at Unet.py:69,74,79,84 disappears); weights are bf16 K-major GEMM operands
 * packed from the fp32 PyTorch-layout master copy by eld_pack_weights.
 */
#ifndef ELD_B200_UNET_H
#define ELD_B200_UNET_H

/* kinds for eld_pack_weights: source layout is PyTorch's (Conv2d OIHW, ConvTranspose2d IOHW).  The packed
 * operand is OPAQUE: the logical K-major matrix listed below, stored as the shared-memory image the tiles
 * consume (blocks [n_tile][tap][64-channel chunk], 64B/128B swizzle applied) so that a block is one linear
 * bulk copy.  Same element count as the source. */
#define ELD_PACK_CONV_FPROP    0  /* [cout][ (kh*3+kw)*cin + ci ]            <- W[co][ci][kh][kw]      */
#define ELD_PACK_CONV_DGRAD    1  /* [cin ][ (kh*3+kw)*cout + co ]           <- W[co][ci][2-kh][2-kw]  */
#define ELD_PACK_DECONV_FPROP  2  /* [(kh*2+kw)*cout + co][ci]               <- Wt[ci][co][kh][kw]     */
#define ELD_PACK_DECONV_DGRAD  3  /* [ci][(kh*2+kw)*cout + co]               <- Wt[ci][co][kh][kw]     */
int eld_pack_weights(eld_ctx* ctx, const float* w, void* packed_bf16, int cout, int cin, int kind, void* stream);

#define ELD_ACT_NONE  0
#define ELD_ACT_LRELU 1  /* max(0.2x, x)  (Unet.py:102-104), fused into the producing tile           */
#define ELD_ACT_MASK  2  /* multiply by d lrelu/dx of `aux` (1 if aux>0 else 0.2): backward of lrelu */

/* y[n,h,w,y_c0:y_c0+cout] = act( conv3x3_pad1(x[n,h,w,x_c0:x_c0+cin]) + bias )   nn.Conv2d(k=3,p=1), Unet.py:11-44.
 * With ELD_PACK_CONV_DGRAD weights (cin/cout swapped) the same tile is the data gradient.
 * cin % 32 == 0, cout % 32 == 0; any h, w (partial tiles are masked).  bias may be NULL.  y (and aux): 32-byte aligned,
 * pitch and first channel multiples of 16 (the epilogue uses 256-bit accesses). */
int eld_conv3x3_bf16(eld_ctx* ctx, const void* x, int x_pitch, int x_c0, int cin, const void* w_packed,
                     const float* bias, void* y, int y_pitch, int y_c0, int cout, int n, int h, int w,
                     int act, const void* aux, int aux_pitch, int aux_c0, void* stream);

/* nn.ConvTranspose2d(cin, cout, 2, stride=2) (Unet.py:30,34,38,42) as GEMM + pixel-shuffle epilogue:
 * y[n, 2h+kh, 2w+kw, y_c0+co] = sum_ci x[n,h,w,ci] Wt[ci][co][kh][kw] + bias[co].  (h, w) = INPUT grid. */
int eld_deconv2x2_bf16(eld_ctx* ctx, const void* x, int x_pitch, int x_c0, int cin, const void* w_packed,
                       const float* bias, void* y, int y_pitch, int y_c0, int cout, int n, int h, int w,
                       void* stream);

/* data gradient of the above: dx[n,h,w,ci] = sum_{kh,kw,co} dy[n,2h+kh,2w+kw,co] Wt[ci][co][kh][kw],
 * optionally times the LeakyReLU derivative of aux (the deconv input activation). */
int eld_deconv2x2_dgrad_bf16(eld_ctx* ctx, const void* dy, int dy_pitch, int dy_c0, int cout,
                             const void* w_packed, void* dx, int dx_pitch, int dx_c0, int cin,
                             int n, int h, int w, int act, const void* aux, int aux_pitch, int aux_c0,
                             void* stream);

/* Weight gradients, accumulated (+=) in f32 into the PyTorch-layout gradient `dw` (zero it once per
 * step): conv dW[co][ci][kh][kw] += sum_pixels dz[.,co] * x[. + (kh-1,kw-1), ci]   (autograd of Unet.py:11-44);
 * deconv dWt[ci][co][kh][kw] += sum_pixels x[n,h,w,ci] * dy[n,2h+kh,2w+kw,co].  (h, w) = x's grid. */
int eld_conv3x3_wgrad_bf16(eld_ctx* ctx, const void* x, int x_pitch, int x_c0, int cin,
                           const void* dz, int dz_pitch, int dz_c0, int cout,
                           float* dw, int n, int h, int w, void* stream);
int eld_deconv2x2_wgrad_bf16(eld_ctx* ctx, const void* x, int x_pitch, int x_c0, int cin,
                             const void* dy, int dy_pitch, int dy_c0, int cout,
                             float* dw, int n, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-network step.  Parameters / gradients / Adam moments are FLAT fp32 device buffers in the
 * reference's state_dict order (conv1_1.weight, conv1_1.bias, ... conv10_1.bias; Unet.py:11-46), conv
 * weights OIHW, deconv weights IOHW - so released checkpoints copy in 1:1 and DDP all-reduces one buffer.
 */
typedef struct eld_unet eld_unet;
size_t eld_unet_param_count(void);                                   /* 7,760,484 for UNetSeeInDark(4,4) */
int    eld_unet_param_offset(const char* layer, int is_bias, size_t* offset, size_t* count);
size_t eld_unet_workspace_bytes(int n, int h, int w, int train);     /* activations (+gradients) + packed weights; train = 1 also
                                                                      * holds what the forward tiles leave for the backward: sign
                                                                      * words (1 bit per masked activation element) and pool codes
                                                                      * (1 byte per pooled element: argmax + signs) */
/* Inference (train = 0): h % 16 == 0 and w % 16 == 0, as for the reference network.  Training (train = 1):
 * h % 128 == 0, w % 256 == 0.  The caller owns `workspace` (device memory) for the lifetime of the object. */
int    eld_unet_create(eld_ctx* ctx, int n, int h, int w, int train, void* workspace, size_t bytes, eld_unet** out);
void   eld_unet_destroy(eld_unet* u);
/* The same network with 3-channel frames on either side (ELDModel.initialize, ELD_model.py:377-389: in_channels = 3 for
 * --stage_in srgb, out_channels = 3 for --stage_out srgb): conv1_1 becomes [32][cin][3][3], conv10_1 [cout][32][1][1];
 * everything else, and the meaning of every other entry point, is unchanged (x is [n][cin][h][w], out / target / dout
 * [n][cout][h][w]).  cin, cout in {3, 4}. */
size_t eld_unet_param_count_io(int cin, int cout);
int    eld_unet_param_offset_io(const char* layer, int is_bias, int cin, int cout, size_t* offset, size_t* count);
int    eld_unet_create_io(eld_ctx* ctx, int n, int h, int w, int train, void* workspace, size_t bytes,
                          int cin, int cout, eld_unet** out);
int    eld_unet_grad_buckets_io(int cin, int cout, size_t* offsets, int max_offsets);
/* ELDModel.forward (ELD_model.py:422-432): x f32 NCHW [n][4][h][w] -> out f32 NCHW [n][4][h][w] */
int    eld_unet_forward(eld_unet* u, const float* params, const float* x, float* out, void* stream);
/* forward + pixel loss (L1: mean |out-target|, losses.py:32; or MSE, see eld_unet_set_loss) + backward (ELD_model.py:411-420): grads is zeroed
 * and filled; *loss (device float) receives the mean absolute error.  No optimizer step, no host sync. */
int    eld_unet_train_step(eld_unet* u, const float* params, const float* x, const float* target,
                           float* out, float* grads, float* loss, void* stream);
/* The autograd seam (ELDModel.backward_G, ELD_model.py:411-420: `loss.backward()` through netG): after an
 * eld_unet_forward on a train = 1 object (activations stay in the workspace), back-propagate the caller's
 * dout = d(loss)/d(out) (f32 NCHW) into `grads` (zeroed and filled, like eld_unet_train_step).  Any loss the caller likes. */
int    eld_unet_backward(eld_unet* u, const float* params, const float* x, const float* dout, float* grads, void* stream);
/* Pixel loss of eld_unet_train_step (models/losses.py:29-36, --loss): 0 = nn.L1Loss (default), 1 = nn.MSELoss. */
#define ELD_LOSS_L1 0
#define ELD_LOSS_L2 1
int    eld_unet_set_loss(eld_unet* u, int kind);
/* Data-parallel overlap (the reference is single-GPU, ELD_model.py:187-190): the flat gradient is final in
 * eld_unet_grad_buckets() = 4 contiguous ranges in backward-completion order (decoder upv6..conv10_1, bottleneck
 * conv5_*, encoder conv2_1..conv4_2, first layer conv1_1..conv1_2); offsets[2k], offsets[2k+1] = first element, element
 * count of bucket k.
 * After eld_unet_bucket_events(u, 1) every eld_unet_train_step records an event on its stream when bucket k of `grads`
 * is final; eld_unet_wait_bucket makes `stream` (the caller's communication stream) wait for it - the caller then
 * all-reduces grads[offset, offset+count) there while the rest of backward runs, and runs Adam on a bucket once its
 * all-reduce has joined the compute stream. */
int    eld_unet_grad_buckets(size_t* offsets, int max_offsets);      /* returns the bucket count (4) */
int    eld_unet_bucket_events(eld_unet* u, int enable);
int    eld_unet_wait_bucket(eld_unet* u, int bucket, void* stream);
/* torch.optim.Adam step (ELD_model.py:400-401,475) on the flat buffers; grads are multiplied by
 * grad_scale first (1/world_size after a SUM all-reduce).  step counts from 1. */
/* Per-launch timing (CUDA events on the launch stream) of the steps issued after eld_unet_profile(u, 1);
 * eld_unet_profile_read synchronises, returns name[32] / ms / algorithmic FLOPs / algorithmic bytes per
 * launch and clears the log.  Used by bench.py for the live roofline numbers. */
int    eld_unet_profile(eld_unet* u, int enable);
/* Measurement aid (no reference counterpart): a one-thread kernel on `stream` that compares %clock64 with %globaltimer
 * for ~20 us and writes the SM clock in MHz that the preceding kernels were running at to *out_mhz_device. */
int    eld_clock_probe(eld_ctx* ctx, float* out_mhz_device, void* stream);
int    eld_unet_profile_read(eld_unet* u, int max, char* names32, float* ms, double* flops, double* bytes, int* count);
int    eld_adam_step(eld_ctx* ctx, float* params, const float* grads, float* m, float* v, size_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     float grad_scale, void* stream);

#endif
