/*
 * t2v_b200 - C ABI of the B200-native text-to-video finetune hot path.
 *
 * The reference (ExponentialML/Text-To-Video-Finetuning) has no FFI of its own: its hot path is the chain
 *   train.py:339-347 (tensor_to_vae_latent) -> train.py:720-836 (finetune_unet) ->
 *   models/unet_3d_condition.py:325-500 (UNet3DConditionModel.forward) -> models/unet_3d_blocks.py ->
 *   diffusers leaf modules -> ATen -> cuDNN/cuBLAS/SDPA.
 * Each entry point below replaces one class of ATen/library call reached from those leaves; the reference call
 * site each one stands in for is cited on the declaration.  The Python host code in
 * text-to-video-finetuning_b200/ binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - All pointers are raw CUDA device pointers owned by the caller (PyTorch's allocator); nothing is allocated
 *     or freed inside, nothing synchronises the device.  `stream` is a cudaStream_t passed as void*.
 *   - Activations are bf16, channels-last: a (N, C, H, W) tensor is stored [N][H][W][C]; a clip (B, C, F, H, W)
 *     is stored [B][F][H][W][C] (== frames-major [B*F][H][W][C]).
 *   - Convolution weights are bf16 [Cout][KH][KW][Cin] (torch.channels_last of the diffusers (Cout,Cin,KH,KW)
 *     parameter); linear weights are [out][in].  Weight gradients are fp32, same layout, ACCUMULATED (+=).
 *   - Channel counts must be multiples of 8 (16-byte TMA rows); 3/4-channel tensors are padded to 8 by the host.
 *   - Return value: 0 on success, negative on error; t2v_last_error() returns a thread-local message.
 *     Functions are re-entrant and hold no thread-local state besides that message (autograd calls backward
 *     entry points from worker threads).
 */
#ifndef T2V_B200_H_
#define T2V_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int t2v_version(void);
const char* t2v_last_error(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
int64_t t2v_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Fused epilogue shared by the tensor-core entry points:  y = alpha * acc + bias[c] + rowbias[n][c] + residual
 */
typedef struct {
    const float* bias;      /* [Cout] fp32 or NULL                                                       */
    const float* rowbias;   /* [N][Cout] fp32 or NULL (ResnetBlock2D time_emb_proj broadcast over H, W)   */
    const void* residual;   /* bf16, same shape/layout as the output, or NULL                             */
    float alpha;
    int32_t out_fp32;       /* 0: bf16 output, 1: fp32 output                                             */
} T2VEpilogue;

/* Implicit-GEMM convolution forward on tcgen05 tensor cores (TMA-fed, zero padding by TMA OOB fill).
 * Replaces nn.Conv2d / nn.Conv3d(3,1,1) / nn.Linear forward as reached from
 *   ResnetBlock2D.conv1/conv2/conv_shortcut, Downsample2D.conv, Upsample2D.conv  (unet_3d_blocks.py:295-306,457-469,506-513,741-744)
 *   TemporalConvLayer.conv1..4  (unet_3d_blocks.py:308-314; tensor viewed as W=H*W, H=F, N=B, KH=3, KW=1)
 *   conv_in / conv_out          (unet_3d_condition.py:132,249)
 *   every nn.Linear             (viewed as W=rows, H=N=1, 1x1)
 * x [N][H][W][Cin], w [Cout][KH][KW][Cin], y [N][Ho][Wo][Cout];  Ho = (H + pad_h0 + pad_h1 - KH)/stride + 1.  */
int t2v_conv_fwd(const void* x, const void* w, void* y, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                 int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                 int32_t pad_w1, const T2VEpilogue* epi, void* stream);

/* Data gradient of the same convolution: dx [N][H][W][Cin] = conv_transpose(dy [N][Ho][Wo][Cout], w).
 * epi->residual (bf16, shape of dx) is added, which lets the caller sum gradient branches for free.
 * (autograd of the call sites above; train.py:861 accelerator.backward)                                */
int t2v_conv_dgrad(const void* dy, const void* w, void* dx, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                   int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                   int32_t pad_w1, const T2VEpilogue* epi, void* stream);

/* Weight gradient: dw [Cout][KH][KW][Cin] (fp32) += dy^T * shifted(x).  Split over pixels to fill the GPU;
 * partial products are reduced with red.global.add.f32 directly into the fp32 gradient buffer.              */
int t2v_conv_wgrad(const void* x, const void* dy, float* dw, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                   int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                   int32_t pad_w1, void* stream);

/* Strided-batched GEMM on the same kernel:  C[z1][z2] = alpha * opA(A[z1][z2]) * opB(B[z1][z2])^T  (+= if accumulate)
 *   a_kmajor=1: A is [M][K] rows (K contiguous);  a_kmajor=0: A is stored [K][M] (M contiguous)
 *   b_kmajor=1: B is [N][K] rows (K contiguous);  b_kmajor=0: B is stored [K][N] (N contiguous)
 * Used for the attention products QK^T, PV and their gradients (Attention in Transformer2DModel /
 * AutoencoderKL mid-block; diffusers AttnProcessor2_0 -> SDPA in the reference, train.py:138-152), and for
 * nn.Linear weight gradients.  out_mode: 0 bf16, 1 fp32, 2 fp32 accumulate (split-K allowed).              */
typedef struct {
    const void* ptr;
    int64_t ld;         /* elements between consecutive rows of the stored matrix */
    int64_t stride_z1;  /* elements */
    int64_t stride_z2;
    int32_t kmajor;
} T2VMat;
int t2v_bgemm(const T2VMat* A, const T2VMat* B, void* C, int64_t ldc, int64_t c_stride_z1, int64_t c_stride_z2,
              int32_t M, int32_t N, int32_t K, int32_t Z1, int32_t Z2, float alpha, int32_t out_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2V_B200_H_ */
