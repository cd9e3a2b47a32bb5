/* hawkeye_b200 — C ABI of the B200-native high-order-pooling hot path.
 *
 * One shared library (hawkeye_b200/libhawkeye_b200.so), plain C types, no torch types.
 * Conventions:
 *   - return 0 = success; <0 = argument/shape/alignment error, nothing was launched;
 *     >0 = cudaError_t from a launch.  hk_last_error() gives the text (thread-local).
 *   - the caller owns every device buffer including workspaces (query *_workspace_bytes);
 *     the library allocates nothing and never synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).  There is no library-owned device state: every entry point is re-entrant and
 *     may run concurrently with anything else on the GPU (no kernel waits on another thread-block cluster).
 *   - tensors are contiguous fp32; pointers 16-byte aligned; no CPU fallback: an unsupported
 *     shape is an error (-3), never a silent slow path.
 * Each entry point cites the reference interface it replaces (paths relative to the Hawkeye tree).
 */
#ifndef HAWKEYE_B200_H
#define HAWKEYE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* hk_version(void);
const char* hk_last_error(void);
long long hk_launch_count(void);      /* kernels launched by this library on the calling thread */
void hk_reset_launch_count(void);

/* ---- precision mode (process-wide; default 0, or $HK_PRECISE at first use) ---------------------------------------
 * 0: single-pass TF32 tensor-core products (tcgen05 kind::tf32 keeps 10 mantissa bits of each operand); every kernel
 *    that produces an operand of a later MMA rounds it to tf32 on store (round-to-nearest), so outputs of
 *    hk_conv3x3_*, hk_bn_*, hk_bilinear_pool_fwd, hk_cbp_fwd carry a 2^-11 relative quantisation.  Meets the 1e-3
 *    forward tolerance of the path; gradients below ReLU / max-pool kinks then differ from an fp32 run by branch
 *    flips (see tests/matched.py).
 * 1: 3xTF32 — every MMA operand is split into (hi, lo) tf32 halves and  A.B ~= Ah.Bh + Al.Bh + Ah.Bl  is accumulated
 *    by the same kernels in three passes; nothing is rounded on store.  fp32-class results (for parity runs against
 *    the fp32 reference) at more than 3x the cost; this mode allocates stream-ordered scratch (cudaMallocAsync). */
void hk_set_precise(int on);
int hk_get_precise(void);

/* ---- generic batched TF32 tensor-core GEMM (tcgen05 + TMA) -------------------------------------
 * C[b] = alpha*alpha_vec[b] * A[b].B[b] + diag*I + beta*beta_vec[b] * D[b]   (ReLU optional; C optionally transposed)
 * A logical [M,K]: a_mn_major=0 -> A[m*lda+k]; 1 -> A[k*lda+m].   B logical [K,N]: b_mn_major=0 -> B[n*ldb+k]; 1 -> B[k*ldb+n].
 * Replaces torch.bmm call sites model/methods/MPNCOV.py:117,132,154-160,178-194 and 1x1 convs resnet.py:34-37. */
int hk_gemm_tf32(const float* A, int a_mn_major, long long lda, long long strideA, const float* B, int b_mn_major,
                 long long ldb, long long strideB, float* C, long long ldc, long long strideC, int trans_c, int M,
                 int N, int K, int batch, float alpha, const float* alpha_vec, float diag, const float* D,
                 long long ldd, long long strideD, float beta, const float* beta_vec, int relu, void* stream);

/* same product, always as 3xTF32 whatever the precision mode (results that feed an exponential, e.g. CIN's softmax(-Gram)) */
int hk_gemm_3xtf32(const float* A, int a_mn_major, long long lda, long long strideA, const float* B, int b_mn_major,
                   long long ldb, long long strideB, float* C, long long ldc, long long strideC, int trans_c, int M,
                   int N, int K, int batch, float alpha, const float* alpha_vec, float diag, const float* D,
                   long long ldd, long long strideD, float beta, const float* beta_vec, int relu, void* stream);

/* ---- BCNN bilinear pooling: model/methods/BCNN.py:13-27 (BilinearPooling.forward) ----------------
 * x [B,C,HW] (NCHW feature map viewed as in BCNN.py:17) -> y [B,C*C] = normalize(sqrt(x x^T/HW + 1e-5)).
 * inv_norm_out (optional, [B]) receives 1/||z||.  Requires C%128==0.  H*W need not be a multiple of 4 (7x7 maps of 224x224
 * inputs): the workspace then also holds a zero-padded copy of x (TMA needs a 16-byte row pitch). */
size_t hk_bilinear_pool_fwd_workspace_bytes(int B, int C, int HW);
int hk_bilinear_pool_fwd(const float* x, float* y, float* inv_norm_out, int B, int C, int HW, void* workspace,
                         size_t workspace_bytes, void* stream);
/* backward of the same (what autograd derives for BCNN.py:13-27): dx [B,C,HW] from dy [B,C*C]; z is recomputed.
 * Default precision mode: y (forward) and dx (backward) are rounded to tf32 on store — they are operands of the next
 * MMA (classifier / last conv dgrad) — and the forward uses sqrt.approx (2^-22 rel.); hk_set_precise(1): fp32 as computed. */
size_t hk_bilinear_pool_bwd_workspace_bytes(int B, int C, int HW);
int hk_bilinear_pool_bwd(const float* x, const float* dy, float* dx, int B, int C, int HW, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- CBCNN compact bilinear pooling: model/methods/CBCNN.py:96-135 (CompactBilinearPooling.forward) ------------
 * h1,h2 int32 [C] and s1,s2 fp32 [C] are the count-sketch hash / sign vectors of CBCNN.py:76-91 (numpy seeds 1/3/5/7,
 * generated bit-exactly on the host).  y [B,d] = normalize(signed_sqrt(tensor-sketch)); pre [B,d] (pre-sqrt sketch)
 * is saved for the backward.  Requires C%128==0; for H*W % 4 != 0 a zero-padded copy of x is made in stream-ordered
 * scratch (cudaMallocAsync) — the one case outside the precise mode where the library allocates. */
int hk_cbp_fwd(const float* x, const int* h1, const int* h2, const float* s1, const float* s2, float* y, float* pre,
               int B, int C, int HW, int d, void* stream);
size_t hk_cbp_bwd_workspace_bytes(int B, int C, int d);
int hk_cbp_bwd(const float* x, const float* pre, const float* dy, const int* h1, const int* h2, const float* s1,
               const float* s2, float* dx, int B, int C, int HW, int d, void* workspace, size_t workspace_bytes,
               void* stream);

/* ---- Fast MPN-COV pooling head: model/methods/MPNCOV.py:105-230 -----------------------------------------------
 * Covpool (:105-134): x [B,C,M] -> cov [B,C,C] = X I_hat X^T; xc [B,C,ceil4(M)] receives the centred features at a 16-byte row
 *   pitch (saved for bwd; equal to [B,C,M] whenever M % 4 == 0).
 * Sqrtm (:137-202): coupled Newton-Schulz, iterN >= 2, forward and the reference's hand-derived backward formulae,
 *   all products as 3xTF32 tcgen05 GEMMs.  `saved` (hk_sqrtm_saved_floats floats) carries A, Y_i, Z_i, normA.
 * Triuvec (:205-230): row-major upper triangle [B,n,n] <-> [B,n(n+1)/2]. */
int hk_covpool_fwd(const float* x, float* cov, float* xc, int B, int C, int M, void* stream);
int hk_covpool_bwd(const float* xc, const float* g, float* dx, int B, int C, int M, void* stream);
size_t hk_sqrtm_saved_floats(int B, int n, int iterN);
size_t hk_sqrtm_fwd_workspace_bytes(int B, int n);
size_t hk_sqrtm_bwd_workspace_bytes(int B, int n);
int hk_sqrtm_fwd(const float* x, float* y, float* saved, int B, int n, int iterN, void* workspace,
                 size_t workspace_bytes, void* stream);
int hk_sqrtm_bwd(const float* x, const float* y, const float* g, float* saved, float* grad_x, int B, int n, int iterN,
                 void* workspace, size_t workspace_bytes, void* stream);
int hk_triuvec_fwd(const float* x, float* y, int B, int n, void* stream);
int hk_triuvec_bwd(const float* g, float* dx, int B, int n, void* stream);

/* ---- VGG-16 backbone: model/backbone/vgg.py:56-70 (Conv2d 3x3 s1 p1 + bias, ReLU, MaxPool2d(2,2)) ----------
 * Activations are NHWC fp32 inside the backbone.  Weights keep the reference layout [Cout,Cin,3,3] in the
 * state_dict and are re-packed per step: w_fwd [9][Cout][Cin], w_dgrad [9][Cin][Cout] (taps flipped). */
int hk_conv3x3_pack_weights(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin, void* stream);
/* y = relu?(conv3x3(x, w) + bias): implicit GEMM on tcgen05, TMA zero-fill = padding.  Cin%32==0, Cout%32==0. */
int hk_conv3x3_fwd(const float* x_nhwc, const float* w_fwd_packed, const float* bias, float* y_nhwc, int N, int H,
                   int W, int Cin, int Cout, int relu, void* stream);
/* relu(conv3x3(x, w) + bias) followed by MaxPool2d(2,2) (vgg.py:59-68: every pool of VGG-16 follows a conv + ReLU) in ONE
 * kernel: the epilogue reduces the 2x2 windows across lanes and the full-resolution map is never written.
 * pooled: [N,H/2,W/2,Cout] NHWC, or [N,Cout,H/2,W/2] when out_nchw (the last pool feeds the pooling heads in NCHW);
 * code (optional, training): one byte per pooled element for hk_maxpool2x2_bwd_idx (bits 0-1 first arg-max in scan order,
 * bit 2 = max > 0).  Bit-identical to hk_conv3x3_fwd + hk_maxpool2x2_fwd_idx.  H, W even; single-pass TF32 mode only
 * (HK_ERR_UNSUPPORTED under hk_set_precise(1): the 3xTF32 passes chain through the full-resolution map). */
int hk_conv3x3_fwd_pool(const float* x_nhwc, const float* w_fwd_packed, const float* bias, float* pooled,
                        unsigned char* code, int N, int H, int W, int Cin, int Cout, int out_nchw, void* stream);
/* same with stride 2 (ResNet v1.5 down-sampling 3x3, resnet.py:116): H, W are the input dims, output is H/2 x W/2 */
int hk_conv3x3_s2_fwd(const float* x_nhwc, const float* w_fwd_packed, const float* bias, float* y_nhwc, int N, int H,
                      int W, int Cin, int Cout, int relu, void* stream);
/* dx = conv3x3^T(dy, w) * (relu_mask_act > 0)  (mask optional: the ReLU output that produced x) */
int hk_conv3x3_dgrad(const float* dy_nhwc, const float* w_dgrad_packed, const float* relu_mask_act, float* dx_nhwc,
                     int N, int H, int W, int Cin, int Cout, void* stream);
/* dw [Cout,Cin,3,3] (reference layout), db [Cout] (optional) from x, dy (dy already ReLU-masked).  Cin%32==0, W%4==0. */
size_t hk_conv3x3_wgrad_workspace_bytes(int Cin, int Cout);
int hk_conv3x3_wgrad(const float* x_nhwc, const float* dy_nhwc, float* dw, float* db, int N, int H, int W, int Cin,
                     int Cout, void* workspace, size_t workspace_bytes, void* stream);
/* accumulate != 0: dw += ..., db += ... (gradient accumulation straight into a parameter's .grad buffer, no temporaries) */
int hk_conv3x3_wgrad_acc(const float* x_nhwc, const float* dy_nhwc, float* dw, float* db, int N, int H, int W, int Cin,
                         int Cout, void* workspace, size_t workspace_bytes, int accumulate, void* stream);
/* first layer (Cin=3, vgg.py:61): NCHW image in, NHWC out, bias+ReLU fused.  The 3x3x3 patches are materialised once
 * as X27 [N*H*W][32] (start of the fwd workspace) and reused by the weight/bias gradient. */
size_t hk_conv3x3_first_fwd_workspace_bytes(int N, int H, int W, int Cout);
int hk_conv3x3_first_fwd(const float* x_nchw, const float* w, const float* bias, float* y_nhwc, int N, int H, int W,
                         int Cout, void* workspace, size_t workspace_bytes, void* stream);
size_t hk_conv3x3_first_wgrad_workspace_bytes(int N, int H, int W, int Cout);
int hk_conv3x3_first_wgrad(const float* x27, const float* dy_nhwc, float* dw, float* db, int N, int H, int W,
                           int Cout, void* workspace, size_t workspace_bytes, void* stream);
int hk_conv3x3_first_wgrad_acc(const float* x27, const float* dy_nhwc, float* dw, float* db, int N, int H, int W,
                               int Cout, void* workspace, size_t workspace_bytes, int accumulate, void* stream);
/* MaxPool2d(2,2) on NHWC; out_nchw=1 writes the pooled map as NCHW (input of the pooling heads).
 * bwd routes dy to the first max (PyTorch semantics) and multiplies by (x>0), i.e. also applies the ReLU backward. */
int hk_maxpool2x2_fwd(const float* x_nhwc, float* y, int N, int H, int W, int C, int out_nchw, void* stream);
int hk_maxpool2x2_bwd(const float* x_nhwc, const float* dy, float* dx_nhwc, int N, int H, int W, int C, int dy_nchw,
                      void* stream);
/* training variants: fwd also records one byte per pooled element (bits 0-1 arg-max window position, bit 2 = max > 0);
 * bwd routes dy from that byte alone instead of re-reading the four pre-pool activations */
int hk_maxpool2x2_fwd_idx(const float* x_nhwc, float* y, unsigned char* code, int N, int H, int W, int C, int out_nchw,
                          void* stream);
int hk_maxpool2x2_bwd_idx(const unsigned char* code, const float* dy, float* dx_nhwc, int N, int H, int W, int C,
                          int dy_nchw, void* stream);
int hk_relu_mask_inplace(float* dy, const float* act, size_t n, void* stream);

/* ---- ResNet-50 v1.5 trunk support (model/backbone/resnet.py:89-252); activations NHWC [P = N*H*W, C] -----------------
 * stem 7x7/s2/p3 (resnet.py:176): patches X147 [P][160] (+ packed weights [64][160]) feed one tcgen05 GEMM. */
int hk_stem_im2col(const float* x_nchw, float* x147, int N, int H, int W, void* stream);
int hk_pack_stem_weights(const float* w, float* w147, int Cout, void* stream);
/* nn.BatchNorm2d in train mode (batch statistics, running stats updated with momentum, eps inside the sqrt):
 * y = [relu]((x-mean)*invstd*gamma + beta [+ residual]);  backward returns dx, dgamma, dbeta and (optionally) the
 * ReLU-masked dy for the residual branch (resnet.py:141-142 `out += identity; relu`). */
size_t hk_bn_workspace_bytes(long long P, int C);
int hk_bn_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y, float* save_mean,
              float* save_invstd, float* running_mean, float* running_var, float momentum, float eps, long long P, int C,
              int relu, void* workspace, size_t workspace_bytes, void* stream);
int hk_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                const float* residual, float* y, long long P, int C, int relu, void* stream);
int hk_bn_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
              const float* save_invstd, float* dx, float* dres, float* dgamma, float* dbeta, long long P, int C, int relu,
              void* workspace, size_t workspace_bytes, void* stream);
/* hk_bn_bwd with the ReLU mask recomputed instead of read: when the forward was relu((x-mean)*invstd*gamma + beta) WITHOUT a
 * residual, pass that beta as beta_for_mask and the backward re-evaluates the forward's own expression on x (same operation
 * order, so the same mask) — y is not read (may be null): a third less HBM traffic in both backward passes.
 * beta_for_mask == null: identical to hk_bn_bwd (mask = y > 0). */
int hk_bn_bwd_ex(const float* x, const float* y, const float* dy, const float* gamma, const float* beta_for_mask,
                 const float* save_mean, const float* save_invstd, float* dx, float* dres, float* dgamma, float* dbeta,
                 long long P, int C, int relu, void* workspace, size_t workspace_bytes, void* stream);
/* nn.MaxPool2d(3, 2, 1) (resnet.py:180) */
/* argmax (optional, [N,Ho,Wo,C] bytes): window position of the first maximum, consumed by the backward */
int hk_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* argmax, int N, int H, int W, int C, void* stream);
int hk_maxpool3x3s2_bwd(const unsigned char* argmax, const float* dy, float* dx, int N, int H, int W, int C,
                        void* stream);
/* stride-2 sampling of an NHWC map (1x1/s2 down-sample convs) and its adjoint (zero insertion); H, W = full-res dims */
int hk_subsample2(const float* x, float* y, int N, int H, int W, int C, void* stream);
int hk_upsample2_zero(const float* y, float* x, int N, int H, int W, int C, void* stream);
int hk_add_inplace(float* a, const float* b, size_t n, void* stream);
int hk_nhwc_to_nchw(const float* x, float* y, int N, int HW, int C, void* stream);
int hk_nchw_to_nhwc(const float* x, float* y, int N, int HW, int C, void* stream);
/* weight gradient of a matrix-form conv (1x1, or im2col'd stem): dw [Cout][K] = dY[P][Cout]^T . X[P][K] */
size_t hk_matconv_wgrad_workspace_bytes(long long P, int K, int Cout);
int hk_matconv_wgrad(const float* x, const float* dy, float* dw, long long P, int K, int Cout, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- channel interaction: model/methods/CIN.py:24-60, ChannelInteractionModule ---------------------
 * The Gram (:31), W.X (:34,:55), the 3x3 conv (:36,:57) and fc (:47-48) use hk_gemm_tf32 / hk_conv3x3_* / hk_linear_*; these are
 * the pieces in between: W_SCI = softmax(-G) row-wise (:32) and its backward; W_CCI = |W_SCI - weight_b * W_SCI[(b+B/2)%B]|
 * (:50-53; `per` = C*C elements per sample, B even) and its backward (d_sci, d_weight [B]); AdaptiveAvgPool1d(1) (:71) as a
 * row mean over the first `cols` of `ld` entries. */
int hk_softmax_neg_rows_fwd(const float* g, float* w, long long rows, int cols, void* stream);
int hk_softmax_neg_rows_bwd(const float* w, const float* dw, float* dg, long long rows, int cols, void* stream);
int hk_cci_weight_fwd(const float* w_sci, const float* weight, float* w_cci, int B, long long per, void* stream);
int hk_cci_weight_bwd(const float* w_sci, const float* weight, const float* d_cci, float* d_sci, float* d_weight, int B,
                      long long per, void* stream);
int hk_row_mean_fwd(const float* x, float* y, long long rows, int cols, int ld, void* stream);
int hk_row_mean_bwd(const float* dy, float* dx, long long rows, int cols, int ld, void* stream);

/* ---- OSME excitation (model/methods/OSME.py:8-24): s = sigmoid(m)[n,c] * x[n,c,:] and its backward; the
 * squeeze (AdaptiveAvgPool2d) is hk_row_mean_*, the two Linear layers are hk_linear_*, ReLU on the bottleneck hk_relu_* */
int hk_se_gate_fwd(const float* x, const float* m, float* s, long long rows, int hw, void* stream);
int hk_se_gate_bwd(const float* x, const float* m, const float* ds, float* dx, float* dm, long long rows, int hw,
                   void* stream);
/* ---- MAMC / N-pairs loss of OSMENet: model/loss/MAMC_loss.py:24-90 -------------------------------------------------------
 * hk_l2norm_rows_*: F.normalize(p=2, dim=1) of [rows, D] and its backward (inv_norm[r] = 1 / max(||x_r||, 1e-12)).
 * hk_npair_loss: prod [n,n] = F F^T of the n = batch x attention anchors (row-normalised features), cls[n] / part[n] the
 * label and attention index of each anchor.  Adds the N-pairs loss (sum of the three terms of eq. 11, divided by n) to the
 * pre-zeroed fp64 accumulator loss_acc[0] and writes d loss / d prod [n,n].  One launch, O(n) per anchor
 * (sum_k exp(neg_k - pos_j) = exp(-pos_j) * sum_k exp(neg_k)) instead of the reference's per-anchor Python loop. */
int hk_l2norm_rows_fwd(const float* x, float* y, float* inv_norm, int rows, int D, void* stream);
int hk_l2norm_rows_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int rows, int D, void* stream);
int hk_npair_loss(const float* prod, const int* cls, const int* part, double* loss_acc, float* dprod, int n, void* stream);
int hk_relu_fwd(const float* x, float* y, size_t n, void* stream);
int hk_relu_bwd(const float* y, const float* dy, float* dx, size_t n, void* stream);

/* ---- classifier nn.Linear (BCNN.py:42, CBCNN.py:26, MPNCOV.py:31) as skinny tcgen05 GEMMs ------------------- */
size_t hk_linear_fwd_workspace_bytes(int B, int F, int N);
int hk_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int F, int N, void* workspace,
                  size_t workspace_bytes, void* stream);
int hk_linear_dgrad(const float* dy, const float* w, float* dx, int B, int F, int N, void* stream);
int hk_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int B, int F, int N, void* stream);

/* ---- nn.CrossEntropyLoss(label_smoothing) fwd+bwd (train.py:211-212, :315-319); labels int64 ----------------
 * loss[0] = mean loss; dlogits (optional) = dloss/dlogits * grad_scale; correct (optional) = #argmax==label.
 * In the default precision mode dlogits is rounded to tf32 on store (it is the operand of the classifier's dgrad / wgrad
 * MMAs, which would otherwise truncate it); hk_set_precise(1) stores it unrounded. */
int hk_softmax_ce_ls(const float* logits, const long long* labels, float* loss, float* dlogits, int* correct, int B,
                     int K, float label_smoothing, float grad_scale, void* stream);

/* ---- input side: transforms.ToTensor + Normalize (dataset/transforms.py:14-19, test.py:80-85) fused on
 * the GPU: uint8 HWC batch [N,H,W,3] -> fp32 NCHW (x/255 - mean_c)/std_c; a quarter of the float pipeline's H2D bytes */
int hk_normalize_u8(const unsigned char* x_nhwc, float* y_nchw, int N, int H, int W, float mean0, float mean1, float mean2,
                    float std0, float std1, float std2, void* stream);

/* ---- optimizers over flat fp32 buffers: torch.optim.SGD (Examples/BCNN.py:40), Adam (Examples/MPN.py:14-18) */
int hk_sgd_momentum(float* p, const float* g, float* buf, size_t n, float lr, float momentum, float weight_decay,
                    float grad_scale, int first_step, void* stream);
int hk_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
            float weight_decay, float grad_scale, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HAWKEYE_B200_H */
