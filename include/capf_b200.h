/*
 * capf_b200.h -- C ABI of libcapf_b200.so: the B200 (sm_100a) implementation of the
 * ContextAware-PoseFormer single-frame lifting path  (CA_PF.forward, conpose.py:30-42).
 *
 * The reference exposes no FFI: all arithmetic is delegated to PyTorch (cuDNN / cuBLAS / ATen).
 * This header is the boundary a maintainer binds instead (ctypes stub: see INTEGRATION.md and
 * contextaware-poseformer_b200/lib.py).  Every entry point cites the reference call it replaces
 * as  file:line  relative to /root/reference/ContextPose/mvn/models/.
 *
 * Conventions
 *   - plain pointers + sizes only; every data pointer is a *device* pointer owned by the caller
 *     (the PyTorch caching allocator in practice).  The library allocates no device memory.
 *   - every call is asynchronous and ordered on the cudaStream_t passed as `void* stream`
 *     (pass torch.cuda.current_stream().cuda_stream); all calls are CUDA-graph capturable.
 *   - return value: 0 = ok, <0 = capf_status; capf_last_error() gives a thread-local message.
 *     Nothing throws or exits across the ABI.
 *   - activations are NHWC (channels innermost); token matrices are row-major [rows][cols].
 *   - there is no CPU path: with no usable device every compute call returns CAPF_ERR_CUDA.
 */
#ifndef CAPF_B200_H_
#define CAPF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPF_ABI_VERSION 13

typedef enum capf_status {
  CAPF_OK = 0,
  CAPF_ERR_ARG = -1,      /* bad shape / dtype / null pointer                       */
  CAPF_ERR_UNSUPPORTED = -2, /* valid request, no kernel variant for it              */
  CAPF_ERR_CUDA = -3      /* CUDA runtime / driver error (message has the detail)   */
} capf_status;

typedef enum capf_dtype { CAPF_F32 = 0, CAPF_F16 = 1, CAPF_BF16 = 2 } capf_dtype;

typedef enum capf_act { CAPF_ACT_NONE = 0, CAPF_ACT_RELU = 1, CAPF_ACT_GELU = 2 } capf_act;

/* Kernel family selector for CAPF_OP_CONV2D. */
typedef enum capf_impl {
  CAPF_IMPL_SIMT = 0,     /* fp32-accumulate CUDA-core implicit GEMM (any shape/dtype) */
  CAPF_IMPL_TCGEN05 = 1   /* TMA -> smem -> tcgen05.mma (TMEM accumulators), f16/bf16  */
} capf_impl;

typedef enum capf_op_kind {
  CAPF_OP_CONV2D = 1,
  CAPF_OP_FUSE_SUM = 2,
  CAPF_OP_MAXPOOL3X3S2 = 3,
  CAPF_OP_BILINEAR = 4,
  CAPF_OP_LAYERNORM = 5,
  CAPF_OP_ATTENTION = 6,
  CAPF_OP_REF_SAMPLE = 7,
  CAPF_OP_DEFORM_SAMPLE = 8,
  CAPF_OP_EMBED_COORD = 9,
  CAPF_OP_LEVELS_TO_JOINT = 10,
  CAPF_OP_CROP_NORMALIZE = 11,
  CAPF_OP_CAST = 12,
  CAPF_OP_PREPROCESS_U8 = 13,
  CAPF_OP_BASICBLOCK = 14,
  CAPF_OP_WARP_AFFINE_U8 = 15,
  CAPF_OP_POSE_ERRORS = 16,
  /* training step of the lifter (reference train.py:186-201, :337-345), fp32 */
  CAPF_OP_GEMM_F32 = 17,
  CAPF_OP_COLSUM = 18,
  CAPF_OP_LAYERNORM_BWD = 19,
  CAPF_OP_GELU = 20,
  CAPF_OP_GELU_BWD = 21,
  CAPF_OP_ATTENTION_BWD = 22,
  CAPF_OP_DEFORM_BWD = 23,
  CAPF_OP_ROWS_AXPY = 24,
  CAPF_OP_JOINT_TO_LEVELS = 25,
  CAPF_OP_ADAMW = 26,
  CAPF_OP_EXPAND_REDUCE = 27,
  CAPF_OP_MLP = 28
} capf_op_kind;

/*
 * One step of a forward "program".  The Python host (program.py) emits a flat array of these from the
 * reference's config keys (STAGE2..4, poseformer.*) and hands it to capf_plan_create().
 * Field meaning per kind (unused fields must be 0 / NULL):
 *
 * CAPF_OP_CONV2D  -- nn.Conv2d(bias=False)+BatchNorm2d(eval) [+residual] [+ReLU]  (pose_hrnet.py:79-136,
 *                    235-277, 382-408; networks/resnet.py:62-85; networks/refineNet.py:26-45) and
 *                    nn.Linear [+GELU] [+residual] (pose_dformer.py:25-31,49,56,132,214,221,240) as a
 *                    1x1 convolution over rows.
 *     i[0..10] = N,H,W,Cin,Cout,KH,KW,stride,pad,Ho,Wo   i[11]=act  i[12]=impl
 *     i[13] = tcgen05 kernel variant hint: 0 automatic, 1 per-tap TMA implicit GEMM, 2 shared-memory halo band
 *             (3x3 / stride 1 / pad 1 whose folded weights fit in shared memory); used by the A/B parity tests
 *     i[14] = with output segments (i[20] > 1): bit s set = segment s SKIPS the activation i[11] (NONE / RELU only); else ignored
 *     i[15] = per-tap kernel tile height hint: 0 automatic, 1 = 128 rows, 2 = 256 rows (two accumulators per B stage)
 *     i[16] = per-tap kernel column-tile width hint (0 automatic, else a multiple of 16 dividing Cout)
 *     i[17] = 2-CTA (cta_group::2) GEMM kernel for Linears over rows: 0 automatic (wide Linears with many rows),
 *             1 never, 2 always when the shape allows (K % 64 == 0, Cout % 16 == 0)
 *     i[18] = 1: split operands (precision "bf16x3"): x is [N,H,W,2*Cin] bf16, the hi | lo planes CAPF_OP_CAST (i[2] = Cin) writes,
 *             w is [Cout][KH*KW*(Wh | Wh)][KH*KW*Wl] bf16; the kernel accumulates hi*Wh + lo*Wh + hi*Wl (per-tap kernel only)
 *     i[19] = Cin2 > 0: a SECOND input x2 [N,H,W,Cin2] in in[5] (1x1 / stride 1 only): out = [x | x2] . w^T with
 *             w = [Cout][Cin + Cin2] -- a Bottleneck's conv3 and its downsample conv as one GEMM (pose_hrnet.py:116-136)
 *     i[20] = S in 2..4: OUTPUT SEGMENTS (tcgen05 per-tap kernel; no residual, no GELU, no split operands / second input).  Sibling
 *             convolutions that read the same x with the same geometry -- the fuse-layer convs of a HighResolutionModule that
 *             start from one branch (pose_hrnet.py:235-277) -- run as ONE GEMM over the Cout-concatenated weights; segment s
 *             (i[21], i[22], i[23] = channel widths of segments 0..2, multiples of 16; the last segment takes the rest of Cout)
 *             is written to its own dense tensor out[s] [N,Ho,Wo,width_s].  Bit-identical to the separate convolutions.
 *     in[0]=x  [N,H,W,Cin]        dtype_in
 *     in[1]=w  SIMT: [KH*KW*Cin][Cout] (tap-major rows, Cout contiguous); TCGEN05: [Cout][KH*KW*Cin];
 *              dtype_in, except x f32 -> w f32.  BatchNorm scale is pre-folded into w by the host.
 *     in[2]=bias[Cout] f32 (folded BN shift or Linear bias; may be NULL)
 *     in[3]=residual [N,Ho,Wo,Cout] dtype_out or NULL (may alias out[0])
 *     out[0]=y [N,Ho,Wo,Cout] dtype_out.    y = relu?( gelu?(conv+bias) + residual )
 *
 * CAPF_OP_EXPAND_REDUCE -- the tail of one Bottleneck and the head of the next (pose_hrnet.py:116-136 inside layer1 :421-427; the
 *                    same shapes in networks/resnet.py layer1) as one kernel over pixels (rows):
 *                        y = relu(t . W3^T + b3 + x)     conv3 (1x1, K1 -> N1) + bn3 + residual + ReLU of block i
 *                        u = relu(y . W1^T + b1)         conv1 (1x1, N1 -> N2) + bn1 + ReLU of block i + 1
 *                    y never makes the HBM round trip between the two GEMMs (it is still written once: it is the next residual).
 *                    Results are bit-identical to the two CAPF_OP_CONV2D ops (program.fuse_expand_reduce emits it for them).
 *     i[0]=rows (N*H*W)  i[1]=K1 (64)  i[2]=N1 (256)  i[3]=N2 (64)          16-bit dtypes, dtype_in == dtype_out
 *     in[0]=t [rows][K1]  in[1]=W3 [N1][K1]  in[2]=b3 [N1] f32 or NULL  in[3]=x [rows][N1]  in[4]=W1 [N2][N1]  in[5]=b1 [N2] f32 or NULL
 *     out[0]=y [rows][N1] (may alias in[3])   out[1]=u [rows][N2]
 *
 * CAPF_OP_MLP -- the Mlp of a 128-wide transformer block (pose_dformer.py:25-31; DeformableBlock :138-141, Block :78) as one kernel over
 *                    token rows:  X = X + gelu(t . W1^T + b1) . W2^T + b2   (fc1 128 -> 256, GELU erf form, hidden rounded to dtype_in and kept in
 *                    shared memory, fc2 256 -> 128, residual add in fp32).  Bit-identical to the two CAPF_OP_CONV2D ops (program.fuse_mlp emits it).
 *     i[0]=rows  i[1]=K1 (128)  i[2]=N1 (256)  i[3]=N2 (128)      dtype_in f16 | bf16, dtype_out f32
 *     in[0]=t [rows][K1] dtype_in (the LayerNorm output)  in[1]=W1 [N1][K1]  in[2]=b1 f32[N1] or NULL  in[3]=X [rows][N2] f32 (residual; may alias
 *     out[0])  in[4]=W2 [N2][N1]  in[5]=b2 f32[N2] or NULL        out[0]=X [rows][N2] f32
 *
 * CAPF_OP_FUSE_SUM -- HighResolutionModule fuse  y_i = ReLU(sum_j f_ij(x_j))  (pose_hrnet.py:294-301) with the
 *                    nn.Upsample(mode='nearest') of the j>i terms (:244) applied on the fly.
 *     i[0..3]=N,H,W,C  i[4]=n_terms(1..4)  i[5..8]=log2 upsample factor per term  i[9]=relu
 *     in[t]=term t, [N,H>>s,W>>s,C] dtype_in (summed in order t=0,1,..)   out[0]=[N,H,W,C] dtype_out
 *
 * CAPF_OP_MAXPOOL3X3S2 -- nn.MaxPool2d(3,2,1) (networks/resnet.py:105).  i[0..5]=N,H,W,C,Ho,Wo
 *
 * CAPF_OP_BILINEAR -- nn.Upsample(mode='bilinear', align_corners=True) (networks/globalNet.py:40,
 *                    networks/refineNet.py:61).   i[0..5]=N,H,W,C,Ho,Wo   in[0]=x  out[0]=y
 *
 * CAPF_OP_LAYERNORM -- nn.LayerNorm over the last dim (pose_dformer.py:65,72,120,138,206), optionally of
 *                    (x + x0) with x0 broadcast every `period` rows (the `x + x_0` of :120).
 *     i[0]=rows i[1]=D i[2]=period(0 = no x0)  i[3]=n_proj(0 = none)   f[0]=eps
 *     in[0]=x [rows][D] f32  in[1]=gamma f32  in[2]=beta f32  in[3]=x0 [period][D] f32 or NULL
 *     out[0]=y [rows][D] dtype_out
 *     n_proj > 0 (<= 8) fuses a narrow nn.Linear behind the norm -- the head LayerNorm(640)+Linear(640->3) of
 *     pose_dformer.py:205-208,240:  in[4]=W [n_proj][D] f32, in[5]=b [n_proj] f32 or NULL, out[0]=[rows][n_proj] f32
 *
 * CAPF_OP_ATTENTION -- Attention.forward softmax(q k^T * scale) v  (pose_dformer.py:47-55) for the two tiny
 *                    sequence shapes of the model: 5 levels of one joint (:231-234) or 17 joints (:235-238).
 *     i[0]=groups i[1]=seq(5|17) i[2]=heads i[3]=head_dim i[4]=token_stride_rows i[5]=group_stride_rows
 *     f[0]=scale   in[0]=qkv [rows][3*heads*head_dim] dtype_in, columns ordered (3,heads,head_dim) (:49)
 *     out[0]=[rows][heads*head_dim] dtype_out.  Row of token t of group g = g*group_stride + t*token_stride.
 *
 * CAPF_OP_REF_SAMPLE -- F.grid_sample(features, ref[B,17,1,2], bilinear, zeros, align_corners=True)
 *                    (pose_dformer.py:216-218) on up to 4 NHWC maps.
 *     i[0]=B i[1]=J i[2]=n_levels  i[3+3l..5+3l]=H_l,W_l,C_l
 *     in[0]=ref [B*J][2] f32 (x,y in [-1,1])  in[1+l]=map l [B,H_l,W_l,C_l] dtype_in
 *     out... see i/o below: out[0]=packed output base dtype_out; level l written at element offset i[15+l]
 *     as a dense [B*J][C_l] matrix.   out[1] (optional, int32 [n_levels][B*J][8]) = x_nw, y_nw, valid-mask
 *     (bit0 nw, bit1 ne, bit2 sw, bit3 se), 0, float bits of the sampled (x,y), 0, 0 -- the integer part of the
 *     gather and the exact position it was derived from, exposed for bit-exact tests.
 *
 * CAPF_OP_DEFORM_SAMPLE -- DeformableBlock sampling (pose_dformer.py:122-135): softmax over the 4 samples of
 *                    each head, tanh offsets + ref, F.grid_sample(..., padding_mode='border',
 *                    align_corners=True), and the sample-weighted sum.  The sum is taken *before*
 *                    embed_proj (legal: the weights sum to 1; changes rounding order only).
 *     i[0]=B i[1]=J i[2]=n_levels i[3+3l..]=H_l,W_l,C_l  i[15+l]=element offset of level l in out[0]
 *     in[0]=ref [B*J][2] f32   in[1+l]=map l dtype_in
 *     in[5]=ow [n_levels*B*J][48] f32: cols 0..15 attention_weights logits (head*4+sample),
 *           cols 16..47 sampling_offsets pre-tanh ((head*4+sample)*2+xy)
 *     out[0]: level l = dense [B*J*4][C_l] dtype_out (row = (b*J+j)*4+head)
 *     out[1] optional int32 [n_levels][B*J][16][8] corner record as above.
 *
 * CAPF_OP_EMBED_COORD -- coord_embed(kp2d) + Spatial_pos_embed[0], and Spatial_pos_embed[1+l] broadcast
 *                    into the level slabs ready for the feat_embed GEMMs to accumulate onto
 *                    (pose_dformer.py:214,223-225).   i[0]=B i[1]=J i[2]=D i[3]=n_slabs(=levels+1)
 *     in[0]=kp2d [B*J][2] f32  in[1]=W [D][2] f32  in[2]=b [D] f32  in[3]=pos [n_slabs][J][D] f32
 *     out[0]=X [n_slabs][B*J][D] f32 (level-major token stream)
 *
 * CAPF_OP_LEVELS_TO_JOINT -- rearrange '(b p) l c -> b p (l c)' (pose_dformer.py:235) from the level-major
 *                    stream.  i[0]=rows(B*J) i[1]=n_slabs i[2]=D   in[0]=[n_slabs][rows][D] f32  out[0]=[rows][n_slabs*D] f32
 *
 * CAPF_OP_CROP_NORMALIZE -- keypoints_2d_cpn_crop /= (96,128); -= 1  in place (conpose.py:34-35).
 *     i[0]=n_points   out[0]=crop [n][2] f32
 *
 * CAPF_OP_CAST -- dtype conversion of a dense array.  i[0],i[1]=element count (lo,hi 31-bit words) in[0] out[0]
 *     i[2] = 0, or C > 0 (f32 -> bf16 only): "split planes" -- every row of C elements becomes 2C bf16 values
 *            [hi (C) | lo (C)], hi = bf16(x), lo = bf16(x - hi): the A operand of a split-operand CONV2D (i[18] = 1)
 *     i[3] = 1 | 2 (with i[2] = C): the B operand instead -- rows [hi | hi | lo] (3C values) of a weight matrix [rows][C]
 *            (2: the source is stored transposed, [C][rows]); used for weights that change every step (training)
 *
 * CAPF_OP_BASICBLOCK -- one fused HRNet BasicBlock (pose_hrnet.py:66-95): y = relu(bn2(conv2(relu(bn1(conv1(x))))) + x), both
 *                    convolutions 3x3 / stride 1 / pad 1, C -> C channels (C = 32), 16-bit NHWC; the intermediate never leaves
 *                    the SM.  The host emits it for conv pairs that match (program.fuse_basic_blocks).
 *     i[0..3]=N,H,W,C   in[0]=x  in[1]=w1 [C][9C]  in[2]=b1 f32[C]  in[3]=w2 [C][9C]  in[4]=b2 f32[C]   out[0]=y
 *
 * CAPF_OP_PREPROCESS_U8 -- the image half of data_prefetcher.preload (mvn/datasets/utils.py:45-50, flip-test copy :67):
 *                    uint8 BGR HWC crops -> fp32 RGB NHWC, (x / 255 - mean[c]) / std[c] with IEEE divisions (bit-exact
 *                    with the torch expression), optionally mirrored along W (torch.flip(images, [2])).
 *     i[0..2]=B,H,W  i[3]=mirror(0|1)  i[4]=apply_std(1: HRNet, 0: CPN `x / 255 - mean`)
 *     in[0]=uint8 [B,H,W,3] (B,G,R)  in[1]=f32[6] device: mean R,G,B then std R,G,B   out[0]=f32 [B,H,W,3] (R,G,B)
 *
 * CAPF_OP_WARP_AFFINE_U8 -- crop_image (mvn/utils/img.py:51-69): cv2.warpAffine(frame, trans, (Wo, Ho), INTER_LINEAR),
 *                    constant 0 border, on a batch of uint8 HWC frames; the same bytes as OpenCV (1/32-pixel source
 *                    coordinates, 10-bit weights, round-half-up of the weighted sum).  The host passes the INVERSE
 *                    (crop -> frame) 2x3 matrices in fp64, computed from `trans` as OpenCV does (host mirror:
 *                    mvn/utils/img.py::invert_affine).  Frames share one padded [Hs,Ws] storage; in[2] optionally gives
 *                    each frame's live (h, w) -- Human3.6M cameras deliver 1000x1000 and 1002x1000.
 *     i[0]=B  i[1..2]=Hs,Ws  i[3..4]=Ho,Wo  i[5]=mode  (mode 1 only: i[6]=mirror along W, i[7]=apply_std)
 *     in[0]=uint8 [B,Hs,Ws,3]  in[1]=f64 [B,6] device  in[2]=int32 [B,2] (h,w) device or NULL  in[3]=f32[6] mean|std (mode 1)
 *     out[0]: mode 0 = uint8 [B,Ho,Wo,3] (channel order kept); mode 1 = f32 [B,Ho,Wo,3] RGB normalised like
 *             CAPF_OP_PREPROCESS_U8 applied to the mode-0 result (crop + data_prefetcher.preload in one pass)
 *
 * CAPF_OP_POSE_ERRORS -- the per-frame terms of evaluate_using_pred (mvn/datasets/human36m.py:358-422): MPJPE
 *                    (mvn/models/loss.py:16-22), P_MPJPE (:25-68, Procrustes-aligned) and MPJVE (:87-101) of frame n against
 *                    the frame prev[n] that precedes it inside its action.  fp64 arithmetic, one row per frame; the host
 *                    sums rows per action.
 *     i[0]=N frames  i[1]=J joints
 *     in[0]=pred f32 [N,J,3]  in[1]=gt f32 [N,J,3]  in[2]=int32 [N] prev (or NULL: no velocity term)   out[0]=f64 [N,3]
 *
 * ---- training step of `volume_net` (the reference trains it through autograd, train.py:186-201; the backbone is frozen,
 *      conpose.py:22-25).  All fp32, every reduction in a fixed order (bit-reproducible steps). ----
 *
 * CAPF_OP_GEMM_F32 -- C[M][N] = op(A) op(B) [+ bias[N]] [+ C]: forward nn.Linear (tb = 1), dgrad dy W (ta = tb = 0),
 *                    wgrad dy^T x (ta = 1, split over the rows).
 *     i[0..2]=M,N,K  i[3]=ta (A stored [K][M])  i[4]=tb (B stored [N][K])  i[5]=lda  i[6]=ldb  i[7]=ldc
 *     i[8]=K splits (<= 1: none)  i[9]=accumulate into C
 *     in[0]=A  in[1]=B  in[2]=bias or NULL   out[0]=C   out[1]=partials workspace [splits][M][N] (split-K only)
 * CAPF_OP_COLSUM -- out[n] = [out[n] +] sum_m x[m * ld + n] (bias gradients, Spatial_pos_embed gradient, partials).
 *     i[0]=M  i[1]=N  i[2]=ld  i[3]=accumulate   in[0]=x   out[0]=f32[N]   out[1]=workspace f32[256][N] or NULL
 * CAPF_OP_LAYERNORM_BWD -- nn.LayerNorm backward: dx (optionally accumulated) and per-block partial sums of dy*xhat / dy.
 *     i[0]=rows  i[1]=D (% 32, <= 640)  i[2]=period of the broadcast x0 rows (0: none)  i[3]=accumulate dx  i[4]=blocks
 *     f[0]=eps   in[0]=x  in[1]=gamma  in[2]=dy  in[3]=x0 or NULL   out[0]=dx [rows][D]  out[1]=partials [blocks][2][D]
 * CAPF_OP_GELU / CAPF_OP_GELU_BWD -- nn.GELU() (erf form) on a dense array / its backward.
 *     i[0],i[1]=element count (lo, hi 31-bit words; % 4)   in[0]=h  (BWD: in[1]=dy)   out[0]=y (BWD: dh)
 * CAPF_OP_ATTENTION_BWD -- backward of CAPF_OP_ATTENTION (same i[0..5], f[0]); probabilities are recomputed.
 *     in[0]=qkv f32 [rows][3*heads*hd]  in[1]=d out [rows][heads*hd]   out[0]=d qkv
 * CAPF_OP_DEFORM_BWD -- backward of CAPF_OP_DEFORM_SAMPLE (same i[0..18], in[0..5]) w.r.t. the attention logits and the
 *                    sampling offsets (grid_sample w.r.t. its grid, tanh, softmax); the maps are constants.
 *     out[0]=d ow f32 [levels*B*J][48]   out[1]=d sampled (INPUT, f32, the layout of DEFORM_SAMPLE's out[0])
 * CAPF_OP_ROWS_AXPY -- y[r][:] = [y[r][:] +] scale[(r % i[2]) / i[3]] * t[r][:]: DropPath residual adds and their backward.
 *     i[0]=rows  i[1]=D (% 4)  i[2]=mod  i[3]=div  i[4]=accumulate   in[0]=t  in[1]=scale or NULL (= 1)   out[0]=y
 * CAPF_OP_JOINT_TO_LEVELS -- inverse of CAPF_OP_LEVELS_TO_JOINT.  i[0]=R  i[1]=slabs  i[2]=D   in[0]=dY [R][slabs*D]  out[0]=dX
 * CAPF_OP_ADAMW -- torch.optim.AdamW update (decoupled weight decay) over one flat parameter buffer.
 *     i[0],i[1]=element count   in[0]=f32[7] device {lr, beta1, beta2, eps, weight_decay, 1-beta1^t, sqrt(1-beta2^t)}
 *     in[1]=grad  in[2]=exp_avg_sq (updated)   out[0]=param (updated)  out[1]=exp_avg (updated)
 */
typedef struct capf_op {
  int32_t kind;
  int32_t dtype_in;
  int32_t dtype_out;
  int32_t reserved;
  int32_t i[24];
  float f[4];
  const void* in[6];
  void* out[4];
} capf_op;

typedef struct capf_plan capf_plan;

/* Library / device ------------------------------------------------------------------------------------- */
int capf_abi_version(void);
const char* capf_last_error(void);
/* SM count, cc major/minor, global memory bytes of `device` (out[4] as int64). */
int capf_device_info(int device, int64_t* out4);

/* Programs ---------------------------------------------------------------------------------------------- */
/* Validates `ops`, selects kernel variants, encodes TMA descriptors for the pointers given.  Pointers are
 * baked in: the host keeps every referenced buffer alive and at the same address for the plan's lifetime. */
int capf_plan_create(const capf_op* ops, int n_ops, int device, capf_plan** out_plan);
/* Enqueue ops [first, first+count) on `stream`; count < 0 means "to the end". */
int capf_plan_run(const capf_plan* plan, int first, int count, void* stream);
int capf_plan_num_launches(const capf_plan* plan);   /* kernels one full run enqueues */
/* Name of the kernel (and, for the tcgen05 kernels, the tile shape) op `k` of the plan launches, written to `buf`
 * (NUL-terminated, at most `cap` bytes).  Used by bench.py / tools to attribute device time to kernels. */
int capf_plan_op_kernel(const capf_plan* plan, int k, char* buf, int cap);
int capf_plan_destroy(capf_plan* plan);
/* One-off execution of a single op (builds a throw-away plan): used by the per-operator parity tests. */
int capf_op_run(const capf_op* op, int device, void* stream);

/* Frame decode (mvn/datasets/human36m.py:565-567, cv2.imread(..., IMREAD_COLOR)) --------------------------
 * JPEG streams in host memory -> interleaved BGR uint8 frames in the padded device storage [n,Hs,Ws,3] that
 * CAPF_OP_WARP_AFFINE_U8 crops from, decoded by nvJPEG (loaded with dlopen on first use; without it these return
 * CAPF_ERR_UNSUPPORTED and capf_jpeg_available() 0).  sizes_hw (host, [n,2], may be NULL) receives each frame's (h, w).
 * Work is enqueued on `stream`; the decoder is bound to the device of its first use (one process per GPU).
 * Environment CAPF_JPEG_BACKEND (read at the first decode): unset / "default" = one nvjpegDecode per frame; "gpu_hybrid",
 * "hybrid", "hardware" = the whole batch through nvjpegDecodeBatched on that nvJPEG backend (CAPF_ERR_UNSUPPORTED if the
 * GPU / driver does not offer it; never a silent fallback). */
int capf_jpeg_available(void);
int capf_jpeg_info(const unsigned char* data, size_t length, int device, int* height, int* width);
int capf_jpeg_decode_batch(const unsigned char* const* data, const size_t* lengths, int n, unsigned char* frames, int Hs, int Ws,
                           int* sizes_hw, int device, void* stream);

/* Convenience wrappers over capf_op_run mirroring the reference call sites ------------------------------ */
int capf_crop_normalize(float* crop_xy, int n_points, void* stream);          /* conpose.py:34-35 */

#ifdef __cplusplus
}
#endif
#endif /* CAPF_B200_H_ */
