/*
 * movedepth_b200 -- C ABI of the B200-native MOVEDepth dense hot path.
 *
 * The reference (JeffWang987/MOVEDepth) has no FFI layer: its hot path is plain PyTorch
 * (SURVEY.md section 8b).  This header is the boundary a maintainer binds instead of the
 * ATen call sites cited on each entry point (file:line relative to the reference root).
 * INTEGRATION.md shows the ctypes stub.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device, fp32, 16-byte aligned,
 *     dense in the layout stated; the caller owns all buffers; nothing is allocated, freed
 *     or synchronised inside the library;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call
 *     only enqueues work on it, so calls are CUDA-graph capturable;
 *   - return 0 on success, <0 for an argument error detected before any launch,
 *     >0 = the cudaError_t of a failed launch.  mvd_last_error_string() returns a
 *     thread-local description of the last failure on the calling thread;
 *   - no C++ exception crosses the boundary; functions are re-entrant.
 */
#ifndef MOVEDEPTH_B200_H_
#define MOVEDEPTH_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MVD_ABI_VERSION 1

/* cost-volume output layouts (grouped volume, G correlation groups) */
#define MVD_LAYOUT_BGDHW 0 /* [B,G,D,h,w] : what reg3d's first Conv3d consumes (NCDHW)       */
#define MVD_LAYOUT_BDHWG 1 /* [B,D,h,w,G] : channels-last-3d view of the same logical tensor */

/* flags */
#define MVD_FLAG_NO_TMA 1   /* no TMA staging: taps are gathered from global memory (debug / parity isolation) */
#define MVD_FLAG_NO_TABLE 2 /* forward: no per-pixel tap-correlation table, recompute each bilinear cell directly */
#define MVD_FLAG_DBG_NO_STORE 4  /* forward, profiling only: compute everything but do not issue the TMA stores */
#define MVD_FLAG_PLAIN_STORE 16  /* forward: TMA-staged loads, but plain st.global instead of TMA stores */
#define MVD_FLAG_BWD_V2 32  /* backward: the round-1 kernel (8 short chunks per pixel, one hypothesis in flight per thread), for A/B runs */
/* backward, tuning: bits 8..11 of flags = number of hypothesis chunks per pixel (0 = default 3) */

int mvd_version(void);
const char* mvd_last_error_string(void);
/* number of SMs the persistent grids were sized for (0 before the first launch) */
int mvd_sm_count(void);

/* ---------------------------------------------------------------------------------------
 * K1  fused homography-warp cost volume, group correlation folded in.
 * Replaces: movedepth/layers.py:778-794 (generate_costvol: BackprojectDepth 581-586,
 * Project3D 601-621, F.grid_sample zeros/bilinear/align_corners=True at 791, product 792)
 * followed by the group mean movedepth/trainer.py:359 (reshape(B,D,C/G,G,h,w).mean(2)).
 *
 *   ref, src : [B,h,w,C] channels-last matching features (C == 32)
 *   prior    : [B,h,w]   mono depth prior, used with `ratio` when hyps == NULL
 *   ratio    : [B,D]     hypothesis d = prior * ratio[b,d]  (separable form of
 *                        schedule_depth_rangev2 / _zv2, layers.py:256-284 / 370-398)
 *   hyps     : [B,D,h,w] explicit hypotheses, or NULL
 *   K, invK, T : [B,4,4] row-major; K/invK at the volume's resolution, T = ref->src pose
 *   out      : grouped volume, G == 16 groups, group g = mean of channels {g, g+G},
 *              laid out per `out_layout`
 * ------------------------------------------------------------------------------------- */
int mvd_costvol_grouped_fwd(const float* ref, const float* src, const float* prior,
                            const float* ratio, const float* hyps, const float* K,
                            const float* invK, const float* T, float* out, int B, int C, int G,
                            int h, int w, int D, int out_layout, int flags, void* stream);

/* Backward of the above (autograd of grid_sample + product + group mean; the sampling grid
 * is built under no_grad in the reference, layers.py:784-790, so only ref/src get gradients).
 *   gout : same layout as `out`;  gref, gsrc : [B,h,w,C], OVERWRITTEN (zeroed inside). */
int mvd_costvol_grouped_bwd(const float* gout, const float* ref, const float* src,
                            const float* prior, const float* ratio, const float* hyps,
                            const float* K, const float* invK, const float* T, float* gref,
                            float* gsrc, int B, int C, int G, int h, int w, int D, int out_layout,
                            int flags, void* stream);

/* Reference-layout volume (public generate_costvol signature, layers.py:778-794):
 *   ref, src : [B,C,h,w] (NCHW, any C);  hyps : [B,D,h,w];  out : [B,D,C,h,w]. */
int mvd_costvol_full_fwd(const float* ref, const float* src, const float* hyps, const float* K,
                         const float* invK, const float* T, float* out, int B, int C, int h, int w,
                         int D, void* stream);
/*   gout : [B,D,C,h,w];  gref, gsrc : [B,C,h,w], OVERWRITTEN (zeroed inside). */
int mvd_costvol_full_bwd(const float* gout, const float* ref, const float* src, const float* hyps,
                         const float* K, const float* invK, const float* T, float* gref,
                         float* gsrc, int B, int C, int h, int w, int D, void* stream);

/* ---------------------------------------------------------------------------------------
 * K3  softmax over D + entropy + local-max depth regression, one pass over the logits.
 * Replaces: F.softmax(cost,1) movedepth/trainer.py:367,395; entropy layers.py:862-863;
 * localmax layers.py:796-812 (argmax, clamped +-radius window counted with duplicates,
 * soft index / (D-1), depth = 1/(inv_a + n*(inv_b-inv_a))).
 *   logits : [B,D,h,w];  inv_a, inv_b : [B,h,w] (1/hyps[:, -1], 1/hyps[:, 0])
 *   prob (nullable) : [B,D,h,w];  entropy : [B,h,w];  depth : [B,h,w];  amax : [B,h,w] int32
 * ------------------------------------------------------------------------------------- */
int mvd_regress_fwd(const float* logits, const float* inv_a, const float* inv_b, float* prob,
                    float* entropy, float* depth, int* amax, int B, int D, int hw, int radius,
                    void* stream);
/* Backward: glogits = d(entropy)/dlogits * g_entropy + d(depth)/dlogits * g_depth
 * (the argmax index is a constant, as in autograd).  Recomputes the softmax from logits. */
int mvd_regress_bwd(const float* logits, const float* inv_a, const float* inv_b, const int* amax,
                    const float* g_entropy, const float* g_depth, float* glogits, int B, int D,
                    int hw, int radius, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4  convex upsampling.  Replaces: convex_upsample movedepth/layers.py:200-214
 * (softmax over the 9 taps of mask.view(B,9,f,f,h,w), zero-padded 3x3 unfold of depth,
 * weighted sum, pixel shuffle).  f = 2**scale.
 *   depth : [B,h,w];  mask : [B,9*f*f,h,w];  out : [B,f*h,f*w]
 * ------------------------------------------------------------------------------------- */
int mvd_convex_up_fwd(const float* depth, const float* mask, float* out, int B, int h, int w, int f,
                      void* stream);
/*   gdepth : [B,h,w] OVERWRITTEN;  gmask : [B,9*f*f,h,w] OVERWRITTEN */
int mvd_convex_up_bwd(const float* depth, const float* mask, const float* gout, float* gdepth,
                      float* gmask, int B, int h, int w, int f, void* stream);

/* ---------------------------------------------------------------------------------------
 * K5+K6  photometric reprojection loss of ONE source frame.
 * Replaces: BackprojectDepth/Project3D + F.grid_sample(border) movedepth/trainer.py:501-507,
 * 519-529, 575-580; SSIM layers.py:646-677; compute_reprojection_loss trainer.py:535-550.
 *   depth : [B,H,W];  src, tgt : [B,3,H,W];  K, invK, T : [B,4,4]
 *   warped (nullable) : [B,3,H,W] the border-padded bilinear warp of src
 *   loss : [B,H,W] = ssim_w * mean_c SSIM(warped,tgt) + (1-ssim_w) * mean_c |tgt-warped|
 *   (ssim_w == 0 -> L1 only).  identity != 0: skip the warp, compare src with tgt directly.
 * ------------------------------------------------------------------------------------- */
int mvd_photometric_fwd(const float* depth, const float* src, const float* tgt, const float* K,
                        const float* invK, const float* T, float* warped, float* loss, int B, int H,
                        int W, float ssim_w, int identity, void* stream);
/* Backward: gloss [B,H,W] -> gdepth [B,H,W] (OVERWRITTEN) and, when gP != NULL, gP [B,3,4]
 * (OVERWRITTEN) = d loss / d (K @ T)[:3,:], from which d loss / d T = K[:3,:]^T @ gP.
 * Needs the forward's `warped`. */
int mvd_photometric_bwd(const float* gloss, const float* depth, const float* src, const float* tgt,
                        const float* warped, const float* K, const float* invK, const float* T,
                        float* gdepth, float* gP, int B, int H, int W, float ssim_w, void* stream);

/* ---------------------------------------------------------------------------------------
 * Loss assembly on the per-pixel reprojection losses of mvd_photometric_fwd.
 * Replaces: min over sources, identity auto-mask and masked mean -- movedepth/trainer.py:687-709
 * (mono), 621-662 (multi-frame: no identity term, mask = ones), 589-609 (fused).
 *   l0, l1 (nullable) : [n] losses of the source frames;  ident (nullable) : [n] identity loss,
 *   already the minimum over the source frames;  noise (nullable) : [n] N(0,1) tie-break noise
 *   reproj : [n] = min(l0, l1);  sel : [n] bytes (bit 0 = selected source, bit 1 = mask);
 *   sums : 2 doubles = [sum(reproj*mask), sum(mask)] (zeroed inside);
 *   loss : 1 float = sums[0] / (sums[1] + 1e-7)
 * Backward: g0, g1 (nullable) [n] OVERWRITTEN = gloss[0] / (sums[1] + 1e-7) where the pixel is
 * unmasked and the source was selected, else 0.
 * ------------------------------------------------------------------------------------- */
int mvd_reproj_select_fwd(const float* l0, const float* l1, const float* ident, const float* noise,
                          float* reproj, unsigned char* sel, double* sums, float* loss, long long n,
                          void* stream);
int mvd_reproj_select_bwd(const float* gloss, const double* sums, const unsigned char* sel, float* g0,
                          float* g1, long long n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Loss glue (csrc/lossglue.cu): the remaining per-scale tensor code of the loss section.
 *
 * mvd_disp_to_depth_{fwd,bwd}: F.interpolate(disp, [H,W], bilinear, align_corners=False) followed by
 * disp_to_depth (movedepth/trainer.py:512-515, layers.py:400-409):
 *   depth[b,y,x] = 1 / (inv_far + range * up(disp)[b,y,x]),  inv_far = 1/max_depth, range = 1/min_depth - 1/max_depth
 *   disp [B,hs,ws], depth/gdepth [B,H,W] (H, W integer multiples of hs, ws), gdisp [B,hs,ws] OVERWRITTEN.
 *
 * mvd_smooth_loss_{fwd,bwd}: get_smooth_loss (layers.py:630-643), with normalize != 0 preceded by the mean
 * normalisation disp / (disp.mean(2).mean(3) + 1e-7) of trainer.py:712-713 (mono) / 663-666 (mvs_smooth_loss):
 *   disp [B,h,w], img [B,3,h,w] (NCHW), loss: 1 float, work: mvd_smooth_loss_workspace_bytes(B) bytes written by fwd
 *   and read by bwd; dot: B doubles of scratch; gdisp [B,h,w] OVERWRITTEN = gloss[0] * d loss / d disp.
 *
 * mvd_masked_smooth_l1_{fwd,bwd}: the masked-augmentation consistency term (trainer.py:398-400):
 *   sel = bilinear(box mask [H,W] -> [h,w], align_corners=True) != 0, box = fh x fw zeros at box_xy = (x, y) (device int64[2],
 *   layers.py:52-69); loss = weight * mean over sel of smooth_l1(a - b); sel [B*h*w] bytes and sums (2 doubles) are
 *   written by fwd and read by bwd; ga, gb [B,h,w] OVERWRITTEN (gb may be NULL).
 * ------------------------------------------------------------------------------------- */
int mvd_disp_to_depth_fwd(const float* disp, float* depth, int B, int hs, int ws, int H, int W, float inv_far,
                          float range, void* stream);
int mvd_disp_to_depth_bwd(const float* gdepth, const float* depth, float* gdisp, int B, int hs, int ws, int H,
                          int W, float range, void* stream);
long long mvd_smooth_loss_workspace_bytes(int B);
int mvd_smooth_loss_fwd(const float* disp, const float* img, double* work, float* loss, int B, int h, int w,
                        int normalize, void* stream);
int mvd_smooth_loss_bwd(const float* gloss, const float* disp, const float* img, const double* work, double* dot,
                        float* gdisp, int B, int h, int w, int normalize, void* stream);
int mvd_masked_smooth_l1_fwd(const float* a, const float* b, const long long* box_xy, unsigned char* sel,
                             double* sums, float* loss, int B, int h, int w, int H, int W, int fh, int fw,
                             float weight, void* stream);
int mvd_masked_smooth_l1_bwd(const float* gloss, const float* a, const float* b, const unsigned char* sel,
                             const double* sums, float* ga, float* gb, long long n, float weight, void* stream);

/* ---------------------------------------------------------------------------------------
 * Skinny 2-D convolutions (csrc/conv2d_small.cu): FPN4's conv0 / conv1 stages (movedepth/networks/resnet_encoder.py:325-341,
 * Conv2d 453-475: 3->8, 8->8 at full resolution; 8->16 5x5 stride 2, 16->16 at half resolution), UncertNet's 8->8 layer
 * (depth_decoder.py:376-381) and the DepthDecoder's finest stage (depth_decoder.py:72-101: upconv(0,1) 16->16 and the
 * disparity head 16->1, which the caller zero-pads to 4 output channels).  Exact fp32 direct convolutions on the CUDA cores
 * (packed fp32x2 FMAs), no bias.  Replaces cuDNN fprop / dgrad / wgrad for these layers.
 *   x  : [B,H,W,cin] channels-last;  w : [cout][k][k][cin] (channels-last storage of the OIHW weight);
 *   pad: zero padding on every side: k/2 ("same"), or 0 for the stride-1 layers ("valid": the decoder's inputs arrive
 *        reflection-padded from mvd_decoder_prep);
 *   y, gy : [B,Ho,Wo,cout], Ho = (H + 2*pad - k)/stride + 1;  gx : [B,H,W,cin] OVERWRITTEN;  gw : like w, OVERWRITTEN
 *   supported (cin,cout,k,stride): (3,8,3,1) (8,8,3,1) (8,16,5,2) (16,16,3,1) (16,4,3,1); dgrad: all but (3,8,3,1) (the image).
 * ------------------------------------------------------------------------------------- */
int mvd_conv2d_small_supported(int cin, int cout, int k, int stride);
int mvd_conv2d_small_fwd(const float* x, const float* w, float* y, int B, int H, int W, int cin, int cout, int k,
                         int stride, int pad, void* stream);
int mvd_conv2d_small_dgrad(const float* gy, const float* w, float* gx, int B, int H, int W, int cin, int cout, int k,
                           int stride, int pad, void* stream);
long long mvd_conv2d_small_wgrad_workspace_bytes(int B, int H, int W, int cin, int cout, int k, int stride, int pad);
int mvd_conv2d_small_wgrad(const float* x, const float* gy, float* gw, void* workspace, long long workspace_bytes,
                           int B, int H, int W, int cin, int cout, int k, int stride, int pad, void* stream);

/* ---------------------------------------------------------------------------------------
 * DepthDecoder glue (csrc/decoder.cu): everything between two 3x3 convolutions of the U-Net decoder
 * (movedepth/networks/depth_decoder.py:72-101; ConvBlock / Conv3x3 / upsample of layers.py:521-553, 624-627) in one pass:
 *   xp = ReflectionPad2d(1)( cat( nearest_up(act(z + bias), up), skip ) ),   x3 = [hi | lo | hi] TF32 split of xp (optional)
 *   z    : [B,h,w,C1] channels-last raw output of the previous conv (bias NOT yet added), or a plain feature map
 *   bias : [C1] or NULL;  act: 0 = identity, 1 = ELU(alpha=1);  up: 1 or 2 (nearest)
 *   skip : [B,up*h,up*w,C2] channels-last or NULL (C2 == 0);  C1, C2 multiples of 4
 *   xp   : [B,up*h+2,up*w+2,C1+C2];  x3: [B,up*h+2,up*w+2,3*(C1+C2)] or NULL
 * bwd: gxp -> gz [B,h,w,C1], gskip [B,up*h,up*w,C2] (or NULL), gbias [C1] (or NULL), all OVERWRITTEN.
 * ------------------------------------------------------------------------------------- */
int mvd_decoder_prep_fwd(const float* z, const float* bias, const float* skip, float* xp, float* x3, int B, int h,
                         int w, int C1, int C2, int up, int act, void* stream);
int mvd_decoder_prep_bwd(const float* gxp, const float* z, const float* bias, float* gz, float* gskip, float* gbias,
                         int B, int h, int w, int C1, int C2, int up, int act, void* stream);

/* ---------------------------------------------------------------------------------------
 * ResNet stem max-pool: MaxPool2d(kernel 3, stride 2, padding 1) on channels-last activations (torchvision resnet,
 * used by ResnetEncoder.forward, movedepth/networks/resnet_encoder.py:113).  x [B,H,W,C] -> y [B,Ho,Wo,C],
 * Ho = (H-1)/2+1; idx [B,Ho,Wo,C] bytes = window position (ky*3+kx) of the first maximum; bwd: gx [B,H,W,C] OVERWRITTEN.
 * ------------------------------------------------------------------------------------- */
int mvd_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, void* stream);
int mvd_maxpool3x3s2_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, void* stream);

/* Pose-net outputs -> 4x4 camera transform: transformation_from_parameters = rot_from_axisangle (Rodrigues, angle + 1e-7) composed
 * with the translation matrix, optionally inverted (movedepth/layers.py:412-429, 464-518; called at trainer.py:462-466).
 *   axisangle, translation : [B,3];  M : [B,4,4] row-major, OVERWRITTEN;  invert : M = R^T * T(-t) instead of T(t) * R.
 *   bwd: g_axisangle, g_translation [B,3] OVERWRITTEN from gM [B,4,4] (forward-mode derivative of the same expression). */
int mvd_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert, void* stream);
int mvd_pose_matrix_bwd(const float* axisangle, const float* translation, const float* gM, float* g_axisangle,
                        float* g_translation, int B, int invert, void* stream);

/* ---------------------------------------------------------------------------------------
 * Device-side image pre-processing (csrc/datapipe.cu): what MonoDataset.preprocess does on the CPU through
 * Pillow / torchvision (movedepth/datasets/mono_dataset.py:104-126 pyramid of `transforms.Resize(..., ANTIALIAS)`,
 * 164/206 flip, 220-223 `transforms.ColorJitter`), on uint8 channels-last images, bit-exact.
 *   mvd_resample_u8   one pass of Pillow's ImagingResample: src [outer][n_in][inner] -> dst [outer][n_out][inner];
 *                     bounds [n_out][2] = (first tap, count), coeff [n_out][ksize] = 22-bit fixed-point Lanczos taps
 *                     (host-computed as Resample.c does); flip (nullable, one byte per image of `outer_per_image`
 *                     rows): taps are read through the mirrored index (horizontal pass of a flipped image)
 *   mvd_flip_copy_u8  copy [N,H,W,C] with the optional per-image mirrored x index
 *   mvd_u8_to_tensor  ToTensor: [N,H,W,3] uint8 -> [N,3,H,W] float, x / 255.0f
 *   mvd_jitter_blend_u8  in place, mode 0 brightness / 1 contrast / 2 saturation = Image.blend(degenerate, img, factor[n]);
 *                     sums: N uint64 of scratch (contrast); active (nullable): images with active[n] == 0 are skipped
 *   mvd_jitter_hue_u8 in place, torchvision adjust_hue: rgb2hsv, h += shift[n] (uint8 wrap), hsv2rgb (Convert.c)
 * ------------------------------------------------------------------------------------- */
int mvd_resample_u8(const unsigned char* src, unsigned char* dst, const int* bounds, const int* coeff, int ksize,
                    long long outer, int n_in, int n_out, int inner, const unsigned char* flip,
                    long long outer_per_image, void* stream);
int mvd_flip_copy_u8(const unsigned char* src, unsigned char* dst, int N, int H, int W, int C,
                     const unsigned char* flip, void* stream);
int mvd_u8_to_tensor(const unsigned char* src, float* dst, int N, int H, int W, void* stream);
int mvd_jitter_blend_u8(unsigned char* img, int N, int H, int W, int mode, const float* factor,
                        unsigned long long* sums, const unsigned char* active, void* stream);
int mvd_jitter_hue_u8(unsigned char* img, int N, int H, int W, const unsigned char* shift,
                      const unsigned char* active, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-tensor Adam on a flat fp32 arena (torch.optim.Adam semantics, no amsgrad, no
 * weight decay; replaces optimizer.step() movedepth/trainer.py:137-141, 272).
 *   step_size = lr / (1 - beta1^t);  bias2 = sqrt(1 - beta2^t)
 * ------------------------------------------------------------------------------------- */
int mvd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                  float beta1, float beta2, float eps, float step_size, float bias2,
                  float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Gradient gather into the flat arena (replaces autograd's per-parameter accumulate kernels, i.e. the reference's
 * optimizer.zero_grad() + AccumulateGrad of movedepth/trainer.py:270-271, with ONE launch per parameter group).
 *   table     : device int64 [nseg][14] = {src pointer (0 = no gradient: zeros are written), dst offset (elements),
 *               numel, linear (1: src is dense in the destination's order), dims[5] (destination physical order,
 *               outermost first, padded with 1), src strides[5] (elements, same order)}
 *   block_map : device int32 [nblocks][2] = {segment, chunk of mvd_gather_chunk() elements}
 * ------------------------------------------------------------------------------------- */
int mvd_gather_chunk(void);
int mvd_gather_segments(const long long* table, const int* block_map, int nblocks, float* dst, void* stream);

/* ---------------------------------------------------------------------------------------
 * 3xTF32 operand split for the tensor-core convolutions of movedepth_b200/precision.py:
 * x = hi + lo (hi = x rounded to TF32).  x: [rows, C] (channels-last), out: [rows, 3C] =
 * [hi, lo, hi] (pattern 0, activations) or [hi, hi, lo] (pattern 1, weights).  n = rows * C.
 * ------------------------------------------------------------------------------------- */
int mvd_split_tf32(const float* x, float* out, long long n, int C, int pattern, void* stream);

/* ---------------------------------------------------------------------------------------
 * reg3d output head: Conv3d(16 -> 1, 3x3x3, stride 1, zero padding 1, no bias) on the
 * channels-last full-resolution volume, exact fp32 (CUDA cores; an HBM-bound stencil).
 * Replaces: `self.prob = nn.Conv3d(base_channels, 1, 3, stride=1, padding=1, bias=False)`
 * movedepth/networks/resnet_encoder.py:254 and its call at 279 (cuDNN fprop/dgrad/wgrad).
 *   x  : [B,D,H,W,16] (channels-last-3d storage of the logical [B,16,D,H,W] tensor)
 *   w  : [16,3,3,3]   (= weight[0] of the module, contiguous)
 *   y, gy : [B,D,H,W];  gx : [B,D,H,W,16] OVERWRITTEN;  gw : [16,3,3,3] OVERWRITTEN
 *   wgrad needs a caller-owned workspace of mvd_conv3d_c16o1_wgrad_workspace_bytes(B,D,H,W).
 * The weights are staged in constant memory by fwd/dgrad (stream-ordered copy): concurrent
 * calls with DIFFERENT weights on different streams are not supported.
 * ------------------------------------------------------------------------------------- */
int mvd_conv3d_c16o1_fwd(const float* x, const float* w, float* y, int B, int D, int H, int W,
                         void* stream);
int mvd_conv3d_c16o1_dgrad(const float* gy, const float* w, float* gx, int B, int D, int H, int W,
                           void* stream);
long long mvd_conv3d_c16o1_wgrad_workspace_bytes(int B, int D, int H, int W);
int mvd_conv3d_c16o1_wgrad(const float* gy, const float* x, float* gw, void* workspace,
                           long long workspace_bytes, int B, int D, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------
 * reg3d first layer: Conv3d(16 -> 16, 3x3x3, stride 1, zero padding 1, no bias) on the
 * channels-last full-resolution volume as a tensor-core implicit GEMM (TMA-staged slices).
 * Replaces: ConvBnReLU3D.conv of `conv0`, movedepth/networks/resnet_encoder.py:178, 231, 258
 * (cuDNN fprop) and its data gradient (cuDNN dgrad).
 *   in, out : [B,D,H,W,16];  w : [16,16,3,3,3] (co, ci, kd, kh, kw), contiguous
 *   mode 0  : out = conv(in, w)            (forward)
 *   mode 1  : out = d loss / d input given in = d loss / d output   (data gradient)
 *   passes 3: 3xTF32 operand split (near-fp32, the forward policy); passes 1: single-pass TF32
 * ------------------------------------------------------------------------------------- */
int mvd_conv3d_c16c16(const float* in, const float* w, float* out, int B, int D, int H, int W,
                      int mode, int passes, void* stream);
/* The same contract on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, every tap's
 * A operand = the TMA-staged slice at a shifted start address.  flags: 0; profiling only (results become meaningless):
 * 2 = skip the MMAs, 4 = skip the global stores, 8 = skip the hi/lo operand split. */
int mvd_conv3d_c16c16_tc(const float* in, const float* w, float* out, double* bn_sums, int B, int D, int H, int W,
                         int mode, int passes, int flags, void* stream);
/* bn_sums (nullable, 32 doubles, ZEROED BY THE CALLER): the epilogue adds the per-channel sums [sum y (16), sum y^2 (16)] of
 * the output it writes -- the BatchNorm statistics of ConvBnReLU3D (resnet_encoder.py:175-182) without another pass over
 * the 283 MB volume; feed them to mvd_bn_finalize. */
/* Weight gradient on tcgen05 (single-pass TF32; MN-major operands, the four M-groups of A are the x slice shifted by
 * kw positions); same workspace / reduction scheme as mvd_conv3d_c16c16_wgrad. */
long long mvd_conv3d_c16c16_wgrad_tc_workspace_bytes(int B, int D, int H, int W);
int mvd_conv3d_c16c16_wgrad_tc(const float* gy, const float* x, float* gw, void* workspace,
                               long long workspace_bytes, int B, int D, int H, int W, void* stream);
/* Weight gradient of the same layer, exact fp32 (packed FFMA2 on the CUDA cores):
 *   gy, x : [B,D,H,W,16];  gw : [16,16,3,3,3] OVERWRITTEN;  workspace of
 *   mvd_conv3d_c16c16_wgrad_workspace_bytes(B,D,H,W) bytes (per-tile partials, reduced deterministically). */
long long mvd_conv3d_c16c16_wgrad_workspace_bytes(int B, int D, int H, int W);
int mvd_conv3d_c16c16_wgrad(const float* gy, const float* x, float* gw, void* workspace,
                            long long workspace_bytes, int B, int D, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training-mode BatchNorm on a channels-last activation viewed as a row-major [M, C] matrix
 * (M = N*D*H*W, C a power of two in [4,1024]), fused with ReLU and the residual add.
 * Replaces: nn.BatchNorm3d + F.relu in ConvBnReLU3D (movedepth/networks/resnet_encoder.py:175-182),
 * nn.BatchNorm2d + F.relu in Conv2d (453-475), bn/relu/add of the torchvision ResNet blocks
 * (74-121) and, under data-parallel training, nn.SyncBatchNorm (movedepth/trainer.py:69-129):
 * the caller all-reduces `sums` / `sums2` (2C doubles) between the two kernels of a pass.
 *   mvd_bn_stats      sums = [sum x (C), sum x^2 (C), arrival counter (1)] = 2C+1 doubles  (zeroed inside)
 *                     peers != NULL (data-parallel training): the last block of the reduction all-reduces the 2C sums
 *                     over NVLink peer memory in place (peers / rank / world / nmax as for mvd_peer_allreduce_f64), so
 *                     no separate exchange launch sits between the reduction and its consumer
 *   mvd_bn_finalize   stats = [mean, invstd, scale = w*invstd, shift = b - mean*scale] (4C floats),
 *                     running_mean/var updated in place (unbiased variance), count = rows over all ranks;
 *                     num_batches_tracked (device int64, nullable) is incremented by one
 *   mvd_bn_apply      y = relu?(x*scale + shift (+ residual));  relu: bit 0 = ReLU, bit 1 = the residual is added AFTER
 *                     the ReLU (U-Net skip, resnet_encoder.py:272-276) instead of before it (ResNet block)
 *   mvd_bn_bwd_reduce sums2 = [sum g (C), sum g*xhat (C), counter] (2C+1 doubles; all-reduced in place when peers != NULL,
 *                     this rank's own sums are then copied to local_sums2 (2C doubles) first),
 *                     g = gy * (y > 0) when relu; y may be NULL when the forward had
 *                     no residual: the mask is then recomputed from x*scale + shift (bit-identical), saving one read
 *   mvd_bn_bwd_apply  gx = w*invstd*(g - sum g/count - xhat * sum g*xhat/count); gres = g (nullable);
 *                     gw = sum g*xhat, gb = sum g (nullable; pass NULL when sums2 was all-reduced)
 * ------------------------------------------------------------------------------------- */
int mvd_bn_stats(const float* x, long long M, int C, double* sums, const unsigned long long* peers, int rank,
                 int world, int nmax, void* stream);
int mvd_bn_finalize(const double* sums, double count, const float* weight, const float* bias,
                    float* running_mean, float* running_var, float momentum, float eps, float* stats,
                    int C, long long* num_batches_tracked, void* stream);
int mvd_bn_apply(const float* x, const float* residual, const float* stats, float* y, long long M,
                 int C, int relu, void* stream);
int mvd_bn_bwd_reduce(const float* gy, const float* x, const float* y, const float* stats, double* sums2,
                      double* local_sums2, long long M, int C, int relu, const unsigned long long* peers, int rank,
                      int world, int nmax, void* stream);
int mvd_bn_bwd_apply(const float* gy, const float* x, const float* y, const float* stats,
                     const float* weight, const double* sums2, double count, float* gx, float* gres,
                     float* gw, float* gb, long long M, int C, int relu, void* stream);

/* One-kernel forms for activations that fit the L2 (most layers of the step): phase 1 reduces, a grid barrier (the last block
 * to arrive performs the SyncBatchNorm exchange when peers != NULL), phase 2 applies with x re-read from the L2.  Same
 * arithmetic and outputs as stats + finalize + apply / bwd_reduce + bwd_apply above.  `workspace` = mvd_bn_workspace_doubles()
 * doubles, zeroed ONCE by the caller; the kernels hand it back zeroed, so consecutive calls on one stream share it (calls
 * that can run concurrently need their own).  The grid never exceeds 2/3 of the SM count at two resident blocks per SM. */
int mvd_bn_workspace_doubles(void);
int mvd_bn_fwd_fused(const float* x, const float* residual, const float* weight, const float* bias, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float momentum, float eps, double count,
                     float* stats, float* y, long long M, int C, int relu, double* workspace,
                     const unsigned long long* peers, int rank, int world, int nmax, void* stream);
int mvd_bn_bwd_fused(const float* gy, const float* x, const float* y, const float* stats, const float* weight, double count,
                     float* gx, float* gres, float* gw, float* gb, double* local_sums2, long long M, int C, int relu,
                     double* workspace, const unsigned long long* peers, int rank, int world, int nmax, void* stream);

/* ---------------------------------------------------------------------------------------
 * SyncBatchNorm statistics exchange over NVLink peer memory (replaces the all_gather / all_reduce
 * of nn.SyncBatchNorm, movedepth/trainer.py:69-129, for the 2C-double vectors of mvd_bn_stats /
 * mvd_bn_bwd_reduce).  `peers` is a DEVICE array of `world` addresses: the same symmetric
 * buffer (mvd_peer_allreduce_buffer_bytes(world, nmax) bytes, zero-initialised once) as mapped
 * on this GPU for every rank.  out[i] = sum over ranks of local[i], added in rank order (bitwise
 * identical on all ranks).  Collective: every rank must launch it the same number of times.
 * Protocol: flag-in-data (every double = two 8-byte words {half, epoch tag} stored into the peers'
 * buffers, the receiver polls its own buffer until the tags match); waits are bounded (~4 s of SM
 * clocks: a time-out writes the epoch to byte 8 of the rank's buffer instead of hanging the GPU).
 * ------------------------------------------------------------------------------------- */
long long mvd_peer_allreduce_buffer_bytes(int world, int nmax);
int mvd_peer_allreduce_f64(const double* local, double* out, int n, const unsigned long long* peers,
                           int rank, int world, int nmax, void* stream);

/* ---------------------------------------------------------------------------------------
 * Measurement helpers (bench.py's live cost-volume roofline): CUDA timing events that also
 * work INSIDE a captured CUDA graph.  mvd_event_record with external != 0 uses
 * cudaEventRecordExternal, i.e. the record becomes an event-record NODE when the stream is
 * being captured and fires at every replay; mvd_event_elapsed_ms is cudaEventElapsedTime.
 * ------------------------------------------------------------------------------------- */
void* mvd_event_create(void);
int mvd_event_record(void* event, void* stream, int external);
int mvd_event_elapsed_ms(void* start, void* stop, float* ms);
int mvd_event_destroy(void* event);

#ifdef __cplusplus
}
#endif
#endif /* MOVEDEPTH_B200_H_ */
