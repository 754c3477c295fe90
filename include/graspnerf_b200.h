/* graspnerf_b200 C ABI  --  libgraspnerf_b200.so
 *
 * Drop-in boundary for GraspNeRF's generalizable-NeRF volumetric TSDF hot path (reference: PKU-EPIC/GraspNeRF,
 * src/nr/network).  The reference has NO native/FFI layer for this path (it is pure PyTorch);
 * these entry points are what a binding for the path binds: each one replaces a chain of reference torch ops,
 * cited per function.  The reference-side binding (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated; the caller owns all buffers (kernels never allocate);
 *   - all launches go to the caller's `stream` (a cudaStream_t passed as void*), no host synchronisation;
 *   - return value: 0 = ok, <0 = argument error, >0 = cudaError_t of the launch; no exceptions cross the ABI;
 *   - re-entrant; the only process-global state is an idempotent per-device cache of kernel attributes; one process
 *     per GPU under data parallelism.
 *
 * Data layout in HBM:
 *   feature maps  : channels-last  [B, V, fh, fw, 32]
 *   images        : RGBA-interleaved [B, V, H, W, 4] (one bilinear tap = one 16-byte texel)
 *   record  `rec` : [B, N, V, 72]  ray_feats32 | dir_diff4 | rgb3, depth | img_feats32 ; n = ray*dn + sample
 *   per-point `pt`: [B, N, 2]      nvalid, view bit mask
 *   `pooled`      : [B, N, 68]     K2a output: mean32 | var32 | mean_v(w), nvalid, 0, 0
 */
#ifndef GRASPNERF_B200_H
#define GRASPNERF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* K1: fused project - sample.
 * Replaces project_points_dict (render_ops.py:82-144), get_img_feats (renderer.py:80-88), get_dir_diff
 * (aggregate_net.py:11-17) and the per-point valid-view count / mask (ibrnet.py:466,490).  volume_mode=1: points are the voxel centres of utils/field_utils.py:12-27 plus bbox_min,
 * in sample_volume's order (renderer.py:166-170).  volume_mode=0: explicit points (RGB head, render_ops.py:27-39). */
typedef struct GnK1Params {
    const void* imgs;         /* [B,V,H,W,4] RGBA-interleaved (A unused): fp32 in [0,1], or uint8 when img_u8 = 1 */
    const float* img_feats;   /* [B,V,fh,fw,32] channels-last, or ray_feats + 32 when feat_stride == 64 */
    const float* ray_feats;   /* [B,V,fh,fw,32] channels-last, or the fused [B,V,fh,fw,64] buffer (ray_feats 32 | img_feats 32 per texel) */
    const float* KRt;         /* [B,V,3,4]  K @ [R|t]  (render_ops.py:94) */
    const float* cam;         /* [B,V,3]    camera centres -R^T t (render_ops.py:112) */
    const float* axis;        /* [R] voxel-centre table (volume mode) */
    const float* bbox_min;    /* [B,3] (volume mode) */
    const float* pts;         /* [B,N,3] (ray mode) */
    const float* que_dir;     /* [B,N/dn,3] unit query direction per ray (ray mode) */
    float* rec;               /* out [B,N,V,72] */
    float* pt;                /* out [B,N,2]: nvalid (float), view bit mask (uint32 bits) */
    int* dbg_feat_idx;        /* optional out [B,N,V,2] int32 (x0,y0) feature-map corner indices, or NULL */
    int B, V, H, W, fh, fw;
    int R;                    /* grid resolution (volume mode) */
    int N;                    /* points per scene */
    int dn;                   /* samples per ray */
    int volume_mode;
    int tiles_per_scene;      /* filled in by the launcher */
    int feat_stride;          /* floats per feature-map texel: 32 (two maps, 0 means 32) or 64 (fused buffer: one address per bilinear tap) */
    int img_u8;               /* 1: imgs holds uint8 texels; the kernel divides by 255 exactly as color_map_forward does (main.py:170,
                                 utils/base_utils.py:492-493), so the planner's PNG bytes cross PCIe as bytes */
    int* valid_count;         /* optional [B,V] int32, ACCUMULATED (caller zeroes it): number of valid projections per view - the
                                 numerator of the reference's "!! too low ratio" diagnostic (renderer.py:174-176) without a host
                                 synchronisation; NULL = not counted */
} GnK1Params;

int gn_k1_forward(const GnK1Params* params, void* stream);

/* K2 weight blob: all head weights packed k-major ([in][out_padded]) in one fp32 buffer.  The table of entries is
 * owned by the library; the host packer enumerates it (so offsets can never disagree). */
int gn_weight_entry_count(void);
int gn_weight_entry(int idx, const char** name, int* offset, int* rows, int* cols, int* cols_padded);
int gn_weight_blob_floats(void);

/* K2a: per-(point,view) head + cross-view pooling.
 * Replaces ray_dir_fc + add and the mask-weighted mean/var (ibrnet.py:457-471), MixtureLogisticsDistDecoder.forward + compute_prob (dist_decoder.py:99-142, via predict_proj_ray_prob
 * renderer.py:62-78), prob_embed (aggregate_net.py:47-54) and IBRNetWithNeuRayNeus.forward lines 466-484 + 507-511
 * (ibrnet.py): neuray_fc, weighted mean/var, base_fc, vis_fc, vis_fc2, pooling, and (with_rgb) rgb_fc + softmax blend.
 * que_dists: NULL -> fixed +-0.005 interval (volume mode, dist_decoder.py:121-124); else [B,N] normalised
 * inverse-depth spacings (render_ops.py:46-52). */
typedef struct GnK2aParams {
    const float* rec;          /* [B,N,V,72] */
    const float* pt;           /* [B,N,2] */
    const float* weights;      /* blob */
    const float* depth_range;  /* [B,V,2] */
    const float* que_dists;    /* [B,N] or NULL */
    float* pooled;             /* out [B,N,68] */
    float* colors;             /* out [B,N,4] (rgb, 0) or NULL */
    float* dbg_rows;           /* optional out [B,N,V,8]: hit, vis, w0, vis2, x0, x1, pe0, pe1 ; or NULL */
    /* tensor-core path only: */
    const void* tc_const;      /* buffer of gn_k2a_tc_const_bytes() bytes filled by gn_k2a_tc_prepare() */
    float* tok;                /* optional out [B,N,20]: geometry_fc output (ibrnet.py:487-489) 16 | nvalid,0,0,0 ; or NULL */
    const float* axis;         /* [R]   (tok, volume mode) */
    const float* bbox_min;     /* [B,3] (tok, volume mode) */
    const float* pts;          /* [B,N,3] (tok, ray mode) */
    int B, N, V, dn;
    int with_rgb;              /* 1: also evaluate rgb_fc + blend into `colors` */
    int R, volume_mode;        /* (tok) */
    int* status;               /* optional sticky flag word (tensor-core path): bit 0 is OR-ed in when a row's hit / visibility or a
                                  point's pooled features / tokens are not finite - the fp16 hi/lo operand split overflows for
                                  activations >= 65504, which the fp32 reference would survive; NULL = no check */
} GnK2aParams;
int gn_k2a_forward_tc(const GnK2aParams* params, void* stream);   /* THE product kernel: tcgen05 / TMEM (fp16 hi/lo split, 3 MMAs per product; 160 TMEM columns per tile, three tiles per SM) */
int gn_k2a_forward(const GnK2aParams* params, void* stream);      /* fp32 CUDA-core implementation of the same math: on-GPU cross-check for tests, never selected by the host code */
int gn_k2a_tc_const_bytes(void);
int gn_k2a_tc_prepare(const float* weights, void* tc_const, void* stream);   /* fp32 blob -> fp16 hi/lo operand images + small constants */

/* K2b: per-ray geometry head.
 * Replaces ibrnet.py:485-495: embed (neus.py:21-66), geometry_fc, + pos_encoding (ibrnet.py:437-445), MultiHeadAttention
 * over the dn samples (ibrnet.py:52-102, query-row mask), LayerNorm, out_geometry_fc, clip(-1,1), invalid -> 1.0.
 * volume_mode=1: writes volume[B,R,R,R] with the final z flip of renderer.py:198 fused; else sdf[B,N].
 * grad (optional, [B,N,3]): d(sum sdf)/d pts  (ibrnet.py:497-504), hand-derived reverse pass. */
typedef struct GnK2bParams {
    const float* pooled;       /* [B,N,68]  (full path: embed + geometry_fc + attention; required when grad != NULL) */
    const float* tok;          /* [B,N,20]  (attention-only path: geometry_fc already done by K2a-TC); used when pooled == NULL */
    const float* weights;      /* blob */
    const float* axis;         /* [R] (volume mode) */
    const float* bbox_min;     /* [B,3] (volume mode) */
    const float* pts;          /* [B,N,3] (ray mode) */
    const float* pos_table;    /* [dn,16] sinusoid table (ibrnet.py:437-445) */
    float* sdf;                /* out: volume [B,R,R,R] or sdf [B,N] */
    float* grad;               /* out [B,N,3] or NULL */
    int B, N, dn, R, volume_mode;
} GnK2bParams;
int gn_k2b_forward(const GnK2bParams* params, void* stream);

/* K3: NeuS alpha + alpha compositing for the RGB head.
 * Replaces _get_alpha_from_sdf (aggregate_net.py:105-123), alpha_values2hit_prob (render_ops.py:72-80), the colour /
 * depth sums (renderer.py:105-106,136) and the eikonal partial sums (aggregate_net.py:139). */
typedef struct GnK3Params {
    const float* sdf;          /* [B,rn,dn] */
    const float* grad;         /* [B,rn,dn,3] */
    const float* colors;       /* [B,rn,dn,4] */
    const float* que_dir;      /* [B,rn,3] */
    const float* depth;        /* [B,rn,dn] */
    float inv_s;               /* exp(10*variance) clipped (aggregate_net.py:107) */
    float cos_anneal_ratio;
    float* alpha;              /* out [B,rn,dn] */
    float* hit_prob;           /* out [B,rn,dn] */
    float* pixel_colors;       /* out [B,rn,3] */
    float* render_depth;       /* out [B,rn] */
    float* eik_partial;        /* out [B,rn]: sum_d (|grad|-1)^2 */
    int B, rn, dn;
} GnK3Params;
int gn_k3_composite(const GnK3Params* params, void* stream);

/* K3 samplers (index tables bit-exact): sample_depth (render_ops.py:146-170, deterministic branch) and
 * sample_fine_depth (render_ops.py:172-229; u supplied by the caller: midpoints in eval, uniform randoms in train). */
int gn_k3_coarse_depths(const float* depth_range /*[B,2]*/, float* depth /*[B,rn,dn]*/, int B, int rn, int dn, void* stream);
int gn_k3_fine_depths(const float* depth /*[B,rn,dn]*/, const float* hit_prob /*[B,rn,dn]*/, const float* depth_range /*[B,2]*/,
                      const float* u /*[B,rn,fdn]*/, float* fine_depth /*[B,rn,fdn] sorted*/, int64_t* inds /*[B,rn,fdn] or NULL*/,
                      int B, int rn, int dn, int fdn, void* stream);

/* Ray set-up of the RGB head in one launch: coords2rays (render_ops.py:4-25), depth2points (27-39) and depth2inv_dists
 * (46-52).  poses are the QUERY views' world->camera [R|t]; depth is the per-ray sample table (gn_k3_coarse_depths /
 * gn_k3_fine_depths).  Outputs feed gn_k1_forward (pts, que_dir) and gn_k2a_forward_tc (inv_dists = que_dists). */
typedef struct GnRaySetupParams {
    const float* coords;       /* [B,rn,2] pixel (x,y) */
    const float* poses;        /* [B,3,4] */
    const float* Ks;           /* [B,3,3] */
    const float* depth;        /* [B,rn,dn] */
    const float* depth_range;  /* [B,2] */
    float* pts;                /* out [B,rn*dn,3] */
    float* que_dir;            /* out [B,rn,3] unit, pointing from the sample back to the camera (render_ops.py:37) */
    float* inv_dists;          /* out [B,rn*dn] spacings in normalised inverse depth, last one 1e6 */
    float* centers;            /* optional out [B,rn,3]: ray origins (coords2rays' first result), or NULL */
    float* dirs;               /* optional out [B,rn,3]: un-normalised ray directions (coords2rays' second result), or NULL */
    int B, rn, dn;
} GnRaySetupParams;
int gn_k3_ray_setup(const GnRaySetupParams* params, void* stream);

/* Layout glue around the encoders (one launch each instead of 2-4 strided torch copies):
 * gn_k6_fuse_features: the two NCHW maps the encoders produce, [planes,32,h*w] each, -> the fused channels-last buffer
 *   [planes,h*w,64] (ray_feats | img_feats per texel) that gn_k1_forward gathers from (ray_feats / img_feats / feat_stride = 64).
 * gn_k6_images_u8: uint8 images [V,H,W,C] (C = 3 or 4) -> fp32 [V,3,H,W] = u8 / 255 (color_map_forward, main.py:170, a true
 *   division) and, if out_rgba is not NULL, the uint8 RGBA texels [V,H,W,4] for gn_k1_forward's img_u8 mode. */
int gn_k6_fuse_features(const float* ray_feats, const float* img_feats, float* out, int planes, int hw, void* stream);
int gn_k6_images_u8(const unsigned char* in, float* out_f, unsigned char* out_rgba, int V, int H, int W, int C, void* stream);

/* Depth-mean head (renderer.py:222-266, predict_mean_for_depth_loss; runs in every eval forward, renderer.py:288-289):
 * bilinear sample of ray_feats at `num` pixels per reference view (ops.py:14-34: grid_sample, border padding) and
 * MixtureLogisticsDistDecoder.predict_mean (dist_decoder.py:148-150) of the coarse and, optionally, the fine decoder.
 * feats is addressed through element strides, so NCHW [V,32,fh,fw] and channels-last maps both work.
 * w_*: the six tensors of mean_decoder in nn.Linear layout: W0 [32,32], b0 [32], W1 [32,32], b1 [32], W2 [2,32], b2 [2]. */
typedef struct GnDepthMeanParams {
    const float* feats;            /* ray_feats of the reference views */
    const long long* coords;       /* [V,num,2] int64, the reference's (row, col) pairs used as (x, y) */
    const float* w_coarse[6];
    const float* w_fine[6];        /* all NULL: coarse decoder only */
    float* mean;                   /* out [V,num,2] */
    float* mean_fine;              /* out [V,num,2] (with w_fine) */
    long long stride_v, stride_c, stride_y, stride_x;
    int V, num, H, W, fh, fw, align_corners;
} GnDepthMeanParams;
int gn_k3_depth_mean(const GnDepthMeanParams* params, void* stream);

/* Grasp post-processing on the device (the step after the path in GraspNeRFPlanner.__call__, main.py:23-84,202-203):
 * `process` = gaussian_filter(qual, sigma 1, 'nearest') + TSDF band mask (binary_dilation of the outside voxels, 2
 * iterations, restricted to the band) + width limits; `select` = threshold, 4^3 maximum-filter NMS ('reflect'), ordered
 * compaction of the surviving voxels.  Volumes are [R,R,R] fp32 (rot [4,R,R,R]); one scene per call.
 * scratch: 3*R^3 floats.  count: int32 number of grasps found (may exceed max_grasps: only the first max_grasps rows are
 * written).  The filtered volume is bit-identical to scipy's (float64 accumulation in scipy's tap order). */
typedef struct GnGraspPostParams {
    const float* tsdf;         /* [R,R,R] */
    const float* qual;         /* [R,R,R] */
    const float* rot;          /* [4,R,R,R] */
    const float* width;        /* [R,R,R] */
    float* qual_out;           /* out [R,R,R]: quality volume after process() (before select's threshold / NMS) */
    float* scratch;            /* [3*R^3] */
    float* grasps;             /* out [max_grasps,9]: i, j, k, score, rot0..3, width (argwhere order, main.py:68-73) */
    int* count;                /* out [1] */
    float sigma, min_width, max_width, tsdf_thres_high, tsdf_thres_low, threshold;
    int R, max_filter_size, max_grasps;
} GnGraspPostParams;
int gn_k4_grasp_post(const GnGraspPostParams* params, void* stream);

/* K5: the VGN 3-D ConvNet that consumes the TSDF volume (src/gd/networks.py:39-97, called at renderer.py:323-330): three
 * strided encoder convolutions, three decoder convolutions with nearest x2 upsampling in between (folded exactly into 8
 * parity classes of 3x3x3 kernels on the low-resolution grid) and the three 5^3 heads (sigmoid / F.normalize / identity)
 * as seven direct fp32 convolutions.  weights: blob prepared by the host (graspnerf_b200.weights.pack_vgn, layout from
 * gn_vgn_layer_info); out per scene: qual [R^3] | rot [4,R^3] | width [R^3] (6 consecutive volumes). */
typedef struct GnVgnParams {
    const float* volume;       /* [B,R,R,R] TSDF volume (K2b's output) */
    const float* weights;      /* [gn_vgn_blob_floats()] */
    float* workspace;          /* [gn_vgn_workspace_floats(R)] */
    float* out;                /* [B] scenes, out_scene_stride floats apart, 6*R^3 floats each */
    int B, R;
    int out_scene_stride;
} GnVgnParams;
int gn_vgn_forward(const GnVgnParams* params, void* stream);
int gn_vgn_layer_info(int layer, int* cin, int* cout, int* cout_t, int* taps, int* classes, int* w_offset, int* b_offset);
int gn_vgn_blob_floats(void);
int gn_vgn_workspace_floats(int R);

/* K6: fused element-wise stages of the 2-D encoders (src/nr/network/ops.py:78-230: ResUNetLight / BasicBlock / conv /
 * upconv; init_net.py:8-35; vis_encoder.py:6-21), inference only.  One launch per layer replaces reflection pad +
 * InstanceNorm (+ affine repeat) + activation + residual add:
 *   out = reflect_pad( act( IN(x)*gamma+beta [+ res | + IN(res)*res_gamma+res_beta] ), pad )
 * x / res may themselves be interiors of padded tensors (x_pad / res_pad = their border width).  NCHW fp32. */
typedef struct GnNormActPadParams {
    const float* x;            /* [N,C,H+2*x_pad,W+2*x_pad] */
    const float* gamma;        /* [C] or NULL (no normalisation of x) */
    const float* beta;         /* [C] or NULL */
    const float* res;          /* [N,C,H+2*res_pad,W+2*res_pad] or NULL */
    const float* res_gamma;    /* [C] or NULL: instance-normalise the residual too (BasicBlock downsample branch, ops.py:96-124) */
    const float* res_beta;
    float* out_padded;         /* [N,C,H+2*pad,W+2*pad] or NULL */
    float* out_unpadded;       /* [N,C,H,W] or NULL */
    int N, C, H, W;
    int pad, x_pad, res_pad;
    int act;                   /* 0 none, 1 ReLU, 2 ELU */
    float eps;
    int x_splits;              /* > 1: x is the SUM of x_splits partial tensors (split-K output of gn_k7_conv_forward), x_split_stride
                                  floats apart; 0 / 1: a plain tensor */
    long long x_split_stride;
} GnNormActPadParams;
int gn_k6_norm_act_pad(const GnNormActPadParams* params, void* stream);
/* F.interpolate(scale_factor=2, bilinear, align_corners=True) (ops.py:142-150) + reflection pad: x [planes,H,W] -> out [planes,2H+2pad,2W+2pad] */
int gn_k6_upsample2x_pad(const float* x, float* out, int planes, int H, int W, int pad, void* stream);

/* K7: nn.Conv2d of the 2-D encoders (ops.py:78-230, init_net.py:21-26, vis_encoder.py:9-14) as an implicit GEMM on tcgen05,
 * inference only.  `in` is NCHW fp32 and ALREADY padded by its producer (gn_k6_*), so this is a "valid" convolution:
 * out[img][co][oy][ox] = bias[co] + sum_k in[img-base + (oy*stride)*Wp + ox*stride + koff[k]] * W[co][k], k = (ci, dy, dx).
 * wimg: fp16 hi/lo operand images per 32-wide k chunk, koff: int32 [Kpad] (both built by the host, graspnerf_b200.ops). */
typedef struct GnConvParams {
    const float* in;           /* [Nimg,Cin,Hp,Wp] */
    const void* wimg;          /* [Kpad/32][2 (hi,lo)][4][Npad][8] fp16 */
    const int* koff;           /* [Kpad] input offsets of k relative to a pixel's window origin (0 for padded k) */
    const float* bias;         /* [Cout] or NULL */
    float* out;                /* [Nimg,Cout,Ho,Wo] */
    long long M;               /* filled in by the launcher: Nimg*Ho*Wo */
    int Nimg, Cin, Hp, Wp, Cout, Npad, Ho, Wo, stride, Kpad;
    int ksplit;                /* 0 / 1: one CTA per 128-pixel tile runs all of K.  S > 1: S CTAs per tile, each a contiguous range of the
                                  k chunks, writing S partial outputs [S][Nimg,Cout,Ho,Wo] (bias in partial 0); the consumer
                                  (gn_k6_norm_act_pad, x_splits = S) sums them in a fixed order - deterministic, no atomics */
} GnConvParams;
int gn_k7_conv_forward(const GnConvParams* params, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Backward (training) entry points of the volume path: d volume -> d weights, d feature maps.  First order only.
 * The reference gets these from torch autograd through renderer.py:164-199; here each forward kernel has a hand-derived
 * reverse kernel that RECOMPUTES its forward from the saved inputs (rec / pt from K1, pooled from K2a) and accumulates
 * with atomics (outputs must be zero-initialised by the caller).  Weight gradients are produced in the blob layout, in
 * DOUBLE precision (same element offsets as `weights`; fused composites nfc.* / bf.wpc / bf.b0c are not touched). */
typedef struct GnK2bBwdParams {
    const float* pooled;       /* [B,N,68] saved K2a (SIMT) output */
    const float* weights;      /* blob */
    const float* axis;         /* [R] (volume mode) */
    const float* bbox_min;     /* [B,3] (volume mode) */
    const float* pts;          /* [B,N,3] (ray mode) */
    const float* pos_table;    /* [dn,16] */
    const float* d_sdf;        /* upstream gradient, SAME layout as the forward output: volume [B,R,R,R] or sdf [B,N] */
    float* d_pooled;           /* out [B,N,68]: gradient of mean32 | var32 | wmean (entries 65..67 written as 0) */
    double* d_weights;         /* accumulated (atomicAdd, fp64: the sum over ~10^5 rows loses no bits) [gn_weight_blob_floats()] */
    int B, N, dn, R, volume_mode;
} GnK2bBwdParams;
int gn_k2b_backward(const GnK2bBwdParams* params, void* stream);

typedef struct GnK2aBwdParams {
    const float* rec;          /* [B,N,V,72] saved K1 output */
    const float* pt;           /* [B,N,2] */
    const float* weights;      /* blob */
    const float* depth_range;  /* [B,V,2] */
    const float* que_dists;    /* [B,N] or NULL (volume mode) */
    const float* d_pooled;     /* [B,N,68] from gn_k2b_backward */
    float* d_rec;              /* out [B,N,V,64]: gradient of the record's ray_feats (32) | img_feats (32) entries */
    double* d_weights;         /* accumulated (atomicAdd, fp64: the sum over ~10^5 rows loses no bits) [gn_weight_blob_floats()] */
    const float* d_colors;     /* [B,N,4] upstream gradient of K2a's blended colours (rgb, pad), or NULL (volume path): adds the
                                  reverse of rgb_fc + softmax blend (ibrnet.py:507-511) */
    int B, N, V, dn;
} GnK2aBwdParams;
int gn_k2a_backward(const GnK2aBwdParams* params, void* stream);

typedef struct GnK1BwdParams {
    const float* KRt;          /* [B,V,3,4] */
    const float* axis;         /* [R] (volume mode) */
    const float* bbox_min;     /* [B,3] (volume mode) */
    const float* pts;          /* [B,N,3] (ray mode) */
    const float* d_rec;        /* [B,N,V,64] from gn_k2a_backward */
    float* d_img_feats;        /* accumulated [B,V,fh,fw,32] channels-last */
    float* d_ray_feats;        /* accumulated [B,V,fh,fw,32] channels-last */
    int B, V, H, W, fh, fw, R, N, volume_mode;
} GnK1BwdParams;
int gn_k1_backward(const GnK1BwdParams* params, void* stream);

const char* gn_version(void);
int gn_sizeof_k1_params(void);
int gn_sizeof_k2a_params(void);
int gn_sizeof_k2b_params(void);
int gn_sizeof_k3_params(void);
int gn_sizeof_k2b_bwd_params(void);
int gn_sizeof_k2a_bwd_params(void);
int gn_sizeof_k1_bwd_params(void);
int gn_sizeof_ray_setup_params(void);
int gn_sizeof_depth_mean_params(void);
int gn_sizeof_grasp_post_params(void);
int gn_sizeof_vgn_params(void);
int gn_sizeof_norm_act_pad_params(void);
int gn_sizeof_conv_params(void);

#ifdef __cplusplus
}
#endif
#endif
