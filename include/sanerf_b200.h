/*
 * sanerf_b200.h -- C ABI of libsanerf_b200.so: hand-written sm_100a CUDA for the SANeRF-HQ
 * volumetric-render hot path.
 *
 * This is the drop-in boundary (SURVEY.md 8b "B-native").  Part 1 replaces, one for one, the
 * eight functions the reference exports from its three pybind `_backend` modules; Part 2 adds
 * the fused per-ray render entry point that replaces the ~250 ATen + encoder launches of
 * `NeRFRenderer.run` (reference nerf/renderer.py:221-385).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current device
 *     unless documented otherwise; all buffers are contiguous;
 *   - the CALLER allocates every output (as the reference's Python does with torch.empty /
 *     torch.zeros); the library never allocates, frees or synchronises;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t; NULL = legacy default
 *     stream, which is what the reference's `<<<grid, block>>>` launches use);
 *   - return value: 0 on success; a positive cudaError_t from the launch; or a negative
 *     SANERF_E_* code for an argument the reference would have rejected with a
 *     std::runtime_error / TORCH_CHECK (unsupported D or C, null pointer, bad dtype).
 *     `sanerf_error_string` explains any of them.  The Python layer turns non-zero into
 *     RuntimeError, matching the reference's pybind behaviour.
 *   - every tensor in the interface is fp32 (the reference forces fp16 off, main.py:217).  Arithmetic: the encoder operators
 *     of Part 1 and all element-wise / sampling / compositing steps of Part 2 are plain fp32 in the reference's operation order
 *     (grid forward bit-exact against the reference kernel).  The dense layers of Part 2 run on the 5th-gen tensor cores in
 *     SPLIT precision with fp32 accumulation: proposal MLPs 3xTF32 (operands hi+lo, products hi*hi + hi*lo + lo*hi, relative
 *     error 2^-22 -- they decide sample placement and the sample_pdf index buffers); grid_mlp, mask_mlp, samvit_mlp bf16 hi+lo
 *     operands (16 significant bits, the lo*lo product dropped, relative error 2^-18 per product); view_mlp fp32 CUDA cores.
 *     Tested against the reference's own fp32 GPU path at 1e-3 relative (tests/test_ref_gpu_frames.py; DESIGN.md section 5
 *     lists the measured margins).
 */
#ifndef SANERF_B200_H
#define SANERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *sanerf_stream_t; /* cudaStream_t */

#define SANERF_ABI_VERSION 2

#define SANERF_OK 0
#define SANERF_E_NULL (-1)        /* required pointer is NULL */
#define SANERF_E_DIM (-2)         /* GridEncoding: D must be 2, 3, 4 or 5 (gridencoder.cu:409) */
#define SANERF_E_CHANNELS (-3)    /* GridEncoding: C must be 1, 2, 4, 8, 16 or 32 (gridencoder.cu:392) */
#define SANERF_E_DEGREE (-4)      /* SH degree must be in [1, 8] (sphere_harmonics.py:70) */
#define SANERF_E_CONFIG (-5)      /* fused render: model/shape combination not supported */
#define SANERF_E_SMEM (-6)        /* fused render: device cannot grant the shared memory needed */

int sanerf_abi_version(void);
const char *sanerf_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Part 1 -- encoder operators (reference pybind exports)
 * ---------------------------------------------------------------------------------------- */

/* replaces grid_encode_forward  (gridencoder/src/gridencoder.h:12, gridencoder.cu:467-490, kernel :82-249)
 * inputs [B,D] in [0,1]; embeddings [sO,C]; offsets [L+1] int32; outputs [L,B,C];
 * dy_dx [B, L*D*C] or NULL; S = log2(per_level_scale); H = base resolution;
 * gridtype 0 hash / 1 tiled; interp 0 linear / 1 smoothstep.  Only levels < max_level are written. */
int sanerf_grid_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                               float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                               uint32_t max_level, float S, uint32_t H, float *dy_dx, uint32_t gridtype,
                               int align_corners, uint32_t interp, sanerf_stream_t stream);

/* Fused variant used by GridEncoder.forward: takes raw positions in [-bound,bound], applies the
 * (x+bound)/(2*bound) mapping of gridencoder/grid.py:156 in-kernel and writes the [B, L*C]
 * layout directly (removes the permute copy of grid.py:63).  Same arithmetic as the above.
 * bound <= 0 means the positions are already in [0,1] (no mapping). */
int sanerf_grid_encode_forward_fused(const float *positions, float bound, const float *embeddings,
                                     const int32_t *offsets, float *outputs_BLC, uint32_t B, uint32_t D,
                                     uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                     uint32_t gridtype, int align_corners, uint32_t interp,
                                     sanerf_stream_t stream);

/* replaces grid_encode_backward (gridencoder.h:13, gridencoder.cu:492-522, kernels :252-378)
 * grad [L,B,C]; grad_embeddings [sO,C] zero-filled by the caller, accumulated with atomics;
 * when dy_dx != NULL also writes grad_inputs [B,D].  `embeddings` is unused (kept for ABI symmetry). */
int sanerf_grid_encode_backward(const float *grad, const float *inputs, const float *embeddings,
                                const int32_t *offsets, float *grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                const float *dy_dx, float *grad_inputs, uint32_t gridtype,
                                int align_corners, uint32_t interp, sanerf_stream_t stream);

/* Backward of the fused variant: grad is [B, L*C]; positions raw in [-bound,bound]. */
int sanerf_grid_encode_backward_fused(const float *grad_BLC, const float *positions, float bound,
                                      const int32_t *offsets, float *grad_embeddings, uint32_t B,
                                      uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S,
                                      uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                      sanerf_stream_t stream);

/* replaces grad_total_variation (gridencoder.h:15, gridencoder.cu:525-668); in place on grad */
int sanerf_grad_total_variation(const float *inputs, const float *embeddings, float *grad,
                                const int32_t *offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                                uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                sanerf_stream_t stream);

/* replaces grad_weight_decay (gridencoder.h:16, gridencoder.cu:670-713); B = rows of embeddings */
int sanerf_grad_weight_decay(const float *embeddings, float *grad, const int32_t *offsets, float weight,
                             uint32_t B, uint32_t C, uint32_t L, sanerf_stream_t stream);

/* Device-evaluated level table: res[l] = (uint32)ceil(exp2f(l*S)*H) exactly as the reference kernels
 * compute it on the GPU (gridencoder.cu:133).  out_res: DEVICE uint32[L]. */
int sanerf_grid_level_resolutions(uint32_t *out_res, uint32_t L, float S, uint32_t H, sanerf_stream_t stream);

/* replaces sh_encode_forward (shencoder/src/shencoder.h:9, shencoder.cu:400-417, kernel :27-355)
 * inputs [B,3] (unit vectors); outputs [B,degree^2]; dy_dx [B,3*degree^2] or NULL. */
int sanerf_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t degree,
                             float *dy_dx, sanerf_stream_t stream);

/* replaces sh_encode_backward (shencoder.h:10, shencoder.cu:358-382, 419-438): grad_inputs[b,d] += sum_ch grad*dy_dx */
int sanerf_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t degree,
                              const float *dy_dx, float *grad_inputs, sanerf_stream_t stream);

/* replaces freq_encode_forward (freqencoder/src/freqencoder.h:7, freqencoder.cu:30-58, 97-111)
 * outputs [B,C], C = D + 2*D*deg. */
int sanerf_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                               float *outputs, sanerf_stream_t stream);

/* replaces freq_encode_backward (freqencoder.h:10, freqencoder.cu:63-94, 114-128) */
int sanerf_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg,
                                uint32_t C, float *grad_inputs, sanerf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Part 2 -- fused render (replaces NeRFRenderer.run, eval / no-grad; perturb=True with caller-drawn random numbers)
 * ---------------------------------------------------------------------------------------- */

#define SANERF_MAX_LEVELS 16
#define SANERF_MAX_PEERS 8

/* One multiresolution hash grid (a reference GridEncoder, gridencoder/grid.py:102-146). */
typedef struct {
    const float *embeddings;                /* [rows, C] device */
    uint32_t num_levels;                    /* L <= SANERF_MAX_LEVELS */
    uint32_t level_dim;                     /* C: 2 (grid, proposals) or 8 (s_grid, m_grid) */
    uint32_t offset[SANERF_MAX_LEVELS + 1]; /* row offsets, host-side rule grid.py:124-135 */
    uint32_t res[SANERF_MAX_LEVELS];        /* kernel-side resolution per level (gridencoder.cu:133) */
} sanerf_grid_t;

/* nn.Linear weights are passed in the checkpoint layout: [out, in] row-major fp32, y = x W^T (+b). */
typedef struct {
    sanerf_grid_t prop_grid[2];     /* prop_encoders.{0,1} */
    const float *prop_w0[2];        /* prop_mlp.i.net.0.weight [16, 2*L] */
    const float *prop_w1[2];        /* prop_mlp.i.net.1.weight [1, 16]   */
    sanerf_grid_t grid;             /* grid */
    const float *grid_w[3];         /* grid_mlp.net.{0,1,2}.weight  [Hg,2L] [Hg,Hg] [16,Hg] */
    uint32_t grid_hidden;           /* Hg: 64 (default) or 16 (config #1) */
    const float *view_w[3];         /* view_mlp.net.{0,1,2}.weight  [Hv,31] [Hv,Hv] [3,Hv] */
    uint32_t view_hidden;           /* Hv: 32 (default) or 16 */
    /* optional SAM feature head (with_sam): NULL embeddings = absent */
    sanerf_grid_t s_grid;
    const float *sam_w[5];          /* samvit_mlp.0.net.{0..4}.weight */
    const float *sam_b[5];          /* samvit_mlp.0.net.{0..4}.bias   */
    const float *sam_ln_w, *sam_ln_b; /* samvit_mlp.1.{weight,bias} (LayerNorm 256, eps 1e-5) */
    /* optional object head (return_mask): NULL embeddings = absent */
    sanerf_grid_t m_grid;
    const float *mask_w[3];         /* mask_mlp.0.net.{0,1,2}.weight [256,143] [256,256] [n_inst,256] */
    uint32_t n_inst;
    float aabb[6];                  /* aabb_train / aabb_infer as selected by the caller (renderer.py:232) */
    float min_near;                 /* opt.min_near */
    float grid_bound;               /* 2 when opt.contract else opt.bound (renderer.py:152-155) */
    uint32_t contract;              /* opt.contract */
    uint32_t last_sample_opaque;    /* opt.background == 'last_sample' (renderer.py:313-315) */
    const float *u65, *u33;         /* DEVICE copies of torch.linspace(.5/T, 1-.5/T, T) for T=65, 33 (renderer.py:97) */
} sanerf_model_t;

typedef struct {
    const float *rays_o;            /* [N,3] */
    const float *rays_d;            /* [N,3], unnormalised */
    uint32_t N;
    const float *cam_near_far;      /* NULL, or [1,2] / [N,2] (renderer.py:233-235) */
    uint32_t cam_near_far_rows;     /* 1 or N */
    const float *bg_color;          /* NULL -> bg_scalar; else [3] or [N,3] */
    uint32_t bg_rows;               /* 1 or N */
    float bg_scalar;                /* used when bg_color == NULL (reference default 1) */
    /* outputs (device, caller-allocated) */
    float *image;                   /* [N,3] */
    float *depth;                   /* [N] */
    float *weights_sum;             /* [N] */
    float *sam_in;                  /* optional [N, 8*Ls+35] (163): cat[f_sam, f_image, image, depth], the samvit_mlp input
                                       (renderer.py:361-367); requires model->s_grid */
    float *mask_in;                 /* optional [N,32, 8*Lm+15] (143): per-sample cat[m_grid(x), geo_feat], the mask_mlp
                                       input (renderer.py:304-305, 378); requires model->m_grid */
    uint32_t mask_in_tiled;         /* 0: row-major as above; 1: the same values tile-transposed [ceil(N*32/128)][K][128]
                                       (row = ray*32+sample); 2: record mode for sanerf_mask_head, which gathers m_grid itself:
                                       mask_in is [ceil(N*32/128)][18][128] = per sample the point in [0,1]^3 (3) and geo_feat
                                       (15), model->m_grid is not read.  N must start at a multiple of 4 rays for 1 and 2. */
    /* optional debug / parity taps (NULL to skip) */
    int16_t *inds0;                 /* [N,65]  searchsorted result of the 1st sample_pdf */
    int16_t *inds1;                 /* [N,33]  of the 2nd */
    float *weights2;                /* [N,32]  final-stage weights */
    float *sigma2;                  /* [N,32]  final-stage sigma */
    float *bins2;                   /* [N,33]  final-stage bins (normalised) */
    float *f_image;                 /* [N,31]  composited deferred-shading feature */
    /* optional in-kernel ray generation (replaces the full-image branch of nerf/utils.py::get_rays :262-287):
     * when cam_w > 0, rays_o / rays_d may be NULL and ray n is the pixel with linear index q = cam_ray0 + n (col = q % cam_w,
     * row = q / cam_w) of a pinhole camera: dirs = ((col+.5-cx)/fx, -(row+.5-cy)/fy, -1), rays_d = R dirs (unnormalised), rays_o = t. */
    uint32_t cam_w, cam_ray0;
    float cam_intrinsics[4];        /* fx, fy, cx, cy */
    float cam_pose[12];             /* rows 0..2 of the 4x4 cam2world matrix, row-major: [R | t] */
    /* optional traversal hint: the N rays are a row-major image (block) of width tile_w (N % (4*tile_w) == 0, tile_w % 4 == 0):
     * the kernel then walks 4x4-pixel tiles instead of 16-pixel row segments (better cache reuse between neighbouring rays).
     * Results are identical; 0 = no assumption. */
    uint32_t tile_w;
    /* optional 8-bit image (trainer.py:1140-1143 `(pred * 255).astype(np.uint8)`): [N,3] uint8 */
    uint8_t *image_u8;
    /* optional multi-GPU fan-out (Part 3): image / depth / weights_sum of every ray are ALSO stored into n_peer_out further
     * buffers -- the other ranks' frame buffers in NVLink peer memory, pointers to the row of ray 0 of this call -- so the
     * kernel's final stores are the all-gather of the narrow outputs.  0 = off. */
    uint32_t n_peer_out;
    float *peer_image[SANERF_MAX_PEERS];
    float *peer_depth[SANERF_MAX_PEERS];
    float *peer_weights_sum[SANERF_MAX_PEERS];
    /* optional cap on the number of persistent CTAs (0 = one per SM): leaves SMs free for a concurrent kernel */
    uint32_t max_ctas;
    /* optional perturbed sampling (perturb=True, renderer.py:267-270 and :99-100): uniform random numbers in [0,1) drawn by the
     * caller, noise0 [N,129] (stage-0 bin jitter), noise1 [N,65], noise2 [N,33] (sample_pdf jitter of the two resamplings) --
     * with torch.rand in this order the stream equals the reference's rand_like calls.  All three or none.  Not available
     * together with sam_in (SANERF_E_CONFIG). */
    const float *noise0, *noise1, *noise2;
    /* REQUIRED device scratch of sanerf_render_workspace_bytes() bytes, 16-byte aligned: the MLP weights re-laid-out as
     * tensor-core operand images, written by a prepare kernel at every call and staged into every CTA's shared memory by TMA */
    void *workspace;
} sanerf_render_args_t;

#define SANERF_RENDER_WORKSPACE_BYTES 65536
size_t sanerf_render_workspace_bytes(void);

/* `model` and `args` are HOST structs (copied at launch); the pointers inside are device pointers.
 * One tiny prepare launch (weights -> operand images in args->workspace) + ONE persistent launch that renders all N rays.
 * Returns launch status. */
int sanerf_render(const sanerf_model_t *model, const sanerf_render_args_t *args, sanerf_stream_t stream);

/* Standalone sample_pdf (renderer.py:84-119), perturb=False: bins [N,T0+1], weights [N,T0] ->
 * new_bins [N,T], inds [N,T] (int16) ; T0+1 <= 129, T in {65,33} with u = the linspace table. */
int sanerf_sample_pdf(const float *bins, const float *weights, const float *u, uint32_t N, uint32_t T0,
                      uint32_t T, float *new_bins, int16_t *inds, sanerf_stream_t stream);

/* Object (instance-mask) head on the tensor cores: replaces `m_grid(xyzs)`, `mask_mlp(cat[masks, geo_feat])` and the weighted
 * sum over samples (nerf/renderer.py:304-305, 376-385; SkipConnMLP 143 -> 256 -> 256 -> n_inst, no bias, leaky_relu 0.01,
 * network.py:119-123).  records: the per-sample (point, geo_feat) records written by sanerf_render with mask_in_tiled = 2,
 * [ceil(n_rays*32/128)][18][128], 16-byte aligned (SANERF_E_CONFIG otherwise: rows travel by TMA bulk copy);
 * weights [n_rays,32] = the final-stage compositing weights; m_grid: the object feature grid (HOST struct, 16 levels x 8
 * channels), gathered inside the kernel by producer warps while the MMAs run; w0 [256,143], w1 [256,256], w2 [n_inst,256] in
 * nn.Linear layout; workspace: device scratch of sanerf_mask_head_workspace_bytes() for the split-precision operand images;
 * logits [n_rays,n_inst].  bf16 hi/lo split operands, fp32 accumulation.  n_inst <= 16. */
size_t sanerf_mask_head_workspace_bytes(void);
int sanerf_mask_head(const float *records, const float *weights, const sanerf_grid_t *m_grid, const float *w0, const float *w1,
                     const float *w2, uint32_t n_inst, uint32_t n_rays, void *workspace, float *logits, sanerf_stream_t stream);

/* SAM feature head on the tensor cores: samvit = LayerNorm(SkipConnMLP(f)) per ray (nerf/renderer.py:361-369,
 * nerf/network.py:113-116): sam_in [n_rays (buffer padded to a multiple of 128 rows), 163] row-major as written by
 * sanerf_render; w[5] / b[5]: samvit_mlp.0.net.{0..4}.{weight,bias} ([256,163] [256,256] [256,419] [256,256] [256,256]);
 * ln_w / ln_b: samvit_mlp.1 (LayerNorm 256, eps 1e-5); out [n_rays,256].  w and b are HOST arrays of device pointers. */
size_t sanerf_samvit_mlp_workspace_bytes(void);
int sanerf_samvit_mlp(const float *sam_in, const float *const *w, const float *const *b, const float *ln_w, const float *ln_b,
                      uint32_t n_rays, void *workspace, float *out, sanerf_stream_t stream);
/* Same with the output layout chosen by the caller: out_nchw = 0 -> [n_rays,256] (as above); 1 -> channel-major [256,n_rays], i.e.
 * the [1,256,H,W] tensor the reference's consumer builds with reshape + permute + contiguous (nerf/trainer.py:540-541), written
 * directly by the head's epilogue. */
int sanerf_samvit_mlp_layout(const float *sam_in, const float *const *w, const float *const *b, const float *ln_w, const float *ln_b,
                             uint32_t n_rays, void *workspace, float *out, uint32_t out_nchw, sanerf_stream_t stream);
/* Fused permute + bilinear resize of a feature frame (nerf/trainer.py:540-546: NHWC -> NCHW, then F.interpolate(mode='bilinear',
 * align_corners=False) to (Ho,Wo)): in_hwc [h,w,C] -> out_chw [C,Ho,Wo], reading only the taps needed. */
int sanerf_feature_resize_nchw(const float *in_hwc, uint32_t h, uint32_t w, uint32_t C, uint32_t Ho, uint32_t Wo, float *out_chw,
                               sanerf_stream_t stream);

/* Validation entry point for the tcgen05 (5th-gen tensor core) MLP path that the fused render uses for grid_mlp
 * (nerf/network.py:9-29 `MLP`, bias-free, ReLU between layers): out[M,16] = relu(relu(x W0^T) W1^T) W2^T with
 * x [M,K], W0 [H,K], W1 [H,H], W2 [16,H] in nn.Linear layout, fp32 in/out, split-precision (3xTF32) tensor-core math.
 * Supported (K,H): (32,64) default network, (8,16) config #1, (16,32). */
int sanerf_mlp3_tc(const float *x, const float *w0, const float *w1, const float *w2, float *out, uint32_t M, uint32_t K,
                   uint32_t H, sanerf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Part 3 -- frame buffers in NVLink peer memory (multi-GPU: row-sharded rendering)
 *
 * Replaces the reference's (dead) eval-time `dist.all_gather(preds)` (nerf/trainer.py:1582-1601): one process per GPU, every
 * rank owns a full-frame buffer that the other ranks map over NVLink (CUDA IPC) and write their row block into -- from the
 * render kernel itself (sanerf_render_args_t::peer_*) or by copy-engine pushes -- followed by one flag barrier.  These are the
 * only entry points that allocate (the buffers must be IPC-exportable cudaMalloc allocations) and the host side of
 * sanerf_peer_alloc / free / open / close synchronises like cudaMalloc does; push and barrier are asynchronous.
 * ---------------------------------------------------------------------------------------- */
#define SANERF_PEER_HANDLE_BYTES 64

int sanerf_peer_alloc(size_t bytes, void **ptr);                                   /* zero-filled device buffer */
int sanerf_peer_free(void *ptr);
int sanerf_peer_export(const void *ptr, uint8_t handle[SANERF_PEER_HANDLE_BYTES]); /* handle to send to the other processes */
int sanerf_peer_open(const uint8_t handle[SANERF_PEER_HANDLE_BYTES], void **ptr);  /* map a peer's buffer (lazy peer access) */
int sanerf_peer_close(void *ptr);
/* n_dst asynchronous device-to-device copies of `bytes` from src to dst[i] on streams[i] (HOST arrays): copy engines, no SMs */
int sanerf_peer_push(void *const *dst, const void *src, size_t bytes, uint32_t n_dst, const sanerf_stream_t *streams);
/* flags: HOST array of `world` device pointers, flags[r] = rank r's flag array (SANERF_MAX_PEERS uint32, zero-initialised) as
 * mapped in this process.  Enqueues on `stream`: signal every peer with `epoch` (strictly increasing from 1), wait until every
 * peer has signalled >= epoch.  After it, everything the peers enqueued before THEIR barrier of the same epoch on the stream
 * they passed is visible here.  status: optional device word set to 1 if a peer did not arrive within timeout_s seconds. */
int sanerf_peer_barrier(uint32_t *const *flags, uint32_t rank, uint32_t world, uint32_t epoch, float timeout_s, uint32_t *status,
                        sanerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SANERF_B200_H */
