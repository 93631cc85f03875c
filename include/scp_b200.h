/*
 * scp_b200.h — C ABI of the B200-native hot path of kywind/self-corr-pose.
 *
 * Every entry point takes raw DEVICE pointers (fp32 unless stated), explicit sizes and an
 * explicit CUDA stream (passed as void* = cudaStream_t).  Nothing allocates; callers own all
 * buffers.  Return value: 0 on success, otherwise a cudaError_t value (or -1 for an argument
 * this build does not support) — never a silent fallback.  All functions are re-entrant per
 * stream.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   scp_softras_forward / scp_softras_backward
 *       third-party/softras/soft_renderer/cuda/soft_rasterize_cuda.cpp:59-91, :94-132
 *       (pybind names forward_soft_rasterize / backward_soft_rasterize, :135-138), kernels
 *       soft_rasterize_cuda_kernel.cu:245-305, :308-483, :486-668.
 *   scp_corr_match_forward / scp_corr_match_backward
 *       model/module/correspondence.py:36-73 (Correspondence.match: bmm + mask + two softmaxes
 *       + two weighted sums), reached through torch ops in the reference.
 *       The rotation-cycle similarity of :105-110 (column softmax, grid.bmm) runs on the same two entry points with
 *       the target pixels in the role of the vertices.
 *   scp_cycle_rows_forward / scp_cycle_rows_backward
 *       model/module/pretrained_corr.py:120-139 (PretrainedCorrespondence.compute_cycle_loss after the DINO matching:
 *       gated softmaxes, corr = Pm . Pi^T column-normalised, match = grid . corr gathered at the top-k pixels, loss).
 *   scp_project_faces_forward / scp_project_faces_backward
 *       model/util/loss_utils.py:38-61 (pinhole_cam, render: camera transform, fp64-promoted projection, y flip,
 *       depth texture), third-party/softras/soft_renderer/transform.py:29-49 + functional/look_at.py:6-62 +
 *       orthogonal.py:4-16 (fixed look_at camera of the model) and functional/face_vertices.py:4-22.
 *   scp_image_losses_forward / scp_image_losses_backward
 *       model/util/loss_utils.py:236-244 (compute_mask_loss), :246-252 (compute_texture_loss), :273-284
 *       (compute_depth_loss), :317-320 (compute_match_loss) with the nearest upsampling of `match`
 *       (model/module/correspondence.py:71), reached through torch ops in the reference.
 *   scp_data_bbox_crop / scp_data_resized_crop
 *       data/dataset_wild6d.py:131-163 (Wild6DDataset.__getitem__ after the file decode: numpy bounding box, crop
 *       intrinsics, three torchvision resized_crop calls on CPU workers in the reference).
 *   scp_vit_*  — see the ViT section below
 *       third-party/zsp/zsp/method/vision_transformer_flexible.py:85-101,116-132,214-262 and
 *       model/module/network/dino.py:102-109.
 */
#ifndef SCP_B200_H
#define SCP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------- */

/* ABI version of this header; bumped on any signature change. */
int scp_abi_version(void);
/* Last error text recorded by a failing call on this host thread (never NULL). */
const char *scp_last_error(void);

/* ---- SoftRas soft rasteriser ----------------------------------------------------------- */
/* enum values are the reference's (functional/soft_rasterize.py:22-25) */
#define SCP_DIST_HARD 0
#define SCP_DIST_BARYCENTRIC 1
#define SCP_DIST_EUCLIDEAN 2
#define SCP_RGB_HARD 0
#define SCP_RGB_SOFTMAX 1
#define SCP_ALPHA_HARD 0
#define SCP_ALPHA_SUM 1
#define SCP_ALPHA_PROD 2
#define SCP_TEX_SURFACE 0
#define SCP_TEX_VERTEX 1

/* Scratch bytes the two SoftRas calls need for (B, nf) (per-face packed records + bboxes). */
size_t scp_softras_workspace_bytes(int B, int nf);

/*
 * forward_soft_rasterize.  faces[B,nf,3,3] screen-space face vertices (x,y in NDC, z depth),
 * textures[B,nf,T,3]; caller pre-fills soft_colors[B,4,is,is] with the background colour and
 * zeroes faces_info[B,nf,27] and aggrs_info[B,2,is,is] (functional/soft_rasterize.py:47-53).
 * dist_eps is the already transformed ln(1/dist_eps-1) (soft_rasterize.py:35).
 * Writes faces_info, aggrs_info, soft_colors in place.
 */
int scp_softras_forward(const float *faces, const float *textures, float *faces_info, float *aggrs_info,
                        float *soft_colors, int B, int nf, int T, int image_size, float near_, float far_,
                        float eps, float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                        int func_id_rgb, int func_id_alpha, int texture_sample_type, int double_side,
                        void *workspace, size_t workspace_bytes, void *stream);

/*
 * backward_soft_rasterize.  grad_faces[B,nf,9] and grad_textures[B,nf,T,3] must be zero-filled
 * by the caller (soft_rasterize.py:88-89) and are accumulated into.
 */
int scp_softras_backward(const float *faces, const float *textures, const float *soft_colors,
                         const float *faces_info, const float *aggrs_info, float *grad_faces,
                         float *grad_textures, const float *grad_soft_colors, int B, int nf, int T,
                         int image_size, float near_, float far_, float eps, float sigma_val, int func_id_dist,
                         float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                         int texture_sample_type, int double_side, void *workspace, size_t workspace_bytes,
                         void *stream);

/*
 * Two renders in ONE traversal (model/module/renderer.py:49-61 of the reference launches them separately): the
 * softmax-RGB render of textures_soft (depth render) and the hard z-buffer render of textures_hard (NOCS map) over the
 * same faces with the same sigma / euclidean distance / 'prod' alpha -- their fragments and alpha channels are
 * identical (soft_rasterize_cuda_kernel.cu:408-417).  Vertex textures [B,nf,3,3]; each output pair is prepared by the
 * caller as for scp_softras_forward.  The *_soft outputs feed scp_softras_backward unchanged.
 */
int scp_softras_forward_dual(const float *faces, const float *textures_soft, const float *textures_hard,
                             float *faces_info, float *aggrs_info_soft, float *soft_colors_soft,
                             float *aggrs_info_hard, float *soft_colors_hard, int B, int nf, int image_size,
                             float near_, float far_, float eps, float sigma_val, float dist_eps, float gamma_val,
                             int double_side, void *workspace, size_t workspace_bytes, void *stream);

/* ---- dense 2D<->3D correspondence (Correspondence.match) --------------------------------- */

/* Scratch bytes of scp_corr_match_forward (foreground block lists, per-row-block column partials; for the shapes the
 * tensor-core forward covers also its split operands and per-tile partials). */
size_t scp_corr_workspace_bytes(int B, int hf, int wf, int N);
/* Scratch bytes of scp_corr_match_backward (foreground block lists only; scp_corr_workspace_bytes is always enough). */
size_t scp_corr_backward_workspace_bytes(int B, int hf, int wf, int N);

/*
 * Correspondence.match (model/module/correspondence.py:36-73), training path.
 *   img_feat[B,C,hf*wf] (unit norm over C), mesh_feat[B,N,C] (unit norm over C), mask_down[B,hf*wf]
 *   (nearest down-sampled mask, 0 = background), pred_v[B,N,3], meshgrid[2,hf*wf], tau = tau_img = tau_mesh.
 * Outputs (any of the two pointcorr pointers may be NULL):
 *   pointcorr_full[B,hf*wf,N]            masked similarity (background rows = -1e5)
 *   pointcorr_pool[B,(hf/2)*(wf/2),N]    its 2x2 mean = F.interpolate(bilinear, 1/2) used by
 *                                        pretrained_corr.py:120-123
 *   match[B,hf*wf,3], imatch[B,2,N]      soft 3D point per pixel / soft 2D location per vertex
 *   rsum[B,hf*wf], csum[B,N]             softmax denominators w.r.t. the reference point tau*1 (saved
 *                                        for the backward; features must be L2-normalised so S <= 1)
 *   A_pool[B,2,N], csum_pool[B,N]        (optional, may be NULL; need pointcorr_pool) column softmax over the
 *                                        POOLED similarity times the pooled meshgrid = "grid.bmm(softmax(tau*
 *                                        pointcorr_src, dim=1))" of pretrained_corr.py:125,131-136, per image
 * Supported shapes: C == 64, wf in {8,16,32,64}, hf*wf a multiple of 128 with 128/wf even.
 * Two machine mappings, same results: with pointcorr_full == NULL and hf*wf a multiple of 256 (every training call of the
 * reference's configs) the similarity is formed on the tcgen05 tensor cores with the accumulator in tensor memory and both
 * soft-maxes in the GEMM epilogue (csrc/scp_corr_tc.cu); otherwise, or with SCP_CORR_FWD=legacy in the environment, by the
 * mma.sync kernel of csrc/scp_corr.cu.
 * Background pixels (mask_down == 0) are not traversed: the kernels walk a per-image list of the 2x2 pixel blocks that
 * contain foreground; the constant outputs of the other blocks (uniform row softmax, -1e5 rows) are filled directly.
 */
int scp_corr_match_forward(const float *img_feat, const float *mesh_feat, const float *mask_down,
                           const float *pred_v, const float *meshgrid, float tau, int B, int hf, int wf, int N,
                           int C, float *pointcorr_full, float *pointcorr_pool, float *match, float *imatch,
                           float *rsum, float *csum, float *A_pool, float *csum_pool, void *workspace,
                           size_t workspace_bytes, void *stream);

/*
 * Backward of the above (autograd of correspondence.py:42-53 in the reference): gradients w.r.t.
 * img_feat and mesh_feat from g_match[B,hf*wf,3], g_imatch[B,2,N] and (optional, may be NULL)
 * g_pointcorr_pool / g_pointcorr_full / g_A_pool (with the saved A_pool, csum_pool).  The similarity tile is
 * recomputed, nothing P x N is read back.  workspace: scp_corr_workspace_bytes (the foreground block list is rebuilt).
 */
int scp_corr_match_backward(const float *img_feat, const float *mesh_feat, const float *mask_down,
                            const float *pred_v, const float *meshgrid, float tau, int B, int hf, int wf, int N,
                            int C, const float *match, const float *imatch, const float *rsum, const float *csum,
                            const float *g_match, const float *g_imatch, const float *g_pointcorr_pool,
                            const float *g_pointcorr_full, const float *A_pool, const float *csum_pool,
                            const float *g_A_pool, float *g_img_feat, float *g_mesh_feat, void *workspace,
                            size_t workspace_bytes, void *stream);

/* ---- frozen DINO ViT-S/8: layer-k key features ----------------------------------------------- */
/* Replaces DINO.forward (model/module/network/dino.py:102-109) / VisionTransformer.get_specific_tokens
 * (third-party/zsp/zsp/method/vision_transformer_flexible.py:249-262): embed 384, 6 heads x 64, MLP 1536,
 * patch 8, LayerNorm eps 1e-6, exact GELU.  All pointers are device pointers.
 *
 * Two precisions of the tensor-core products (`precision` argument):
 *   SCP_VIT_X3   (default of the Python layer, the parity mode): fp32-class.  Every operand is carried as a pair of bf16
 *                numbers hi = bf16(x), lo = bf16(x - hi) and every product as hi*hi + hi*lo + lo*hi with fp32
 *                accumulation (~2^-16 relative): the reference's ViT is fp32 and its consumer is arg-max / top-k.
 *                Linear weights are SPLIT bf16 [out][2*in] in the "i32" layout: groups of 32 input columns stored as
 *                [32 hi | 32 lo].
 *   SCP_VIT_BF16 (labelled fast mode): plain bf16 operands [out][in]; features 5e-3 off the fp32 reference.
 * Everything else (biases, LayerNorm vectors, position embedding, residual stream, softmax) is fp32 in both modes. */
#define SCP_VIT_MAX_BLOCKS 12
#define SCP_VIT_BF16 0
#define SCP_VIT_X3 1

typedef struct scp_vit_block {
    const float *ln1_w, *ln1_b;     /* [384] */
    const void *qkv_w;              /* bf16 [1152][384]  (X3: split [1152][768]) */
    const float *qkv_b;             /* [1152] */
    const void *proj_w;             /* bf16 [384][384]   (X3: split [384][768]) */
    const float *proj_b;            /* [384] */
    const float *ln2_w, *ln2_b;     /* [384] */
    const void *fc1_w;              /* bf16 [1536][384]  (X3: split [1536][768]) */
    const float *fc1_b;             /* [1536] */
    const void *fc2_w;              /* bf16 [384][1536]  (X3: split [384][3072]) */
    const float *fc2_b;             /* [384] */
} scp_vit_block;

typedef struct scp_vit_weights {
    const void *patch_w;            /* bf16 [384][192]: conv weight (384,3,8,8) flattened  (X3: split [384][384]) */
    const float *patch_b;           /* [384] */
    const float *cls_pos0;          /* [384]  = cls_token + pos_embed[0] */
    const float *pos;               /* [np][384] patch position embeddings ALREADY resampled to the np = (H/8)*(W/8)
                                       grid (vision_transformer_flexible.py:192-212) */
    scp_vit_block blocks[SCP_VIT_MAX_BLOCKS];
} scp_vit_weights;

size_t scp_vit_workspace_bytes(int B, int H, int W, int precision);

/*
 * feat[B][384][H/8][W/8] (fp32) = keys of block `n_blocks` (0-based; the reference uses 9) without the CLS
 * token, channel = head*64 + d, after running blocks 0..n_blocks-1 on img[B][3][H][W] (fp32, raw [0,1] RGB).
 */
int scp_vit_s8_keys(const scp_vit_weights *w, const float *img, float *feat, void *feat_tokens, int B, int H, int W,
                    int n_blocks, int precision, void *workspace, size_t workspace_bytes, void *stream);
/* feat_tokens (may be NULL): the same features token-major, the operand layout of scp_dino_argmatch: bf16
 * [B][(H/8)*(W/8)][384] (SCP_VIT_BF16) or split pairs [B][(H/8)*(W/8)][768] in the i32 layout (SCP_VIT_X3).
 *
 * Arg-max matching of DINO features (model/module/pretrained_corr.py:85-89) without materialising the similarity:
 * for pair p and pixel r of image a_idx[p], best[p][r] encodes max over the pixels c of image w_idx[p] with
 * w_mask[p][c] > 0 of <tokens[a_idx[p]][r], tokens[w_idx[p]][c]> as (order-preserving similarity bits << 32 |
 * 0xffffffff - c); 0 = no unmasked column.  np = pixels per image, a multiple of 256; best is zeroed by the callee. */
int scp_dino_argmatch(const void *tokens, const long long *a_idx, const long long *w_idx, const float *w_mask, int B,
                      int np, int NP, int precision, unsigned long long *best, void *stream);

/* Building blocks of the above, exported for unit tests and reuse:
 *   C[M][N] (fp32) = A[M][K] (bf16) * W[N][K]^T (bf16) + bias[N] (may be NULL); N % 128 == 0, K % 64 == 0.
 *   Persistent tcgen05/TMEM GEMM fed by TMA. */
int scp_gemm_bf16_tn(const void *A, const void *W, const float *bias, float *C, int M, int N, int K, void *stream);
/* The same GEMM at fp32-class precision: A [M][2K] and W [N][2K] are split bf16 pairs in the i32 layout; K % 32 == 0. */
int scp_gemm_bf16x3_tn(const void *A, const void *W, const float *bias, float *C, int M, int N, int K, void *stream);
/* o[B][T][384] = softmax(q k^T / 8) v per head; q,k,v bf16 [B*6][T][64] (vision_transformer_flexible.py:85-101). */
int scp_attention_bf16(const void *q, const void *k, const void *v, void *o, int B, int T, void *stream);
/* Same result on the tcgen05 tensor cores (S and O tiles in TMEM, TMA-staged operands); v is passed TRANSPOSED:
 * vt[B*6][64][Tp] bf16, Tp = T rounded up to a multiple of 8, columns t >= T zero.  Used by scp_vit_s8_keys. */
int scp_attention_tc5(const void *q, const void *k, const void *vt, void *o, int B, int T, void *stream);
/* fp32-class attention (split operands, P split by the softmax warps): qk = split q | k token-major [B*T][1536] (head h of
 * q: physical columns [128h, 128h+128), of k: 768 + the same; i32 layout), vt = two planes [2][B*6*64][Tp] (hi, lo) with
 * zero padding, o = split [B*T][768]. */
int scp_attention_x3(const void *qk, const void *vt, void *o, int B, int T, void *stream);

/* ---- symmetry regulariser: rotated surface samples + 1-nearest-neighbour ------------------ */
/*
 * CanonicalMesh.compute_symmetry_loss (model/module/mesh.py:53-62): mesh m = b*k + i has the vertices pred_v[b][N][3];
 * its S surface samples are pts[m][s] = sum_c w[m][s][c] * pred_v[b][faces[face_idx[m][s]][c]] (the random face indices
 * and barycentric weights are drawn by the caller, pytorch3d.ops.sample_points_from_meshes semantics), rotated
 * y = pts . rots[i] (row vector times the 3x3 matrix, `sample_pts.bmm(symm_rots)`).  Outputs, as pytorch3d's
 * knn_points(K=1) inside chamfer_distance_single_way (model/util/chamfer.py:152-156):
 *   dist[m][n] = min_s |pred_v[b][n] - y[m][s]|^2,  nn_idx[m][n] = arg min (lowest s on ties).
 * faces: int32 [nf][3]; face_idx: int64 [B*k][S]; w: [B*k][S][3]; rots: [k][3][3].
 */
int scp_symmetry_nn_forward(const float *pred_v, const int *faces, const long long *face_idx, const float *w,
                            const float *rots, int B, int k, int N, int S, float *dist, int *nn_idx, void *stream);
/* g_pred_v[B][N][3] (zeroed by the callee) = d(sum_mn g_dist[m][n] * dist[m][n]) / d(pred_v): through the vertex itself and
 * through the three face vertices of its nearest sample (face index and weights are constants). */
int scp_symmetry_nn_backward(const float *pred_v, const int *faces, const long long *face_idx, const float *w,
                             const float *rots, const int *nn_idx, const float *g_dist, int B, int k, int N, int S,
                             float *g_pred_v, void *stream);

/* ---- encoder input: photometric jitter + normalisation ------------------------------------ */
/*
 * out = Normalize(mean, std)(ColorJitter(img)) of Encoder.encode_img (model/module/encoder.py:30-32), torchvision tensor
 * semantics (one parameter set for the whole batch): img [B][3][HW] fp32 planar, values in [0,1]; out planar like img, or
 * channels-last [B][HW][3] when nhwc_out == 1 (the layout cuDNN's tensor-core convolutions want), or channels-last with a
 * zero fourth channel [B][HW][4] when nhwc_out == 2 (the stem convolution with a zero-padded weight then runs cuDNN's
 * vectorised NHWC kernels instead of the 3-channel fallback; same result; nhwc_out == 3: padded to eight channels).
 * order[4]: torchvision step ids in application order (0 brightness, 1 contrast, 2 saturation, 3 hue; -1 skips a step);
 * ratios[6] = (ratio, 1 - ratio) of brightness, contrast, saturation AS ROUNDED BY THE CALLER (torchvision forms 1 - ratio
 * in double precision); hue in [-0.5, 0.5]; mean/std [3] host arrays.  All of order/ratios/mean/std are HOST pointers
 * (read during the call).  workspace: scp_color_jitter_workspace_bytes(B) device bytes (per-image grey sums, fp64).
 */
typedef struct scp_jitter_params {     /* device-resident parameter block of scp_color_jitter_normalize_dparams */
    int order[4];                      /* step ids in application order, -1 = skip */
    float r1[3], r2[3];                /* (ratio, 1 - ratio) of brightness, contrast, saturation */
    float hue;
    float mean[3], std[3];
} scp_jitter_params;
size_t scp_color_jitter_workspace_bytes(int B);
/* As scp_color_jitter_normalize, with the parameters read from DEVICE memory at run time (for CUDA-graph replays). */
int scp_color_jitter_normalize_dparams(const float *img, float *out, int B, int HW, const void *params_dev, int nhwc_out,
                                       void *workspace, size_t workspace_bytes, void *stream);
int scp_color_jitter_normalize(const float *img, float *out, int B, int HW, const int *order, const float *ratios, float hue,
                               const float *mean, const float *std, int nhwc_out, void *workspace, size_t workspace_bytes,
                               void *stream);

/* ---- multi-GPU: small-vector exchange over NVLink peer memory (SyncBatchNorm statistics) ------------------------ */
/*
 * The reference converts every BatchNorm of the encoder to SyncBatchNorm (model/trainer.py:66): per training step 80
 * collectives of at most 2*512+1 floats.  These entry points replace NCCL for them by ONE kernel per collective that stores
 * the rank's vector into every peer's buffer over NVLink, raises a flag and waits for the peers' flags (csrc/scp_peer.cu).
 * Set-up (once per channel, host side): every rank creates a buffer and publishes its 64-byte IPC handle; every rank opens
 * the other ranks' handles; `peers[q]` = address of rank q's buffer in THIS process (own buffer for q == rank).
 * scp_peer_exchange: dst[world][n] = the ranks' vectors in rank order (reduce == 0) or dst[n] = their sum in rank order
 * (reduce != 0).  `counter` is one zero-initialised device word per channel (advanced by the kernel: CUDA-graph replays keep
 * counting).  All ranks must issue the same sequence of exchanges on a channel, in stream order.  A peer that never arrives
 * traps the kernel after a bounded spin instead of hanging the GPU.
 */
#define SCP_PEER_SLOTS 4
#define SCP_PEER_MAX_WORLD 16
#define SCP_PEER_MAX_FLOATS 1088
size_t scp_peer_buffer_bytes(void);
int scp_peer_buffer_create(void **ptr, unsigned char *handle64);
int scp_peer_buffer_open(const unsigned char *handle64, void **ptr);
int scp_peer_buffer_close(void *ptr, int own);
int scp_peer_exchange(void *const *peers, const float *src, int n, int rank, int world, unsigned *counter, float *dst,
                      int reduce, void *stream);

/* ---- pre-training cycle loss: the k gathered target rows of every image pair ------------------------------ */
/*
 * pointcorr_pool[B,P4,N] (2x2-pooled similarity), A_pool[B,2,N] (= pooled grid . softmax over pixels, from
 * scp_corr_match_forward), depth_weight[B,N]; pairs p = 0..NP-1: images src_idx[p], tgt_idx[p] (int64), gathered target
 * rows rows[NP,k] (int64, pooled-pixel ids), pts_src[NP,2,k], mask_k[NP,k].  Outputs: pair_loss[p] = sum_j
 * |match_j - pts_src_j|_2 mask_j and match[NP,2,k] with
 *   match_j = sum_n A[src,:,n] w_n Pi_j[n] / (sum_n w_n Pi_j[n] + 1e-5), Pi_j = softmax_n(tau * pointcorr_pool[tgt, rows_j, :]),
 *   w_n = [depth_weight[src,n] >= 0.5] [depth_weight[tgt,n] >= 0.5].
 * (the reference's loss is sum_p pair_loss[p] / (NP * k), pretrained_corr.py:139)
 */
int scp_cycle_rows_forward(const float *pointcorr_pool, const float *A_pool, const float *depth_weight,
                           const long long *src_idx, const long long *tgt_idx, const long long *rows,
                           const float *pts_src, const float *mask_k, float tau, int B, int P4, int N, int NP, int k,
                           float *pair_loss, float *match, void *stream);
/* g_pointcorr_pool[B,P4,N] and g_A_pool[B,2,N] are zeroed and accumulated by the callee (forward recomputed). */
int scp_cycle_rows_backward(const float *pointcorr_pool, const float *A_pool, const float *depth_weight,
                            const long long *src_idx, const long long *tgt_idx, const long long *rows,
                            const float *pts_src, const float *mask_k, float tau, int B, int P4, int N, int NP, int k,
                            const float *g_pair_loss, float *g_pointcorr_pool, float *g_A_pool, void *stream);

/* ---- screen-space geometry shared by the renders of a step ---------------------------------------------- */
/*
 * screen_v[B,N,3] = (x', y', z): c = pred_v . rotation + translation (row vectors), x' = pp_x + c_x f_x / z,
 * y' = -(pp_y + c_y f_y / z) with the fp64 intrinsics foc[B,2], pp[B,2] (evaluated in fp64, rounded to fp32), z = c_z.
 * Optional (faces[nf,3] int32 given): face_vertices[B,nf,3,3] = screen_v gathered per face corner with z + z_offset
 * (the model's look_at camera at (0,0,-z_offset) with the identity rotation, orthographic scale 1) and
 * face_textures[B,nf,3,3] = screen_v gathered (the depth render's texture = the screen-space vertices themselves).
 */
int scp_project_faces_forward(const float *pred_v, const float *rotation, const float *translation, const double *foc,
                              const double *pp, const int *faces, int B, int N, int nf, float z_offset,
                              float *screen_v, float *face_vertices, float *face_textures, void *stream);
/*
 * Gradients w.r.t. pred_v[B,N,3] (may be NULL), rotation[B,3,3], translation[B,3] given any of g_screen_v,
 * g_face_vertices, g_face_textures (NULL = zero).  csr_offsets[N+1] / csr_corners[3*nf]: for every vertex the
 * face corners (face*3 + corner) that reference it, ascending -- the face gradients are gathered per vertex in a
 * fixed order (no atomics on the vertex gradients).
 */
int scp_project_faces_backward(const float *pred_v, const float *rotation, const float *translation, const double *foc,
                               const double *pp, const int *csr_offsets, const int *csr_corners, int B, int N, int nf,
                               const float *g_screen_v, const float *g_face_vertices, const float *g_face_textures,
                               float *g_pred_v, float *g_rotation, float *g_translation, void *stream);

/* out[B,N,3] = S . in[B,N,3] per batch element, S an N x N sparse matrix in CSR form (row_offsets[N+1], cols, vals):
 * the graph-Laplacian product of the smoothness loss (model/util/loss_utils.py:63-97: torch.matmul with the dense
 * N x N buffer); the backward is the same call with the CSR of the transpose. */
int scp_spmm3(const int *row_offsets, const int *cols, const float *vals, const float *in, float *out, int B, int N,
              void *stream);

/* ---- image-space loss terms (silhouette pyramid, texture, depth, 3D match) ------------------------------- */
/*
 * A "map" is a device pointer to a (B, [3,] H, W) fp32 view whose planes are contiguous (channel stride H*W) and
 * whose batch stride is given separately in floats -- so channels of the (B,4,H,W) SoftRas outputs are passed
 * without a copy.  16-byte aligned, W % 16 == 0.  Input map order (maps[] / bstrides[], 11 entries):
 *   0 img(3ch) 1 mask 2 depth 3 mask_render 4 tex_render(3ch) 5 tex_mask 6 depth_render 7 depth_mask
 *   8 match_gt(3ch) 9 match_mask 10 match at full resolution (3ch; only read when hf == 0)
 * match_lr[B][hf*wf][3]: the correspondence kernel's 3D match; upsampled (nearest, top-left rule) on the fly when hf > 0.
 * losses[B][4] = per-image (mask, texture, depth, match) loss values, un-weighted, as the reference's compute_*
 * functions return them.  workspace: scp_image_losses_workspace_bytes(B), written by forward, read by backward
 * (batch-global depth scale sums and its gradient coupling).
 */
size_t scp_image_losses_workspace_bytes(int B);
int scp_image_losses_forward(const void *const *maps, const long long *bstrides, const float *match_lr, int B, int H,
                             int W, int hf, int wf, int use_depth, float *losses, void *workspace, void *stream);
/*
 * g_losses[B][4]: upstream gradient of every loss value.  Gradient maps (g_maps[] / g_bstrides[], 5 entries, same view
 * convention, every element written): 0 mask_render 1 tex_render(3ch) 2 tex_mask 3 depth_render 4 match at full
 * resolution (3ch; hf == 0 only).  g_match_lr[B][hf*wf][3] (hf > 0) is zeroed and accumulated by the callee.
 */
int scp_image_losses_backward(const void *const *maps, const long long *bstrides, const float *match_lr, int B, int H,
                              int W, int hf, int wf, int use_depth, const float *g_losses, const void *workspace,
                              void *const *g_maps, const long long *g_bstrides, float *g_match_lr, void *stream);

/* ---- data path: crop box + resized crop of decoded frames ------------------------------------------ */
/*
 * Wild6DDataset.__getitem__ after the file decode (data/dataset_wild6d.py:131-163): per frame the foreground bounding box of
 * the mask, center / length as the reference computes them (integer `//`, python int() of the float64 products with
 * rand_scale[b] = the two np.random.uniform(1.2, 1.5) draws), the crop box, and the crop intrinsics in float64.
 * mask[B][H][W] uint8 (non-zero = foreground), rand_scale[B][2], intr[B][4] = (fx, fy, cx, cy) float64.
 * Outputs (device): crop[B][4] = (top, left, height, width) int32; center[B][2], length[B][2] int64 (x, y);
 * foc_crop[B][2], pp_crop[B][2] float64 (pixels of the crop, before Trainer.batch_reshape's NDC scaling);
 * status[B] = 1 where the mask has no foreground pixel (the reference raises there; such a frame's crop is empty).
 */
int scp_data_bbox_crop(const unsigned char *mask, const double *rand_scale, const double *intr, int B, int H, int W,
                       int img_size, int no_stretch, int *crop, long long *center, long long *length, double *foc_crop,
                       double *pp_crop, int *status, void *stream);
/*
 * The three torchvision.transforms.functional.resized_crop calls of data/dataset_wild6d.py:152-160 on a batch:
 * img[B][H][W][3] uint8 (RGB, or BGR as cv2 decodes when bgr != 0) -> img_out[B][3][S][S] fp32 in [0,1]: bilinear resize of
 * the zero-padded crop, evaluated in float64 on u8 / 255.0 like the reference's float64 tensor (antialias = 0: torchvision
 * 0.11, the reference's pinned environment; antialias = 1: the triangle filter current torchvision applies by default);
 * mask[B][H][W] uint8 -> mask_out[B][S][S] fp32 {0,1} and depth[B][H][W] uint16 -> depth_out[B][S][S] fp32: nearest resize
 * with torch's float32 source-index rule.  Any of the three output pointers may be NULL (that map is skipped).
 */
int scp_data_resized_crop(const unsigned char *img, const unsigned char *mask, const unsigned short *depth, const int *crop,
                          int B, int H, int W, int img_size, int bgr, int antialias, float *img_out, float *mask_out,
                          float *depth_out, void *stream);

/* ---- channels-last glue of the convolutional encoder ------------------------------------------------ */
/*
 * NHWC fp32 tensors, C % 4 == 0.  Replaces at::native's NHWC kernels reached from the reference through
 * torchvision resnet18's `maxpool` (model/module/network/image_encoder.py:122,127-130), F.interpolate(mode='bilinear') of the
 * feature decoder (image_encoder.py:170-178) and F.normalize(img_feat, 2, 1) (model/module/encoder.py:36).
 * max-pool: kernel 3, stride 2, padding 1; idx[B][OH][OW][C] = winning window position kh*3+kw (first maximum in scan
 * order, NaN propagates -- at::native's rule); OH = (H - 1) / 2 + 1.
 */
int scp_nhwc_maxpool3x3s2_forward(const float *x, float *y, unsigned char *idx, int B, int H, int W, int C, void *stream);
int scp_nhwc_maxpool3x3s2_backward(const float *gy, const unsigned char *idx, float *gx, int B, int H, int W, int C,
                                   void *stream);
/* bilinear resize, align_corners = False (at::native::upsample_bilinear2d formulas); the backward is the exact 2x case
 * (gy[B][2H][2W][C] -> gx[B][H][W][C], gather form: no atomics) */
int scp_nhwc_upsample_bilinear_forward(const float *x, float *y, int B, int H, int W, int C, int OH, int OW, void *stream);
int scp_nhwc_upsample2x_bilinear_backward(const float *gy, float *gx, int B, int H, int W, int C, void *stream);
/* y[B][C][P] = x[B][P][C] / max(||x[b][p][:]||_2, eps) (NHWC in, channel-major out: the correspondence kernel's operand),
 * inv_norm[B][P] = 1 / max(norm, eps), negated where the clamp was active; backward: gx[B][P][C] from gy[B][C][P].  C <= 128. */
int scp_nhwc_l2norm_forward(const float *x, float *y, float *inv_norm, int B, int P, int C, float eps, void *stream);
int scp_nhwc_l2norm_backward(const float *gy, const float *y, const float *inv_norm, float *gx, int B, int P, int C,
                             void *stream);

/* ---- inference pose fit: the point-cloud passes of RANSAC + Umeyama -------------------------------- */
/*
 * Replaces, inside Tester.pose_fitting -> estimateSimilarityTransform (model/tester.py:324-427, model/util/umeyama.py:95-159),
 * the per-round evaluateModel calls (residual norm of a candidate transform over all correspondences, :143-159) and the
 * inlier selection + means / covariance of the final closed-form fit (:28-41, :165-201), for a batch of L images at once.
 *   src, tgt [L][n_max][3]     model-space / camera-space correspondences, image l's counts[l] real ones first
 *   hyp_A [L][H][9], hyp_t [L][H][3]   candidate transforms x -> A x + t (A = scale * rotation, row-major), H <= 128
 *   partial [L][scp_posefit_chunks(n_max)][H]   sums of squared point residuals per 512-point chunk; the residual norm of
 *                                      candidate h of image l is sqrt(sum over the chunks)
 */
size_t scp_posefit_chunks(int n_max);
int scp_posefit_residual_table(const float *src, const float *tgt, const int *counts, const float *hyp_A,
                               const float *hyp_t, int L, int n_max, int H, float *partial, void *stream);
/*
 * Winning transform best_A [L][9], best_t [L][3], pass threshold pass_t [L], found [L] (1 = a RANSAC round was accepted).
 * out [L][18] = points used, inlier count, mean of src (3), mean of tgt (3), sum of (tgt - mean)(src - mean)^T (9, target
 * rows), sum of |src - mean|^2 -- over the inliers (point residual < pass_t) when the image is accepted (found and at least
 * 10 % inliers, umeyama.py:29-31), over all real points otherwise.  Fixed-order reductions: bit-reproducible.
 */
int scp_posefit_inlier_moments(const float *src, const float *tgt, const int *counts, const float *best_A,
                               const float *best_t, const float *pass_t, const unsigned char *found, int L, int n_max,
                               float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCP_B200_H */
