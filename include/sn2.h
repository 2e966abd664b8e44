/*
 * sn2.h -- C ABI of libsn2_b200.so: the B200 (sm_100a) implementation of the StrataNet2
 * PointNet++ forward/backward + 2D coverage-projection hot path.
 *
 * The reference is pure Python (no FFI of its own).  The interfaces these entry points replace are
 * the third-party operator calls on the reference hot path and the reference's own projection
 * functions; each entry cites the call site (file:line under the reference tree) it stands in for.
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions:
 *   - plain pointers and sizes only; every buffer is owned by the caller (device memory unless the
 *     parameter name ends in _host); the library never allocates device memory and never
 *     synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - dense batches: B plots of exactly N points (reference contract, model/point_net2.py:112-116).
 *   - return value: 0 on success, negative SN2_E* otherwise (sn2_error_string() explains).  Asynchronous
 *     kernel faults surface at the caller's next synchronisation, as with any CUDA library.
 *   - stateless and re-entrant per stream; one process per GPU.
 *   - positions are "pos4": float4 per point (x, y, z, unused) in metres.
 *   - indices are int32 inside the library (the Python operator wrappers widen to int64).
 *   - fp32 everywhere; distances are d2 = ((dx*dx + dy*dy) + dz*dz) with each operation rounded
 *     separately (no FMA contraction) so that index outputs are bit-exact against the CPU oracle.
 */
#ifndef SN2_H_
#define SN2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SN2_ABI_VERSION 1

enum {
    SN2_OK = 0,
    SN2_EINVAL = -1,      /* bad shape / null pointer / unsupported argument */
    SN2_EUNSUPPORTED = -2, /* valid request outside what this build implements (e.g. N too large) */
    SN2_ECUDA = -3,       /* a CUDA runtime call or launch failed (cudaGetLastError) */
};

/* fixed network widths of the reference architecture (model/point_net2.py:78-95) */
#define SN2_F0 8   /* point features after dropping x,y                    (:77, :118) */
#define SN2_C1 16  /* SA1 MLP [11,16,16]                                    (:79)       */
#define SN2_C2 32  /* SA2 MLP [19,32]                                       (:80)       */
#define SN2_C3 64  /* SA3 MLP [35,64], FP3 MLP [96,64]                      (:81, :88)  */
#define SN2_CF 34  /* FP2 MLP [80,34], FP1 MLP [42,34]                      (:89-90)    */
#define SN2_CF_LD 36 /* row stride (floats) of 34-wide feature buffers: 16-byte aligned rows */
#define SN2_GRID_MAX 64                       /* max cells per x / y axis of the per-plot grid */
#define SN2_GRID_CELLS (SN2_GRID_MAX * SN2_GRID_MAX)
#define SN2_GRID_HDR 12                       /* floats per plot in the grid header */

int sn2_abi_version(void);
const char *sn2_error_string(int code);
/* last CUDA error text recorded by this thread's most recent failing call (never NULL) */
const char *sn2_last_cuda_error(void);

/* ---- a0-a2: input layout.  Replaces get_long_form + column drop (model/point_net2.py:107-118,
 * 155-158).  xyz (B,3,N), cloud (B,F,N) fp32 -> pos4 [B*N] float4, feat [B*N, F-2] row-major.
 * F must be 10 (n_input_feats, config.py:101).  Either input (with its output) may be NULL: positions and features
 * can be ingested by separate launches, so that FPS starts as soon as xyz has arrived while cloud is still in flight. */
int sn2_ingest(const float *xyz, const float *cloud, int B, int N, int F, float *pos4, float *feat,
               void *stream);

/* ---- a3/a4: farthest point sampling.  Replaces torch_cluster fps (model/point_net2.py:22).
 * start: [B] local start index per plot, or NULL for 0 (canonical random_start=False).
 * idx_out [B*M] GLOBAL row indices (b*N + local), pos4_out [B*M] float4 (may be NULL).
 * Arg-max ties -> lowest index.  M <= N, N <= sn2_fps_max_points(). */
int sn2_fps_max_points(void);
int sn2_fps(const float *pos4, int B, int N, int M, const int *start, int *idx_out,
            float *pos4_out, void *stream);
/* Same contract with an explicit algorithm: both produce identical indices.
 * BRUTE: every point updated every iteration; BUCKETED: Morton-sorted 64-point buckets with exact
 * bounding-box pruning (csrc/fps.cu); AUTO picks BUCKETED above 4096 points, BRUTE below. */
enum { SN2_FPS_AUTO = 0, SN2_FPS_BRUTE = 1, SN2_FPS_BUCKETED = 2,
       SN2_FPS_BUCKETED_SPEC4 = 3 /* experimental: speculative 4-sample rounds, same indices, currently slower */,
       SN2_FPS_CLUSTER4 = 4 /* the bucketed kernel on a 4-CTA cluster per plot (always used above 16384 points) */,
       SN2_FPS_BUCKETED_ILP2 = 5 /* two active buckets of a warp updated per loop trip (interleaved TMEM / REDUX chains) */,
       SN2_FPS_BUCKETED_NW16 = 6 /* 16 warps x 16 bucket slots instead of 8 x 32 */,
       SN2_FPS_BUCKETED_NW16_ILP2 = 7 };
int sn2_fps_algo(const float *pos4, int B, int N, int M, const int *start, int *idx_out,
                 float *pos4_out, int algo, void *stream);

/* ---- a3/a4: radius ball query.  Replaces torch_cluster radius (model/point_net2.py:23-25).
 * Step 1: bin the N points of every plot into an xyz grid (<= SN2_GRID_CELLS cells) whose cell edge is >= r.
 *   grid_hdr [B*SN2_GRID_HDR] float, cell_start [B*(SN2_GRID_CELLS+1)] int32,
 *   sorted4 [B*N] float4 (x, y, z, local index as int bits). */
int sn2_grid_build(const float *pos4, int B, int N, float r, float *grid_hdr, int *cell_start,
                   float *sorted4, void *stream);
/* r < 0 selects an automatic cell edge for about -r points per cell (used by sn2_knn3_grid). */
/* Step 2: neighbours per query = min(K, #{i in plot : d2(i, q) < r2}); cnt [B*M]. */
int sn2_ball_count(const float *grid_hdr, const int *cell_start, const float *sorted4,
                   const float *qpos4, int B, int N, int M, float r2, int K, int *cnt, void *stream);
/* Step 3: exclusive scan cnt -> rowptr [B*M+1] (int32); scratch [B] int32. */
int sn2_rowptr_scan(const int *cnt, int B, int M, int *rowptr, int *scratch, void *stream);
/* Step 4: col[rowptr[q] .. rowptr[q+1]) = the first K in-radius points in ASCENDING GLOBAL INDEX. */
int sn2_ball_fill(const float *grid_hdr, const int *cell_start, const float *sorted4,
                  const float *qpos4, int B, int N, int M, float r2, int K, const int *rowptr,
                  int *col, void *stream);

/* ---- a3/a4/a9: PointConv (eval-mode BN).  Replaces torch_geometric PointConv + scatter max
 * (model/point_net2.py:19,27).  level 1: feat [*,8] -> MLP[11,16,16] -> out [Q,16];
 * level 2: feat [*,16] -> MLP[19,32] -> out [Q,32].  w_host: packed weights (see sn2/weights.py),
 * nw floats, HOST memory (passed to the kernel by value). */
int sn2_pointconv_fwd(int level, const float *pos4, const float *feat, const float *qpos4,
                      const int *rowptr, const int *col, int Q, const float *w_host, int nw,
                      float *out, void *stream);

/* ---- a3/a4 fused (eval): ball query + PointConv in one kernel, no neighbour list (csrc/sa_fused.cu).
 * (grid_hdr, cell_start, sorted4) = sn2_grid_build of the N points with r; qsorted4 [B*M] = the queries in
 * any spatially coherent order with their local query index in .w (sn2_grid_build of the queries gives it).
 * u_scratch [B*N, 16|32]: per-point half of the first layer (written by this call).
 * ovf_scratch [B*M + 1] int32: overflow list of the centroids whose cap binds (redone exactly).
 * out [B*M, 16|32] indexed by the ORIGINAL query index; cnt_out (optional) = min(K, #neighbours).
 * Same edges as sn2_ball_* (including when the cap K binds); values equal sn2_pointconv_fwd to fp32 rounding. */
int sn2_sa_fused_fwd(int level, const float *grid_hdr, const int *cell_start, const float *sorted4,
                     const float *qsorted4, const float *pos4, const float *feat, float *u_scratch,
                     int *ovf_scratch, int B, int N, int M, float r2, int K, const float *w_host, int nw,
                     float *out, int *cnt_out, int tensor_core, void *stream);
/* tensor_core != 0 (level 1 only): the 16x16 second layer runs as tcgen05.mma kind::tf32 tiles (one private
 * M=128 tile per warp, accumulators in TMEM) with a 3xTF32 operand split, so the product keeps fp32 accuracy. */

/* ---- a5: global set abstraction.  Replaces MLP[35,64] + global_max_pool (model/point_net2.py:37-42).
 * x2 [B*M,32], pos4 [B*M] -> g [B,64]. */
int sn2_global_sa_fwd(const float *x2, const float *pos4, int B, int M, const float *w_host, int nw,
                      float *g, void *stream);

/* ---- a6: FP3 (k=1 interpolation from the plot vector + skip + MLP[96,64]); model/point_net2.py:62-67,91.
 * g [B,64], x2 [B*M,32], pos4 [B*M] -> out [B*M,64]. */
int sn2_fp3_fwd(const float *g, const float *x2, const float *pos4, int B, int M, const float *w_host,
                int nw, float *out, void *stream);

/* ---- a7/a8 search part: 3 nearest sources of the same plot per query (torch_cluster knn inside
 * knn_interpolate, model/point_net2.py:63).  Ties -> lower source index.  Ms >= 3.
 * nbr [B*Nq,3] int32 GLOBAL source indices (ascending distance), w [B*Nq,3] = 1/max(d2,1e-16). */
int sn2_knn3(const float *spos4, const float *qpos4, int B, int Ms, int Nq, int *nbr, float *w,
             void *stream);

/* Same result from the xy grid of the SOURCES (sn2_grid_build(spos4, B, Ms, r = -3.0f, ...)): ring search
 * with an exact termination bound, ~100x fewer distance evaluations than the scan above. */
int sn2_knn3_grid(const float *grid_hdr, const int *cell_start, const float *sorted4, const float *qpos4,
                  int B, int Ms, int Nq, int *nbr, float *w, int q_sorted, void *stream);
/* q_sorted != 0: qpos4 is a cell-ordered copy of the queries with the original local index in .w (the sorted4
 * array of an sn2_grid_build of the queries); rows of nbr / w are still indexed by the original query. */

/* ---- a7: FP2 = interpolate(f3 [B*Ms,64]) ++ x1 [B*Nq,16] -> MLP[80,34] -> out [B*Nq, SN2_CF_LD]. */
int sn2_fp2_fwd(const float *f3, const int *nbr, const float *w, const float *x1, int Q,
                const float *w_host, int nw, float *out, void *stream);

/* ---- a8+a10: FP1 = interpolate(f2 [*,SN2_CF_LD]) ++ feat [Q,8] -> MLP[42,34] -> lin1+ReLU -> lin2 ->
 * softmax(4) / sigmoid(1) -> cov = proba*density (model/point_net2.py:141-151).
 * cov, proba: [Q,4] row-major. */
int sn2_fp1_head_fwd(const float *f2, const int *nbr, const float *w, const float *feat, int Q,
                     const float *w_host, int nw, float *cov, float *proba, void *stream);
/* Same contract on the tensor cores: 128-point tiles, MLP[42,34] as tcgen05.mma kind::tf32 with a 3xTF32 hi/lo
 * operand split (fp32-accurate, same 1e-3 bound), accumulators in TMEM, head in the epilogue. */
int sn2_fp1_head_fwd_tc(const float *f2, const int *nbr, const float *w, const float *feat, int Q,
                        const float *w_host, int nw, float *cov, float *proba, void *stream);

/* ---- a12: plot-wise coverages.  Replaces project_to_plotwise_coverages (model/project_to_2d.py:7-55).
 * cloud (B,F,N) device (rows 0,1 = normalised x,y), pred [B*N,4].
 * out [B,4] = [low, 1-low, med, high] mean over occupied pixels.
 * Optional (NULL to skip): pix [B*N] int32 = px*(D+1) + py (px, py in [0, D]); pmax [B,3,(D+1)*(D+1)] fp32 (0 where empty),
 * parg [B,3,(D+1)*(D+1)] int32 (both indexed like pix) GLOBAL point index of the per-pixel max (first index on ties; -1 empty). */
int sn2_project_plotwise(const float *cloud, const float *pred, int B, int N, int F, int D, float *out,
                         int *pix, float *pmax, int *parg, void *stream);

/* ---- a13: rasters.  Replaces project_to_2d_rasters (model/project_to_2d.py:58-113), batched.
 * cloud (B,F,N); cov element (plot b, point n, channel c) at cov[b*cov_sb + n*cov_sn + c*cov_sc];
 * rasters [B,3,D,D] float64, NaN where empty, rows flipped (np.flip axis 0); scale = 10*D/diam_meters,
 * shift = diam_meters // 2.  Optional pix [B*N] int32 = y*D + x (unflipped). */
int sn2_project_rasters(const float *cloud, const float *cov, long long cov_sb, long long cov_sn,
                        long long cov_sc, int B, int N, int F, int D, float scale, float shift,
                        double *rasters, int *pix, void *stream);

/* ================= training-path operators (forward + backward), csrc/train_ops.cu =================
 * In training mode BatchNorm uses batch statistics over all edge messages of the batch, so
 * the message matrix is materialised as in the reference and Linear/BatchNorm stay in torch; these entry
 * points replace the torch_geometric / torch_scatter operators around them and their autograd. */

/* PointConv.message (model/point_net2.py:27): msg [E, C+3] = [x[col[e]], pos[col[e]] - qpos[row(e)]]. */
int sn2_edge_msg_fwd(const float *x, const float *pos4, const float *qpos4, const int *rowptr,
                     const int *col, int Q, int C, float *msg, void *stream);
/* dx [P, C] += dmsg[e][0:C] at row col[e] (dx zero-initialised by the caller; atomics).
 * rows_dev (nullable, device int32): when given, only the first min(E, *rows_dev) edges are processed -- the edge
 * arrays then have a fixed CAPACITY E and the live count stays on the device (CUDA-graph replay of a training step
 * whose edge count changes from batch to batch).  Same meaning wherever `rows_dev` appears below. */
int sn2_edge_msg_bwd(const float *dmsg, const int *col, long long E, const int *rows_dev, int C, float *dx,
                     void *stream);
/* scatter max over CSR rows (PointConv aggr='max', global_max_pool): out [Q,C], arg [Q,C] = first edge
 * attaining the max (-1 and value 0 for an empty row).  C in {16, 32, 64}.
 * ss (nullable, device [2*C] = per-channel scale | shift): the max is taken over vals * scale + shift, i.e. the
 * BatchNorm of the producing block applied on load instead of in a pass of its own (train-mode blocks below). */
int sn2_segment_max_fwd(const float *vals, const float *ss, const int *rowptr, int Q, int C, float *out, int *arg,
                        void *stream);
/* dvals [E,C] (zero-initialised) receives dout at the arg-max edges. */
int sn2_segment_max_bwd(const float *dout, const int *arg, long long Q, int C, long long E, float *dvals, void *stream);
/* knn_interpolate, k=3 (model/point_net2.py:63): y [Q,C] = ((w0 x0 + w1 x1) + w2 x2) / ((w0 + w1) + w2). */
int sn2_interp3_fwd(const float *x, int ldx, const int *nbr, const float *w, long long Q, int C, float *y,
                    void *stream);
/* dx [S,C] (zero-initialised) += (dy / den) * w_k at rows nbr_k. */
int sn2_interp3_bwd(const float *dy, const int *nbr, const float *w, long long Q, int C, float *dx,
                    void *stream);
/* knn_interpolate, k=1 from the single plot vector at the origin (fp3): y [B*M,C] = (g[b] * w) / w. */
int sn2_interp_plot_fwd(const float *g, const float *pos4, int B, int M, int C, float *y, void *stream);
int sn2_interp_plot_bwd(const float *dy, const float *pos4, int B, int M, int C, float *dg, void *stream);
/* backward of sn2_project_plotwise: dpred [B*N,4] (zero-initialised) from dout [B,4] and parg [B,3,D+1,D+1]. */
int sn2_project_plotwise_bwd(const float *dout, const int *parg, int B, int D, float *dpred, void *stream);

/* Weight / bias gradient of a Linear applied to E rows (E ~ millions, Co, Ci <= 64): dW [Co,Ci] = dy^T x,
 * db [Co] = column sums of dy.  partial: scratch [nblk, Co*(Ci+1)]; deterministic two-stage reduction.
 * Supported Ci: 11, 16, 19, 34, 35, 42 (the edge / point MLP and head input widths), Co <= 64. */
int sn2_linear_wgrad_supported(int Co, int Ci);
int sn2_linear_wgrad(const float *dy, const float *x, long long E, int Co, int Ci, float *partial, int nblk,
                     float *dW, float *db, void *stream);

/* ---- train-mode MLP block Linear -> ReLU -> BatchNorm1d (batch statistics), csrc/train_mlp.cu ----------------
 * Replaces, for one block of the reference's MLP() (model/point_net2.py:45-53) applied to R rows, the torch
 * sequence addmm / relu / batch_norm forward and their three backward nodes.  Supported (Ci, Co): (11,16) (16,16)
 * (19,32) (80,34) (42,34) and, as tiled kernels for the short wide blocks, (35,64) (96,64): every block of the network.  All pointers are device pointers.
 *   sn2_lrb_fwd        y [R,Co] = relu(x W^T + b); stats [2*Co+1] fp64 = {sum y, sum y^2, rows} (zeroed here);
 *                      rows = R, or min(R, *rows_dev) when rows_dev is given.  x must be 16-byte aligned (its row
 *                      tiles are fetched with TMA bulk copies).
 *                      For SyncBatchNorm the caller all-reduces stats across ranks before sn2_bn_finalize.
 *   sn2_bn_finalize    stats -> ss [4*Co] = {scale, shift, mean, invstd} (biased variance, eps); running_mean /
 *                      running_var (nullable) updated with `momentum` and the unbiased variance like torch;
 *                      *num_batches_tracked (nullable, int64) += 1.
 *   sn2_bn_apply       z = y * scale + shift.  A consumer that is itself one of these kernels can skip this pass:
 *                      `in_ss` (nullable, device [2*Ci] = scale | shift of the PRODUCING block, i.e. the first half of
 *                      its ss) makes sn2_lrb_fwd / sn2_lrb_bwd read x * scale + shift instead of x, and
 *                      sn2_segment_max_fwd takes the same through its `ss`; sn2_lrb_block_fwd skips it when z is NULL.
 *                      The gradient wrt that virtual z is what sn2_lrb_bwd returns in dx / expects in dz.
 *   sn2_lrb_bwd_reduce sums [2*Co] fp64 = {sum dz, sum dz*y} (zeroed here; all-reduced by the caller for SyncBN).
 *   sn2_lrb_bwd        dx [R,Ci] (nullable) = dy W with dy = relu'(y) * BN'(dz); dW [Co,Ci] = dy^T x; db [Co];
 *                      partial: scratch [nblk, Co*(Ci+1)], fixed-order two-stage reduction. */
int sn2_lrb_supported(int Co, int Ci);
int sn2_lrb_fwd(const float *x, const float *in_ss, const float *W, const float *b, long long R, const int *rows_dev,
                int Co, int Ci, float *y, double *stats, void *stream);
int sn2_bn_finalize(const double *stats, const float *gamma, const float *beta, float eps, float momentum,
                    float *running_mean, float *running_var, long long *num_batches_tracked, float *ss, int Co,
                    void *stream);
int sn2_bn_apply(const float *y, const float *ss, long long R, const int *rows_dev, int Co, float *z, void *stream);
int sn2_lrb_bwd_reduce(const float *dz, const float *y, long long R, const int *rows_dev, int Co, double *sums,
                       void *stream);
/* BatchNorm affine gradients of THIS rank from its own (not all-reduced) sums: dbeta = sum dz, dgamma = sum dz*yhat. */
int sn2_bn_param_grad(const double *sums, const float *ss, int Co, float *dgamma, float *dbeta, void *stream);
int sn2_lrb_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                const double *sums, const double *stats, long long R, const int *rows_dev, int Co, int Ci, float *dx,
                float *partial, int nblk, float *dW, float *db, void *stream);
/* The same block in one call each way when no all-reduce sits between the kernels (plain BatchNorm1d):
 * fwd = sn2_lrb_fwd + sn2_bn_finalize (num_batches_tracked += 1 when given) + sn2_bn_apply;
 * bwd = sn2_lrb_bwd_reduce + sn2_bn_param_grad + sn2_lrb_bwd. */
int sn2_lrb_block_fwd(const float *x, const float *in_ss, const float *W, const float *b, const float *gamma, const float *beta,
                      float eps, float momentum, float *running_mean, float *running_var,
                      long long *num_batches_tracked, long long R, const int *rows_dev, int Co, int Ci, float *y,
                      double *stats, float *ss, float *z, void *stream);
int sn2_lrb_block_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                      const double *stats, long long R, const int *rows_dev, int Co, int Ci, double *sums,
                      float *dgamma, float *dbeta, float *dx, float *partial, int nblk, float *dW, float *db,
                      void *stream);

/* ---- train-mode head + point-wise losses, csrc/train_head.cu ---------------------------------------------------
 * sn2_head_fwd   replaces model/point_net2.py:141-153 under train(): relu(lin1) -> lin2 -> softmax(4) x sigmoid(1) on
 *                f1 [R,34] (row stride 34; in_ss nullable = scale|shift [2*34] of the producing block, applied on
 *                load) with the LIVE parameter tensors W1 [16,34], b1 [16], W2 [5,16], b2 [5] (device) -> cov, proba [R,4].
 * sn2_head_bwd   recomputes the head from f1; df1 [R,34] (nullable) = gradient of the (BatchNorm-applied) row; dW1, db1,
 *                dW2, db2 via `partial` [nblk, sn2_head_bwd_partials()] and a fixed-order reduction.  dcov / dproba
 *                [R,4] nullable (= zero).  dropout p must be 0 (config.py:76 default); otherwise the caller uses torch.
 * sn2_pointwise_loss_fwd/bwd   learning/loss_functions.py:19-57: out[0] = mean_i -log((p0+p1) pdf0 + p2 pdf1 + p3 pdf2)
 *                (fp64, pdf [R,3] fp64), out[1] = mean binary entropy of columns 2,3 (EPS 1e-4); sums [2] fp64 scratch;
 *                bwd: dproba [R,4] = g[0] d out[0] + g[1] d out[1], g [2] fp64 on the device.
 * sn2_kde_lut    pdf [B*N,3] fp64 = linear interpolation (scipy interp1d / learning/kde_mixture.py:64-75) of Y [3,K]
 *                over sorted knots X [K] (fp64) at z = fp32(cloud[b,2,n] * z_max), clamped to the grid. */
int sn2_head_fwd(const float *f1, const float *in_ss, const float *W1, const float *b1, const float *W2, const float *b2,
                 long long R, float *cov, float *proba, void *stream);
int sn2_head_bwd_partials(void);
int sn2_head_bwd(const float *f1, const float *in_ss, const float *W1, const float *b1, const float *W2, const float *b2,
                 const float *dcov, const float *dproba, long long R, float *df1, float *partial, int nblk, float *dW1,
                 float *db1, float *dW2, float *db2, void *stream);
int sn2_pointwise_loss_fwd(const float *proba, const double *pdf, long long R, double *sums, double *out, void *stream);
int sn2_pointwise_loss_bwd(const float *proba, const double *pdf, const double *g, long long R, float *dproba, void *stream);
int sn2_kde_lut(const float *cloud, int B, int F, int N, float z_max, const double *X, const double *Y, int K, double *pdf,
                void *stream);

/* ---- train-mode SA1 block without materialised messages, csrc/train_sa.cu ------------------------------------------
 * Replaces, under model.train(), model/point_net2.py:21-29 for sa1 (PointConv(local_nn = MLP([11,16,16])), aggr = max):
 * the message MLP is RECOMPUTED from a per-point table in every sweep over the CSR neighbour lists (rowptr [M+1], col
 * [>= E] global point indices, live edge count = rowptr[M]) instead of writing [E,11] / [E,16] / [E,16] arrays.  LIVE
 * parameter tensors (device): W1 [16,11], b1 [16], W2 [16,16], b2 [16], gamma2 [16].  Statistics and backward sums are
 * raw fp64 sums as in the lrb_* entry points; the caller runs sn2_bn_finalize[_sync] / sn2_bn_param_grad /
 * sn2_bn_bwd_sync between the calls (SyncBatchNorm: same collectives as the materialising path).  queue: device int32
 * scratch, the cursor of a sweep's work queue (reset by the call itself).
 *   sn2_sa1t_pre      u [P,16] = W1[:, :8] feat + W1[:, 8:] pos
 *   sn2_sa1t_stats1   stats1 [33] fp64 = {sum a1, sum a1^2, E}, a1 = relu(u[col] + b1 - W1p q)              -> ss1
 *   sn2_sa1t_stats2   stats2 [33] of a2 = relu(W2 BN1(a1) + b2); key / arg [M,16]: arg-max edge of sign(gamma2) * a2
 *                     (first edge on ties, -1 for an empty row)                                             -> ss2
 *   sn2_sa1t_finish   x1 [M,16] = BN2(a2[arg]) (0 for an empty row), amax [M,16] = a2[arg]
 *   sn2_sa1t_bwd_sums sums2 [32] fp64 = {sum dx1, sum dx1 * amax}
 *   sn2_sa1t_bwd_w2   dW2 [16,16], db2 [16] and sums1 [32] fp64 (BatchNorm 1's raw backward sums); partial
 *                     [sn2_sa1t_blocks(), sn2_sa1t_partials()] scratch
 *   sn2_sa1t_bwd_in   du [P,16] (zeroed here, vector atomics) = sum over the edges of a point of dz1, dc [M,16] = row sums
 *   sn2_sa1t_bwd_w1   dW1 [16,11], db1 [16] from du, dc (same scratch); the gradient of the point features, when needed,
 *                     is du W1[:, :8] (a plain GEMM of the caller). */
int sn2_sa1t_partials(void);
int sn2_sa1t_blocks(void);
int sn2_sa1t_pre(const float *feat, const float *pos4, long long P, const float *W1, float *u, void *stream);
int sn2_sa1t_stats1(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                    const float *b1, double *stats1, int *queue, void *stream);
int sn2_sa1t_stats2(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                    const float *b1, const float *W2, const float *b2, const float *ss1, const float *gamma2, double *stats2,
                    float *key, int *arg, int *queue, void *stream);
int sn2_sa1t_finish(const float *key, const int *arg, const float *gamma2, const float *ss2, int M, float *x1, float *amax,
                    void *stream);
int sn2_sa1t_bwd_sums(const float *dout, const float *amax, const int *arg, int M, double *sums2, void *stream);
int sn2_sa1t_bwd_w2(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                    const float *b1, const float *W2, const float *b2, const float *gamma2, const float *ss1, const float *ss2,
                    const double *stats2, const double *sums2, const float *dout, const int *arg, float *partial, float *dW2,
                    float *db2, double *sums1, int *queue, void *stream);
int sn2_sa1t_bwd_in(const float *u, const float *qpos4, const int *rowptr, const int *col, long long P, int M, const float *W1,
                    const float *b1, const float *W2, const float *b2, const float *gamma2, const float *ss1, const float *ss2,
                    const double *stats1, const double *stats2, const double *sums1, const double *sums2, const float *dout,
                    const int *arg, float *du, float *dc, int *queue, void *stream);
int sn2_sa1t_bwd_w1(const float *du, const float *dc, const float *feat, const float *pos4, const float *qpos4, long long P,
                    int M, float *partial, float *dW1, float *db1, void *stream);

/* sa2 (PointConv(local_nn = MLP([19,32])), one block), same scheme with lane = channel: LIVE W [32,19], b [32], gamma [32].
 *   sn2_sa2t_pre      u [P,32] = W[:, :16] x + W[:, 16:] pos
 *   sn2_sa2t_fwd      stats [65] fp64 of a = relu(u[col] + b - Wp q); key / arg [M,32]                       -> ss
 *   sn2_sa2t_finish   x2 [M,32] = BN(a[arg]), amax [M,32]
 *   sn2_sa2t_bwd_sums sums [64] fp64 = {sum dx2, sum dx2 * amax}
 *   sn2_sa2t_bwd      du [P,32] (zeroed here) += relu'(a) BN'(dz) over the edges of a point, dc [M,32] = row sums
 *   sn2_sa2t_bwd_w    dW [32,19], db [32] (partial as above); the gradient of x is du W[:, :16] (caller's GEMM) */
int sn2_sa2t_pre(const float *x, const float *pos4, long long P, const float *W, float *u, void *stream);
int sn2_sa2t_fwd(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W, const float *b,
                 const float *gamma, double *stats, float *key, int *arg, int *queue, void *stream);
int sn2_sa2t_finish(const float *key, const int *arg, const float *gamma, const float *ss, int M, float *x2, float *amax,
                    void *stream);
int sn2_sa2t_bwd_sums(const float *dout, const float *amax, const int *arg, int M, double *sums, void *stream);
int sn2_sa2t_bwd(const float *u, const float *qpos4, const int *rowptr, const int *col, long long P, int M, const float *W,
                 const float *b, const float *gamma, const float *ss, const double *stats, const double *sums, const float *dout,
                 const int *arg, float *du, float *dc, int *queue, void *stream);
int sn2_sa2t_bwd_w(const float *du, const float *dc, const float *x, const float *pos4, const float *qpos4, long long P, int M,
                   float *partial, float *dW, float *db, void *stream);

/* ---- peer-memory collectives for data-parallel training, csrc/comm.cu ----------------------------
 * The reference has no distributed code; these stand where a DistributedDataParallel / SyncBatchNorm wrapper around
 * learning/train.py:52-66 would call NCCL.  One region per rank (the ONLY device memory this library allocates),
 * exported / imported with CUDA IPC; a communicator is the table of all ranks' regions.  Collectives are one-shot
 * all-reduces over NVLink peer stores, executed inside the kernels that need the result; all collectives of a rank
 * must be enqueued on one stream and in the same order on every rank.  comm == NULL means a single rank (no exchange).
 *   sn2_bn_finalize_sync   sn2_bn_finalize with stats [2Co+1] summed over the ranks first (in place).
 *   sn2_bn_bwd_sync        sn2_bn_param_grad on this rank's sums, then sums [2Co] summed over the ranks (in place).
 *   sn2_adam_step          grad <- sum_r gscale_r * grad_r (in place), then torch.optim.Adam's update of the flat
 *                          parameter bucket (L2 weight decay, bias correction) with lr and step read from the device.
 *   sn2_comm_status        synchronous: collectives completed, sticky error word (non-zero: a wait timed out). */
size_t sn2_comm_region_bytes(void);
int sn2_comm_max_world(void);
int sn2_comm_max_bytes(void);
int sn2_comm_region_alloc(void **region);
int sn2_comm_region_free(void *region);
int sn2_comm_ipc_export(void *region, void *handle64_host);
int sn2_comm_ipc_import(const void *handle64_host, void **peer_region);
int sn2_comm_ipc_release(void *peer_region);
int sn2_comm_create(int rank, int world, void *const *regions_host, void **comm_out);
int sn2_comm_destroy(void *comm);
int sn2_comm_status(void *comm, long long *seq_out, long long *err_out);
int sn2_comm_allreduce_f64(void *comm, double *buf, int n, void *stream);
int sn2_comm_allreduce_f32(void *comm, float *buf, int n, float scale, void *stream);
int sn2_bn_finalize_sync(void *comm, double *stats, const float *gamma, const float *beta, float eps, float momentum,
                         float *running_mean, float *running_var, long long *num_batches_tracked, float *ss, int Co,
                         void *stream);
int sn2_bn_bwd_sync(void *comm, double *sums, const float *ss, int Co, float *dgamma, float *dbeta, void *stream);
int sn2_adam_step(void *comm, float *grad, float gscale, float *param, float *m, float *v, long long n, const float *lr_dev,
                  float b1, float b2, float eps, float wd, long long *step_dev, void *stream);

/* ================= local-map fusion (BASELINE config 4), csrc/fusion.cu =============
 * Weighted-average mosaic of per-plot rasters into the parcel grid; replaces add_weights_band_to_rasters +
 * rasterio.merge(method=_weighted_average_of_rasters) (inference/geotiff_raster.py:103-118, 199-235, 294-347).
 * rasters [P,3,D,D] float64 (NaN = no value), offsets [P,2] int32 = (row, col) of each plot's top-left pixel in
 * the parcel grid; num/den [3,H,W] and wsum [H,W] float64 accumulators (zero-initialised by the caller; may be
 * all-reduced across ranks before finalize).  out [4,H,W]: 3 averaged bands + sum of weights, NaN where empty. */
int sn2_fuse_accumulate(const double *rasters, const int *offsets, int P, int D, int H, int W, double *num,
                        double *den, double *wsum, void *stream);
int sn2_fuse_finalize(const double *num, const double *den, const double *wsum, int H, int W, double *out,
                      void *stream);

/* ---- device-side loader transforms, csrc/loader.cu -------------------------------------------
 * augment (data_loader/loader.py:161-214) + rescale_cloud (:135-158) of a batch: in [B,10,N] fp32 = centred raw plots (metres,
 * raw feature units, fake ground points included).  angle [B] float64 radians + flip [B,2] uint8 (nullable together: no
 * rotation / flips), noise [B,6,N] float64 standard-normal draws for x, y, R, G, B, NIR (nullable: no noise).
 * Out: xyz [B,3,N] (rotated / flipped positions, metres) and cloud [B,10,N] (augmented, rescaled): the model input. */
int sn2_augment_rescale(const float *in, int B, int N, const double *angle, const unsigned char *flip, const double *noise,
                        float z_max, float *xyz, float *cloud, void *stream);

/* ================= parcel tiling / plot preparation / band finalisation, csrc/parcel.cu ====
 * sn2_parcel_grid_build  uniform xy grid (cell side `cell`, nx*ny cells from (x0,y0)) over a parcel cloud of P points:
 *                        cell_start [nx*ny+1], sorted4 [P] float4 = (x, y, z, index bits) grouped by cell (16-byte aligned),
 *                        pos_of [P] = position of every point in sorted4; cell_of [P], count and cursor [nx*ny] are scratch.
 *                        Stands where prepare.py:75-76 builds a KDTree over the parcel.  cell >= 1.05 * znorm_radius.
 * sn2_extract_plots      one CTA per plot centre (centers [C,2] float64): points with float64 d2 <= radius^2 in ascending
 *                        index (inference/prepare_utils.py:47-53), plots with <= min_points points skipped (:67-69,
 *                        prepare.py:91-94), z - min z within znorm_radius (utils/load_data.py:237-249), centred, the fake
 *                        ground points of data_loader/loader.py:90-105 appended, xyz kept, rescaled (:135-158), sub- /
 *                        up-sampled to S (:233-255, canonical hash rule, seeds [C] uint32).  xyz = parcel rows 0-2
 *                        [3,P] fp32, feat = rows 3-9 [7,P].  Out: out_xyz [C,3,S], out_cloud [C,10,S] fp32, out_n [C]
 *                        (points in the disk; negative = more than sn2_plot_capacity()), out_src [C,S] nullable (parcel
 *                        index of every output point, -1-k for fake point k).
 * sn2_finalize_mosaic    fused [4,H,W] float64 (Vb, Vm, Vh, weights; NaN = none) -> out [5,H,W] = (Vb, Vm_soft, Vh, Vm_hard,
 *                        weights) per inference/geotiff_raster.py:121-146, 273-291 (without the admissibility band);
 *                        hist [10002] uint32 and scratch [4] float64 are workspace; scratch[2] = threshold, [3] = target. */
int sn2_parcel_grid_build(const float *x, const float *y, const float *z, long long P, float x0, float y0, float cell, int nx, int ny,
                          int *cell_of, int *count, int *cell_start, int *cursor, float *sorted4, int *pos_of, void *stream);
int sn2_plot_capacity(void);
int sn2_extract_plots(const float *xyz, const float *feat, long long P, float x0, float y0, float cell, int nx, int ny,
                      const int *cell_start, const float *sorted4, const int *pos_of, const double *centers, const unsigned *seeds, int C, int S,
                      float radius, float znorm_radius, float z_max, int diam_meters, int min_points, float *out_xyz,
                      float *out_cloud, int *out_n, int *out_src, void *stream);
unsigned sn2_sample_hash(unsigned seed, unsigned j);
int sn2_finalize_mosaic(const double *fused, int H, int W, unsigned *hist, double *scratch, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SN2_H_ */
