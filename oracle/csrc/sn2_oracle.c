/*
 * sn2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the third-party point-cloud searches that the
 * reference hot path calls (they live in un-vendored wheels):
 *
 *   o_fps     <- torch_cluster.fps       called at  /root/reference/model/point_net2.py:22
 *   o_radius  <- torch_cluster.radius    called at  /root/reference/model/point_net2.py:23-25
 *   o_knn     <- torch_cluster.knn       called through torch_geometric.nn.knn_interpolate at
 *                                         /root/reference/model/point_net2.py:63
 *
 * Semantics follow SURVEY.md Appendix A (A1, A2, A5), which is the project's
 * normative statement of torch-cluster==1.5.9 behaviour
 * (/root/reference/setup_environment/torch_extensions.txt:1-4).
 * PARITY WITH THE REAL WHEELS IS UNPINNED: they are not installable here.  The
 * pins we do have are the committed golden vectors under tests/golden/
 * (reference model files run verbatim on these ops) and known-answer cases.
 *
 * All distances are IEEE fp32, d2 = ((dx*dx + dy*dy) + dz*dz), every product
 * and sum individually rounded: compile with -ffp-contract=off (see Makefile).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float d2f(const float *a, const float *b)
{
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float xx = dx * dx;
    float yy = dy * dy;
    float zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

/*
 * A1.  Farthest point sampling per segment.
 *   pos      [ptr[B], 3] fp32 row-major
 *   ptr      [B+1] segment offsets into pos
 *   optr     [B+1] segment offsets into out (optr[b+1]-optr[b] = m_b, computed by the caller
 *            as ceil(float32(n_b) * float32(ratio)))
 *   start    [B] local start index per segment, or NULL for 0 (canonical random_start=False)
 *   out      [optr[B]] int64 GLOBAL row indices
 * dist[i] = min(dist[i], d2(i, last)); next = argmax(dist), ties -> lowest index.
 */
void o_fps(const float *pos, const int64_t *ptr, const int64_t *optr, const int64_t *start,
           int64_t B, int64_t *out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        int64_t n = ptr[b + 1] - ptr[b];
        int64_t m = optr[b + 1] - optr[b];
        if (n <= 0 || m <= 0)
            continue;
        const float *p = pos + 3 * ptr[b];
        int64_t *o = out + optr[b];
        float *dist = (float *)malloc(sizeof(float) * (size_t)n);
        for (int64_t i = 0; i < n; ++i)
            dist[i] = INFINITY;
        int64_t last = start ? start[b] : 0;
        o[0] = ptr[b] + last;
        for (int64_t it = 1; it < m; ++it) {
            const float *q = p + 3 * last;
            float best = -1.0f;
            int64_t besti = 0;
            for (int64_t i = 0; i < n; ++i) {
                float d = d2f(p + 3 * i, q);
                float cur = dist[i];
                cur = d < cur ? d : cur;
                dist[i] = cur;
                if (cur > best) { /* strict: first (lowest) index wins ties */
                    best = cur;
                    besti = i;
                }
            }
            last = besti;
            o[it] = ptr[b] + last;
        }
        free(dist);
    }
}

/*
 * A2.  Radius (ball) query.  For each query q of segment b (ascending), the points of the
 * same segment of x in ascending index with d2 < r2 (strict), at most K of them.
 *   mode 0: write cnt[q] only.          mode 1: write col at rowptr[q] (global x indices).
 */
void o_radius(const float *x, const int64_t *ptr_x, const float *y, const int64_t *ptr_y,
              int64_t B, float r2, int64_t K, int mode, int64_t *cnt, const int64_t *rowptr,
              int64_t *col)
{
    for (int64_t b = 0; b < B; ++b) {
        int64_t x0 = ptr_x[b], x1 = ptr_x[b + 1];
#pragma omp parallel for schedule(static)
        for (int64_t q = ptr_y[b]; q < ptr_y[b + 1]; ++q) {
            const float *yq = y + 3 * q;
            int64_t c = 0;
            int64_t *dst = mode ? col + rowptr[q] : NULL;
            for (int64_t i = x0; i < x1 && c < K; ++i) {
                /* A2 / upstream CUDA kernel: diff = x - y */
                float d = d2f(x + 3 * i, yq);
                if (d < r2) {
                    if (dst)
                        dst[c] = i;
                    ++c;
                }
            }
            if (!mode)
                cnt[q] = c;
        }
    }
}

/*
 * A5 (search part).  k nearest sources of the same segment for each query; ascending
 * distance, ties -> lower source index first (strict '<' insertion in ascending index scan).
 *   idx [Ny, k] global source indices (-1 where the segment has fewer than k sources)
 *   d2  [Ny, k] fp32 squared distances (diff = source - query, as PyG's knn_interpolate forms it)
 */
void o_knn(const float *x, const int64_t *ptr_x, const float *y, const int64_t *ptr_y,
           int64_t B, int64_t k, int64_t *idx, float *d2)
{
    for (int64_t b = 0; b < B; ++b) {
        int64_t x0 = ptr_x[b], x1 = ptr_x[b + 1];
#pragma omp parallel for schedule(static)
        for (int64_t q = ptr_y[b]; q < ptr_y[b + 1]; ++q) {
            const float *yq = y + 3 * q;
            int64_t *bi = idx + q * k;
            float *bd = d2 + q * k;
            for (int64_t j = 0; j < k; ++j) {
                bi[j] = -1;
                bd[j] = INFINITY;
            }
            for (int64_t i = x0; i < x1; ++i) {
                float d = d2f(x + 3 * i, yq);
                if (d < bd[k - 1] || bi[k - 1] < 0) {
                    int64_t j = k - 1;
                    while (j > 0 && (bi[j - 1] < 0 || d < bd[j - 1])) {
                        bd[j] = bd[j - 1];
                        bi[j] = bi[j - 1];
                        --j;
                    }
                    bd[j] = d;
                    bi[j] = i;
                }
            }
        }
    }
}

int o_abi_version(void) { return 1; }
