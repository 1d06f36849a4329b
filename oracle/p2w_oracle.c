/*
 * p2w_oracle.c -- CPU restatement of the neighbourhood primitives on PointsToWood's
 * inference hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
 * checker or the timed CPU arm; never by the product path (pointstowood_b200/).
 *
 * PARITY UNPINNED for the primitives in this file: the reference tree holds no tests,
 * golden vectors or fixtures, and the arithmetic lives in
 * un-vendored third-party wheels that are absent from /root/reference and cannot be
 * installed here: torch-cluster ~1.6.3 (knn, radius, grid, fps), torch-scatter ~2.1.2,
 * torch-geometric ~2.6 (README.md:42-48 of the reference pins only the wheel index).
 * What follows restates their published CUDA-kernel semantics (SURVEY.md Appendix A.2,
 * A.3, A.4, A.11) and is anchored on the reference's call sites:
 *   knn     pointstowood/src/model.py:120 (SA2/SA3, k=32), :149 (knn_interpolate, k=2)
 *   radius  pointstowood/src/model.py:118 (SA1, r=0.08, max 32)
 *   grid    pointstowood/src/model.py:104, pointstowood/src/preprocessing.py:33,58
 *   fps     not called by the reference; torch_cluster.fps semantics (A.11)
 *
 * Distances are accumulated exactly as nvcc compiles upstream's
 *   tmp += (x[d]-y[d])*(x[d]-y[d])
 * i.e. one FP32 subtract and one fused multiply-add per dimension; fmaf() is the exact
 * single-rounding FMA, and -ffp-contract=off keeps gcc from inventing others.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* gcc in this image has no libgomp spec, so the query loops are spread over plain
 * pthreads: contiguous blocks of `grain` iterations handed out from a shared counter. */
typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t lo, hi, grain; int64_t next; pthread_mutex_t mu; } pf_job;

static void *pf_worker(void *arg) {
    pf_job *j = (pf_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t a = j->next;
        j->next += j->grain;
        pthread_mutex_unlock(&j->mu);
        if (a >= j->hi) break;
        int64_t b = a + j->grain < j->hi ? a + j->grain : j->hi;
        j->fn(a, b, j->ctx);
    }
    return NULL;
}

static int orc_threads = 0;
void orc_set_threads(int n) { orc_threads = n; }
int orc_get_threads(void) {
    if (orc_threads > 0) return orc_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(int64_t lo, int64_t hi, int64_t grain, range_fn fn, void *ctx) {
    int nt = orc_get_threads();
    if (nt > 256) nt = 256;
    if (hi - lo <= grain || nt <= 1) { if (hi > lo) fn(lo, hi, ctx); return; }
    pf_job j = { fn, ctx, lo, hi, grain, lo, PTHREAD_MUTEX_INITIALIZER };
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nt - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &j) == 0) started++;
    pf_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

static inline float sqdist(const float *a, const float *b, int dim) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) {
        float t = a[i] - b[i];
        d = fmaf(t, t, d);
    }
    return d;
}

/* A.2: k nearest sources of the same example, ascending by (d, index); list seeded with
 * (1e10, -1); a candidate is inserted at the first slot whose distance is strictly
 * greater.  nbr is [Ny, k] int64, -1 padded; d2 (optional) the matching distances. */
typedef struct { const float *x, *y; int64_t x0, x1; int dim, k; int64_t *nbr; float *d2; } knn_ctx;

static void knn_range(int64_t lo, int64_t hi, void *vc) {
    const knn_ctx *c = (const knn_ctx *)vc;
    const int k = c->k, dim = c->dim;
    for (int64_t q = lo; q < hi; q++) {
        float bd[100];
        int64_t bi[100];
        for (int e = 0; e < k; e++) { bd[e] = 1e10f; bi[e] = -1; }
        const float *yq = c->y + q * dim;
        for (int64_t n = c->x0; n < c->x1; n++) {
            float d = sqdist(c->x + n * dim, yq, dim);
            if (!(bd[k - 1] > d)) continue;               /* cannot enter the list */
            int e1 = 0;
            while (!(bd[e1] > d)) e1++;
            for (int e2 = k - 1; e2 > e1; e2--) { bd[e2] = bd[e2 - 1]; bi[e2] = bi[e2 - 1]; }
            bd[e1] = d; bi[e1] = n;
        }
        for (int e = 0; e < k; e++) {
            c->nbr[q * k + e] = bi[e];
            if (c->d2) c->d2[q * k + e] = bd[e];
        }
    }
}

int orc_knn(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
            int B, int dim, int k, int64_t *nbr, float *d2) {
    if (k < 1 || k > 100) return -1;      /* upstream: AT_ASSERTM(k <= 100) */
    for (int b = 0; b < B; b++) {
        knn_ctx c = { x, y, ptr_x[b], ptr_x[b + 1], dim, k, nbr, d2 };
        parallel_for(ptr_y[b], ptr_y[b + 1], 64, knn_range, &c);
    }
    return 0;
}

/* A.3 (CUDA semantics): ascending index scan, emit while d < (float)(r*r), stop after
 * max_nbr hits.  nbr is [Ny, max_nbr] int64, -1 padded; cnt [Ny]. */
typedef struct { const float *x, *y; int64_t x0, x1; int dim, max_nbr; float r2; int64_t *nbr; int32_t *cnt; } rad_ctx;

static void radius_range(int64_t lo, int64_t hi, void *vc) {
    const rad_ctx *c = (const rad_ctx *)vc;
    const int dim = c->dim, mx = c->max_nbr;
    for (int64_t q = lo; q < hi; q++) {
        int cnt = 0;
        const float *yq = c->y + q * dim;
        for (int e = 0; e < mx; e++) c->nbr[q * mx + e] = -1;
        for (int64_t n = c->x0; n < c->x1 && cnt < mx; n++) {
            float d = sqdist(c->x + n * dim, yq, dim);
            if (d < c->r2) c->nbr[q * mx + cnt++] = n;
        }
        c->cnt[q] = cnt;
    }
}

int orc_radius(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
               int B, int dim, double r, int max_nbr, int64_t *nbr, int32_t *cnt) {
    const float r2 = (float)(r * r);
    for (int b = 0; b < B; b++) {
        rad_ctx c = { x, y, ptr_x[b], ptr_x[b + 1], dim, max_nbr, r2, nbr, cnt };
        parallel_for(ptr_y[b], ptr_y[b + 1], 64, radius_range, &c);
    }
    return 0;
}

/* A.11: farthest point sampling, m_b = ceil(ratio * n_b) per example, start = first
 * point, running min-distance seeded with 5e4, arg-max with the LOWEST index on ties
 * (the CPU kernel's order; the CUDA one is thread-order dependent, so this is the pin).
 * out holds global indices, example-major, selection order; out_ptr [B+1]. */
typedef struct { const float *src; const int64_t *ptr, *out_ptr; int dim; int64_t *out; } fps_ctx;

static void fps_range(int64_t lo, int64_t hi, void *vc) {
    const fps_ctx *c = (const fps_ctx *)vc;
    const int dim = c->dim;
    for (int64_t b = lo; b < hi; b++) {
        const int64_t s = c->ptr[b], n = c->ptr[b + 1] - c->ptr[b];
        const int64_t m = c->out_ptr[b + 1] - c->out_ptr[b];
        if (m <= 0) continue;
        float *dist = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        for (int64_t i = 0; i < n; i++) dist[i] = 5e4f;
        int64_t old = 0;
        c->out[c->out_ptr[b]] = s + old;
        for (int64_t j = 1; j < m; j++) {
            float best = -1.0f;
            int64_t besti = 0;
            const float *po = c->src + (s + old) * dim;
            for (int64_t i = 0; i < n; i++) {
                float d = sqdist(c->src + (s + i) * dim, po, dim);
                float v = dist[i] < d ? dist[i] : d;
                dist[i] = v;
                if (v > best) { best = v; besti = i; }
            }
            old = besti;
            c->out[c->out_ptr[b] + j] = s + old;
        }
        free(dist);
    }
}

int orc_fps(const float *src, const int64_t *ptr, int B, int dim, double ratio,
            int64_t *out, int64_t *out_ptr) {
    out_ptr[0] = 0;
    for (int b = 0; b < B; b++) {
        int64_t n = ptr[b + 1] - ptr[b];
        float mf = ceilf((float)n * (float)ratio);      /* deg.to(float) * ratio, ceil */
        out_ptr[b + 1] = out_ptr[b] + (int64_t)mf;
    }
    fps_ctx c = { src, ptr, out_ptr, dim, out };
    parallel_for(0, B, 1, fps_range, &c);
    return 0;
}

/* A.4: c = sum_d (int64)((pos[d]-start[d]) / size[d]) * prod_{d'<d} ((int64)((end-start)/size)+1),
 * FP32 IEEE subtract and divide, truncation toward zero. */
int orc_grid(const float *pos, int dim, const float *size, const float *start,
             const float *end, int64_t n, int64_t *out) {
    for (int64_t i = 0; i < n; i++) {
        int64_t c = 0, k = 1;
        for (int d = 0; d < dim; d++) {
            float p = pos[i * dim + d] - start[d];
            c += (int64_t)(p / size[d]) * k;
            k *= (int64_t)((end[d] - start[d]) / size[d]) + 1;
        }
        out[i] = c;
    }
    return 0;
}
