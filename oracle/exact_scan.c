/*
 * oracle/exact_scan.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the arithmetic on the hybrid-retrieval hot path of
 * rnaarla/advanced-rag-milvus.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (advanced-rag-milvus_b200/) never does.
 *
 * PARITY STATUS
 *   * dense / sparse search (rows S1, S2 of SURVEY.md section 8a): "parity unpinned".  In the reference
 *     the scoring happens inside a remote Milvus 2.3.3 server reached through pymilvus
 *     (reference src/advanced_rag/indexing.py:505-523, docker-compose.yml:41); neither is present under
 *     /root/reference and no reference test pins a dense or sparse score.  What is restated here is the
 *     published semantics of those two calls as the reference configures them:
 *       - semantic_index / domain_index: metric_type COSINE (indexing.py:143-180, retrieval.py:93-96),
 *         i.e. inner product of L2-normalised rows; we run it as an EXACT flat scan.
 *       - sparse_index: metric_type IP over SPARSE_FLOAT_VECTOR (indexing.py:163-164,
 *         retrieval.py:97-101), i.e. sum over shared term ids of query value * document value.
 *     and the rule this repo states for ties: score descending, then integer row id ascending.
 *   * RRF fusion / MMR (rows F1, F2): pinned -- see oracle/fusion.py, which is checked against golden
 *     vectors produced by executing the reference's own _fuse_results / _mmr_diversify.
 *
 * CANONICAL ARITHMETIC (shared, bit for bit, with the CUDA path)
 *   dense score  s(q,x) = sum_d q[d]*x[d] over the STORED 16-bit values (fp16 or bf16), each product
 *                exact in fp64, accumulated in fp64 in 8 interleaved lanes (lane j takes d = j mod 8 in
 *                increasing d) and combined as ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)).  D is treated as
 *                zero-padded to a multiple of 8.
 *   cosine       rows and queries are normalised ONCE: n2 = sum_d (double)x[d]^2 sequentially in d,
 *                y[d] = round_to_nearest_even_16bit( (double)x[d] / sqrt(n2) )  (all-zero rows stay zero);
 *                the score is then the dense score above of the stored values.
 *   sparse score acc = fmaf(qv_t, w_td, acc) in fp32, query terms taken in ascending term id.
 *   top-k        (score desc, id asc); dense returns min(k,N) hits, sparse only documents that share at
 *                least one term with the query.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_F16 0
#define ORC_BF16 1

/* ------------------------------------------------------------------ 16-bit <-> fp64 conversions */

static inline double f16_bits_to_double(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15);
    int32_t e = (h >> 10) & 0x1f;
    uint32_t m = h & 0x3ff;
    double v;
    if (e == 0) v = ldexp((double)m, -24);                 /* zero / subnormal */
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexp((double)(m | 0x400), e - 25);
    return sign ? -v : v;
}

static inline double bf16_bits_to_double(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
}

double orc_bits_to_double(uint16_t h, int dtype) {
    return dtype == ORC_F16 ? f16_bits_to_double(h) : bf16_bits_to_double(h);
}

/* Round a double to the nearest-even value of a binary format with `mbits` stored mantissa bits,
 * minimum normal exponent emin (unbiased) and maximum exponent emax; returns the rounded double. */
static inline double round_to_format(double v, int mbits, int emin, int emax, int *overflow) {
    *overflow = 0;
    if (v == 0.0 || isnan(v) || isinf(v)) return v;
    int e;
    double a = fabs(v);
    (void)frexp(a, &e);          /* a = f * 2^e, f in [0.5,1)  => unbiased exponent = e-1 */
    int ue = e - 1;
    if (ue < emin) ue = emin;    /* subnormal range shares the quantum of emin */
    double quantum = ldexp(1.0, ue - mbits);
    double r = nearbyint(a / quantum) * quantum;   /* default rounding mode = nearest even; a/quantum exact */
    if (r >= ldexp(1.0, emax + 1)) *overflow = 1;
    return v < 0 ? -r : r;
}

static inline uint16_t double_to_f16_bits(double v) {
    uint16_t sign = signbit(v) ? 0x8000 : 0;
    if (isnan(v)) return 0x7e00;
    int ovf;
    double r = fabs(round_to_format(v, 10, -14, 15, &ovf));
    if (isinf(v) || ovf) return sign | 0x7c00;
    if (r == 0.0) return sign;
    int e;
    (void)frexp(r, &e);
    int ue = e - 1;
    if (ue < -14) return sign | (uint16_t)llround(ldexp(r, 24));       /* subnormal */
    uint32_t m = (uint32_t)llround(ldexp(r, 10 - ue)) & 0x3ff;
    return sign | (uint16_t)((ue + 15) << 10) | (uint16_t)m;
}

static inline uint16_t double_to_bf16_bits(double v) {
    uint16_t sign = signbit(v) ? 0x8000 : 0;
    if (isnan(v)) return 0x7fc0;
    int ovf;
    double r = fabs(round_to_format(v, 7, -126, 127, &ovf));
    if (isinf(v) || ovf) return sign | 0x7f80;
    float f = (float)r;           /* exact: r has <= 8 significant bits inside fp32 range */
    uint32_t u;
    memcpy(&u, &f, 4);
    return sign | (uint16_t)(u >> 16);
}

uint16_t orc_double_to_bits(double v, int dtype) {
    return dtype == ORC_F16 ? double_to_f16_bits(v) : double_to_bf16_bits(v);
}

void orc_round_f32(const float *in, uint16_t *out, int64_t n, int dtype) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = orc_double_to_bits((double)in[i], dtype);
}

void orc_bits_to_f32(const uint16_t *in, float *out, int64_t n, int dtype) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = (float)orc_bits_to_double(in[i], dtype);
}

/* Canonical row normalisation (cosine metric): see header. */
void orc_normalize_rows(const float *in, uint16_t *out, int64_t n_rows, int dim, int dtype) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) {
        const float *x = in + r * (int64_t)dim;
        uint16_t *y = out + r * (int64_t)dim;
        double n2 = 0.0;
        for (int d = 0; d < dim; ++d) n2 += (double)x[d] * (double)x[d];
        double nrm = sqrt(n2);
        for (int d = 0; d < dim; ++d)
            y[d] = nrm > 0.0 ? orc_double_to_bits((double)x[d] / nrm, dtype) : 0;
    }
}

/* ------------------------------------------------------------------ canonical dense score */

static inline double dot8(const double *q, const double *x, int dim8) {
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int d = 0; d < dim8; d += 8)
        for (int j = 0; j < 8; ++j) p[j] += q[d + j] * x[d + j];   /* products exact -> fma == mul+add */
    return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}

typedef struct { double s; int64_t id; } hit_t;

/* a ranks before b under (score desc, id asc) */
static inline int hit_before(hit_t a, hit_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }

/* heap[0] is the WORST retained hit */
static inline void heap_sift_down(hit_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && hit_before(h[w], h[l])) w = l;
        if (r < n && hit_before(h[w], h[r])) w = r;
        if (w == i) return;
        hit_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}
static inline void heap_offer(hit_t *h, int *n, int k, hit_t x) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = x;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (hit_before(h[p], h[i])) { hit_t t = h[p]; h[p] = h[i]; h[i] = t; i = p; } else break;
        }
    } else if (hit_before(x, h[0])) {
        h[0] = x;
        heap_sift_down(h, k, 0);
    }
}
static int hit_cmp(const void *a, const void *b) {
    hit_t x = *(const hit_t *)a, y = *(const hit_t *)b;
    return hit_before(x, y) ? -1 : (hit_before(y, x) ? 1 : 0);
}

/* Scores of ONE query against all rows (fp64, canonical). */
void orc_dense_scores(const uint16_t *corpus, int64_t n_rows, int dim, int dtype,
                      const uint16_t *query, double *out) {
    int dim8 = (dim + 7) & ~7;
    double *q = (double *)calloc(dim8, sizeof(double));
    for (int d = 0; d < dim; ++d) q[d] = orc_bits_to_double(query[d], dtype);
#pragma omp parallel
    {
        double *x = (double *)calloc(dim8, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n_rows; ++r) {
            const uint16_t *row = corpus + r * (int64_t)dim;
            for (int d = 0; d < dim; ++d) x[d] = orc_bits_to_double(row[d], dtype);
            out[r] = dot8(q, x, dim8);
        }
        free(x);
    }
    free(q);
}

/*
 * Exact flat top-k.  corpus [n_rows, dim] and queries [n_q, dim] hold 16-bit patterns of `dtype`.
 * out_scores f64 [n_q,k], out_ids i64 [n_q,k]; unused slots: score -inf, id -1.  Returns 0.
 * Follows the boundary call reference src/advanced_rag/indexing.py:505-523 (one ranked hit list per
 * query vector, best first) with the exact-scan semantics stated in the header.
 */
int orc_dense_topk(const uint16_t *corpus, int64_t n_rows, int dim, int dtype,
                   const uint16_t *queries, int n_q, int k, int64_t id_offset,
                   double *out_scores, int64_t *out_ids) {
    if (k <= 0 || dim <= 0 || n_q < 0 || n_rows < 0) return -1;
    int dim8 = (dim + 7) & ~7;
    double *qd = (double *)calloc((size_t)n_q * dim8, sizeof(double));
    for (int b = 0; b < n_q; ++b)
        for (int d = 0; d < dim; ++d)
            qd[(size_t)b * dim8 + d] = orc_bits_to_double(queries[(size_t)b * dim + d], dtype);

    int n_thr = 1;
#ifdef _OPENMP
    n_thr = omp_get_max_threads();
#endif
    hit_t *heaps = (hit_t *)malloc((size_t)n_thr * n_q * k * sizeof(hit_t));
    int *cnt = (int *)calloc((size_t)n_thr * n_q, sizeof(int));
    enum { RB = 32 };
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        hit_t *my = heaps + (size_t)t * n_q * k;
        int *myc = cnt + (size_t)t * n_q;
        double *xb = (double *)calloc((size_t)RB * dim8, sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int64_t r0 = 0; r0 < n_rows; r0 += RB) {
            int nr = (int)((n_rows - r0) < RB ? (n_rows - r0) : RB);
            for (int i = 0; i < nr; ++i) {
                const uint16_t *row = corpus + (r0 + i) * (int64_t)dim;
                for (int d = 0; d < dim; ++d) xb[(size_t)i * dim8 + d] = orc_bits_to_double(row[d], dtype);
            }
            for (int b = 0; b < n_q; ++b) {
                const double *q = qd + (size_t)b * dim8;
                for (int i = 0; i < nr; ++i) {
                    hit_t h = { dot8(q, xb + (size_t)i * dim8, dim8), id_offset + r0 + i };
                    heap_offer(my + (size_t)b * k, &myc[b], k, h);
                }
            }
        }
        free(xb);
    }
    hit_t *all = (hit_t *)malloc((size_t)n_thr * k * sizeof(hit_t));
    for (int b = 0; b < n_q; ++b) {
        int n = 0;
        for (int t = 0; t < n_thr; ++t) {
            int c = cnt[(size_t)t * n_q + b];
            memcpy(all + n, heaps + ((size_t)t * n_q + b) * k, (size_t)c * sizeof(hit_t));
            n += c;
        }
        qsort(all, n, sizeof(hit_t), hit_cmp);
        for (int j = 0; j < k; ++j) {
            out_scores[(size_t)b * k + j] = j < n ? all[j].s : -INFINITY;
            out_ids[(size_t)b * k + j] = j < n ? all[j].id : -1;
        }
    }
    free(all); free(cnt); free(heaps); free(qd);
    return 0;
}

/* ------------------------------------------------------------------ sparse inner-product top-k */

/*
 * Postings layout (term-major CSR, i.e. the transpose of the doc-major CSR the reference assembles at
 * indexing.py:379-404): term_ptr i64 [V+1], post_doc i32 [nnz] ascending inside a term, post_w f32 [nnz].
 * Queries: q_ptr i64 [n_q+1], q_terms i32 ascending inside a query, q_vals f32.
 * Output: out_scores f32 [n_q,k], out_ids i64 [n_q,k] (-inf / -1 padded), out_counts i32 [n_q].
 * Follows the sparse branch of the boundary call (indexing.py:472,487-498: dict{indices,values} -> 1-row
 * CSR, metric IP).
 */
int orc_sparse_topk(const int64_t *term_ptr, const int32_t *post_doc, const float *post_w,
                    int64_t n_docs, int32_t n_terms,
                    const int64_t *q_ptr, const int32_t *q_terms, const float *q_vals,
                    int n_q, int k, int64_t id_offset,
                    float *out_scores, int64_t *out_ids, int32_t *out_counts) {
    if (k <= 0) return -1;
    int err = 0;
#pragma omp parallel
    {
        float *acc = (float *)calloc((size_t)n_docs, sizeof(float));
        uint8_t *touched = (uint8_t *)calloc((size_t)n_docs, 1);
        int32_t *tlist = (int32_t *)malloc((size_t)(n_docs > 0 ? n_docs : 1) * sizeof(int32_t));
        hit_t *heap = (hit_t *)malloc((size_t)k * sizeof(hit_t));
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < n_q; ++b) {
            int64_t nt = 0;
            for (int64_t j = q_ptr[b]; j < q_ptr[b + 1]; ++j) {
                int32_t t = q_terms[j];
                if (t < 0 || t >= n_terms) { err = -2; continue; }
                float qv = q_vals[j];
                for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
                    int32_t d = post_doc[p];
                    acc[d] = fmaf(qv, post_w[p], acc[d]);
                    if (!touched[d]) { touched[d] = 1; tlist[nt++] = d; }
                }
            }
            int n = 0;
            for (int64_t i = 0; i < nt; ++i) {
                int32_t d = tlist[i];
                hit_t h = { (double)acc[d], id_offset + d };
                heap_offer(heap, &n, k, h);
                acc[d] = 0.0f;
                touched[d] = 0;
            }
            qsort(heap, n, sizeof(hit_t), hit_cmp);
            for (int j = 0; j < k; ++j) {
                out_scores[(size_t)b * k + j] = j < n ? (float)heap[j].s : -INFINITY;
                out_ids[(size_t)b * k + j] = j < n ? heap[j].id : -1;
            }
            out_counts[b] = n;
        }
        free(heap); free(tlist); free(touched); free(acc);
    }
    return err;
}

/* ------------------------------------------------------------------ k-way merge of ranked lists */

/* cand_scores f64 [n_q, n_cand], cand_ids i64 [n_q, n_cand]; id < 0 marks an empty slot.
 * The step a sharded deployment needs after gathering per-shard top-k (Milvus' proxy does the same
 * reduce server-side across its num_shards=4, reference indexing.py:91,234-239). */
int orc_merge_topk(const double *cand_scores, const int64_t *cand_ids, int n_q, int n_cand, int k,
                   double *out_scores, int64_t *out_ids) {
#pragma omp parallel
    {
        hit_t *buf = (hit_t *)malloc((size_t)(n_cand > 0 ? n_cand : 1) * sizeof(hit_t));
#pragma omp for schedule(static)
        for (int b = 0; b < n_q; ++b) {
            int n = 0;
            for (int j = 0; j < n_cand; ++j) {
                int64_t id = cand_ids[(size_t)b * n_cand + j];
                if (id >= 0) { buf[n].s = cand_scores[(size_t)b * n_cand + j]; buf[n].id = id; ++n; }
            }
            qsort(buf, n, sizeof(hit_t), hit_cmp);
            for (int j = 0; j < k; ++j) {
                out_scores[(size_t)b * k + j] = j < n ? buf[j].s : -INFINITY;
                out_ids[(size_t)b * k + j] = j < n ? buf[j].id : -1;
            }
        }
        free(buf);
    }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
