/*
 * fps_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the two reference hot paths of leonardodalinky/fpsample v1.0.2, written
 * from the algorithm's description with flat arrays; it is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker / reported CPU baseline.
 *
 * Parity pinning: the reference has no tests and no golden vectors, so this file is
 * pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref (the unmodified reference compiled by
 * oracle/build_ref.sh) in tests/test_oracle.py, and the committed fixtures under tests/golden/ that
 * tests/golden/make_golden.py generated from that same build.
 *
 * Arithmetic contract (reference: src/lib.cpp:214-219, src/_ext/Point.h:48-53, src/_ext/utils.h:12-16):
 * every sub / mul / add is an individually rounded binary32 operation, summed over dimensions in
 * index order from 0.0f.  Build with -ffp-contract=off and never with -ffast-math.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_DIM 8 /* src/wrapper.hpp:8-11 (BUCKET_FPS_MAX_DIM) */

/* squared distance, reference order: lib.cpp:214-219 == Point.h:48-53 (float '+' commutes) */
static inline float sqdist(const float *a, const float *b, size_t d) {
    float acc = 0.0f;
    for (size_t j = 0; j < d; ++j) {
        float t = a[j] - b[j];
        acc = acc + t * t;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------------
 * Vanilla FPS.  Follows src/lib.cpp:188-246 (single start) and :111-186 (forced start list):
 *   dist_min = +inf; the first n_starts picks are forced in order; before every pick but the first
 *   the previous pick min-updates all points (strict '<'); free picks take the LAST index holding the
 *   maximum ('>=' scan from max_val = -1, lib.cpp:223-230).
 * returns 0 ok, 2 bad start, 3 bad sizes.
 * ------------------------------------------------------------------------------------------------ */
int oracle_fps_vanilla(const float *pts, size_t n, size_t d, size_t k, const size_t *starts,
                       size_t n_starts, size_t *out) {
    if (n == 0 || d == 0 || k == 0 || k > n || n_starts == 0 || n_starts > k) return 3;
    for (size_t s = 0; s < n_starts; ++s)
        if (starts[s] >= n) return 2;
    float *dm = (float *)malloc(n * sizeof(float));
    if (!dm) return 3;
    for (size_t i = 0; i < n; ++i) dm[i] = INFINITY;
    size_t cur = starts[0];
    out[0] = cur;
    for (size_t t = 1; t < k; ++t) {
        const float *q = pts + cur * d;
        for (size_t i = 0; i < n; ++i) {
            float v = sqdist(pts + i * d, q, d);
            if (v < dm[i]) dm[i] = v;
        }
        if (t < n_starts) {
            cur = starts[t];
        } else {
            float best = -1.0f;
            size_t bi = 0;
            for (size_t i = 0; i < n; ++i)
                if (dm[i] >= best) {
                    best = dm[i];
                    bi = i;
                }
            cur = bi;
        }
        out[t] = cur;
    }
    free(dm);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * FPS with the nearest-point-distance-updating heuristic over an INDEX window (fps_npdu_sampling).
 * Follows src/lib.cpp:272-340: a full min-update against the start point, then per pick a min-update of the
 * points whose index lies within k/2 of the last pick (window shifted, not shrunk, at the array ends) and an
 * arg-max over ALL points with strict '>' from -1 (lowest index among equal maxima).
 * returns 0 ok, 2 bad start, 3 bad sizes.
 * ------------------------------------------------------------------------------------------------ */
int oracle_fps_npdu(const float *pts, size_t n, size_t d, size_t n_samples, size_t k, size_t start, size_t *out) {
    if (n == 0 || d == 0 || n_samples == 0 || n_samples > n) return 3;
    if (start >= n) return 2;
    float *dm = (float *)malloc(n * sizeof(float));
    if (!dm) return 3;
    for (size_t i = 0; i < n; ++i) dm[i] = sqdist(pts + i * d, pts + start * d, d);   /* min(+inf, .) */
    size_t cur = start;
    out[0] = cur;
    const long long P = (long long)n, hw = (long long)(k / 2);
    for (size_t t = 1; t < n_samples; ++t) {
        long long s = (long long)cur - hw, e = (long long)cur + hw;
        if (s < 0) { e -= s; s = 0; }
        if (e >= P) { s = s - (e - P + 1); if (s < 0) s = 0; e = P - 1; }
        const float *q = pts + cur * d;
        for (long long i = s; i <= e; ++i) {
            float v = sqdist(pts + (size_t)i * d, q, d);
            if (v < dm[i]) dm[i] = v;
        }
        float best = -1.0f;
        size_t bi = 0;
        for (size_t i = 0; i < n; ++i)
            if (dm[i] > best) { best = dm[i]; bi = i; }
        cur = bi;
        out[t] = cur;
    }
    free(dm);
    return 0;
}

/*
 * FPS with the nearest-point-distance-updating heuristic over the k NEAREST points (fps_npdu_kdtree_sampling).
 * Reference: fps_npdu_kdtree_sampling_py, src/lib.cpp:369-465.  After the full min-update against the start point every
 * pick min-updates its k_use = min(k, n) nearest points (lib.cpp:407, 421-436; distances as in PointCloud::kdtree_distance,
 * lib.cpp:33-41) and the next pick is the first maximum over all points (strict '>' from -1, lib.cpp:438-442).  The
 * reference obtains the neighbours from nanoflann (vendored third party, src/nanoflann.hpp v1.8.0: KDTreeSingleIndexAdaptor,
 * leaf size 10, KNNResultSet); what it computes only depends on the SET of the k nearest points, which this restatement
 * finds by brute force.  When more points share the k-th nearest distance than there are places left, the reference keeps
 * whichever nanoflann's traversal met first; here they are taken in index order (the one documented difference -- the
 * golden vectors pin the restatement on clouds without such ties).
 */
typedef struct { float d; size_t i; } nk_t;
static int nk_cmp(const void *a, const void *b) {
    const nk_t *x = (const nk_t *)a, *y = (const nk_t *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
int oracle_fps_npdu_kdtree(const float *pts, size_t n, size_t d, size_t n_samples, size_t k, size_t start, size_t *out) {
    if (n == 0 || d == 0 || n_samples == 0 || n_samples > n) return 3;
    if (start >= n) return 2;
    float *dm = (float *)malloc(n * sizeof(float));
    nk_t *nb = (nk_t *)malloc(n * sizeof(nk_t));
    if (!dm || !nb) return 3;
    for (size_t i = 0; i < n; ++i) dm[i] = sqdist(pts + i * d, pts + start * d, d);   /* min(+inf, .), lib.cpp:446-453 */
    size_t cur = start;
    out[0] = cur;
    const size_t k_use = k < n ? k : n;
    for (size_t t = 1; t < n_samples; ++t) {
        /* lib.cpp:412-436: the query is the previous pick (for t == 1 that is the start point again) */
        const float *q = pts + cur * d;
        for (size_t i = 0; i < n; ++i) { nb[i].d = sqdist(pts + i * d, q, d); nb[i].i = i; }
        qsort(nb, n, sizeof(nk_t), nk_cmp);
        for (size_t x = 0; x < k_use; ++x)
            if (nb[x].d < dm[nb[x].i]) dm[nb[x].i] = nb[x].d;
        float best = -1.0f;
        size_t bi = 0;
        for (size_t i = 0; i < n; ++i)
            if (dm[i] > best) { best = dm[i]; bi = i; }
        cur = bi;
        out[t] = cur;
    }
    free(dm);
    free(nb);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * kd-line tree build.  Follows src/_ext/KDTreeBase.h:84-207 + src/_ext/KDLineTree.h:37-39,87-92.
 * Works on a permuted row copy q[n][d] and the permutation perm[n] (perm[pos] = original id).
 * Emits leaves in DFS-left-first order == ascending position order.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    size_t n, d, h;
    float *q;         /* permuted rows            */
    size_t *perm;     /* position -> original id  */
    size_t *leaf_lo;  /* per leaf [lo,hi)         */
    size_t *leaf_hi;
    float *leaf_box;  /* per leaf: d lows then d highs */
    size_t n_leaves;
    float *rowtmp;
} kdl_t;

static void box_of(const kdl_t *t, size_t lo, size_t hi, float *low, float *high) {
    /* KDTreeBase.h:181-207 */
    for (size_t j = 0; j < t->d; ++j) {
        low[j] = FLT_MAX;
        high[j] = -FLT_MAX;
    }
    for (size_t i = lo; i < hi; ++i)
        for (size_t j = 0; j < t->d; ++j) {
            float v = t->q[i * t->d + j];
            if (v < low[j]) low[j] = v;   /* std::min(cur, v) */
            if (high[j] < v) high[j] = v; /* std::max(cur, v) */
        }
}

static void swap_rows(kdl_t *t, size_t a, size_t b) {
    size_t bytes = t->d * sizeof(float);
    memcpy(t->rowtmp, t->q + a * t->d, bytes);
    memcpy(t->q + a * t->d, t->q + b * t->d, bytes);
    memcpy(t->q + b * t->d, t->rowtmp, bytes);
    size_t p = t->perm[a];
    t->perm[a] = t->perm[b];
    t->perm[b] = p;
}

static void divide(kdl_t *t, size_t lo, size_t hi, const float *low, const float *high, size_t depth) {
    size_t count = hi - lo;
    if (depth == t->h || count == 1) { /* KDLineTree.h:37-39 */
        size_t L = t->n_leaves++;
        t->leaf_lo[L] = lo;
        t->leaf_hi[L] = hi;
        memcpy(t->leaf_box + L * 2 * t->d, low, t->d * sizeof(float));
        memcpy(t->leaf_box + L * 2 * t->d + t->d, high, t->d * sizeof(float));
        return;
    }
    /* split dim: first dim with strictly largest span, KDTreeBase.h:160-179 */
    size_t dim = 0;
    float span = 0.0f;
    for (size_t j = 0; j < t->d; ++j) {
        float s = high[j] - low[j];
        if (s > span) {
            span = s;
            dim = j;
        }
    }
    /* split value: strictly sequential binary32 sum in current order, then one binary32 divide,
     * KDTreeBase.h:151-158 (the accumulate lambda narrows the accumulator to float every step) */
    float sum = 0.0f;
    for (size_t i = lo; i < hi; ++i) sum = sum + t->q[i * t->d + dim];
    float val = sum / (float)count;
    /* Hoare partition '< val' | '>= val', KDTreeBase.h:123-149 */
    ptrdiff_t a = (ptrdiff_t)lo, b = (ptrdiff_t)hi - 1;
    for (;;) {
        while (a <= b && t->q[(size_t)a * t->d + dim] < val) ++a;
        while (a <= b && t->q[(size_t)b * t->d + dim] >= val) --b;
        if (a > b) break;
        swap_rows(t, (size_t)a, (size_t)b);
        ++a;
        --b;
    }
    size_t nleft = (size_t)a - lo;
    if ((size_t)a == lo) nleft = 1;
    if ((size_t)a == hi) nleft = count - 1;
    float *cl = (float *)malloc(2 * t->d * sizeof(float));
    box_of(t, lo, lo + nleft, cl, cl + t->d);
    divide(t, lo, lo + nleft, cl, cl + t->d, depth + 1);
    box_of(t, lo + nleft, hi, cl, cl + t->d);
    divide(t, lo + nleft, hi, cl, cl + t->d, depth + 1);
    free(cl);
}

static kdl_t *kdl_build(const float *pts, size_t n, size_t d, size_t h) {
    kdl_t *t = (kdl_t *)calloc(1, sizeof(kdl_t));
    t->n = n;
    t->d = d;
    t->h = h;
    t->q = (float *)malloc(n * d * sizeof(float));
    t->perm = (size_t *)malloc(n * sizeof(size_t));
    size_t max_leaves = (h < 40 && ((size_t)1 << h) < n) ? ((size_t)1 << h) : n;
    t->leaf_lo = (size_t *)malloc(max_leaves * sizeof(size_t));
    t->leaf_hi = (size_t *)malloc(max_leaves * sizeof(size_t));
    t->leaf_box = (float *)malloc(max_leaves * 2 * d * sizeof(float));
    t->rowtmp = (float *)malloc(d * sizeof(float));
    memcpy(t->q, pts, n * d * sizeof(float));
    for (size_t i = 0; i < n; ++i) t->perm[i] = i;
    float box[2 * ORACLE_MAX_DIM];
    box_of(t, 0, n, box, box + d);
    divide(t, 0, n, box, box + d, 0);
    return t;
}

static void kdl_free(kdl_t *t) {
    free(t->q);
    free(t->perm);
    free(t->leaf_lo);
    free(t->leaf_hi);
    free(t->leaf_box);
    free(t->rowtmp);
    free(t);
}

/* exported: build only.  perm[n]; leaf_bounds[n_leaves+1]; leaf_box[n_leaves][2][d] (may be NULL). */
int oracle_kdline_build(const float *pts, size_t n, size_t d, size_t h, size_t *perm,
                        size_t *leaf_bounds, float *leaf_box, size_t *n_leaves) {
    if (d == 0 || d > ORACLE_MAX_DIM) return 1;
    if (n == 0 || h == 0) return 3;
    kdl_t *t = kdl_build(pts, n, d, h);
    memcpy(perm, t->perm, n * sizeof(size_t));
    for (size_t L = 0; L < t->n_leaves; ++L) leaf_bounds[L] = t->leaf_lo[L];
    leaf_bounds[t->n_leaves] = n;
    if (leaf_box) memcpy(leaf_box, t->leaf_box, t->n_leaves * 2 * d * sizeof(float));
    *n_leaves = t->n_leaves;
    kdl_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * kd-line sampling, LAZY form: follows src/wrapper.hpp:45-60, src/_ext/KDNode.h:84-166 and
 * src/_ext/KDLineTree.h:56-85.  Points never move after the build, so a deferred reference point is
 * remembered by its position.  stats (may be NULL): [0] point-updates, [1] bucket tests,
 * [2] leaf flushes, [3] deferred, [4] dropped.
 * returns 0 ok, 1 bad dim, 2 bad start (wrapper.hpp:121-127), 3 bad sizes.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    size_t *v;
    size_t len, cap;
} vec_t;

static void vec_push(vec_t *a, size_t x) {
    if (a->len == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 8;
        a->v = (size_t *)realloc(a->v, a->cap * sizeof(size_t));
    }
    a->v[a->len++] = x;
}

/* one full scan of a leaf against reference row r: min-update, first strict max (KDNode.h:151-160) */
static void leaf_scan(const kdl_t *t, float *dis, size_t lo, size_t hi, const float *r, size_t *mpos,
                      float *mdis) {
    float best = -FLT_MAX;
    size_t bp = *mpos;
    for (size_t i = lo; i < hi; ++i) {
        float v = sqdist(t->q + i * t->d, r, t->d);
        float cur = dis[i];
        cur = (v < cur) ? v : cur; /* std::min(dis, v), Point.h:82-86 */
        dis[i] = cur;
        if (cur > best) {
            best = cur;
            bp = i;
        }
    }
    *mpos = bp;
    *mdis = best;
}

int oracle_kdline_sample(const float *pts, size_t n, size_t d, size_t k, size_t start, size_t h,
                         size_t *out, uint64_t *stats) {
    if (d == 0 || d > ORACLE_MAX_DIM) return 1;
    if (start >= n) return 2;
    if (n == 0 || k == 0 || k > n || h == 0) return 3;
    kdl_t *t = kdl_build(pts, n, d, h);
    size_t nl = t->n_leaves;
    float *dis = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) dis[i] = FLT_MAX; /* Point.h:61-65 */
    size_t *mpos = (size_t *)calloc(nl, sizeof(size_t));
    float *mdis = (float *)calloc(nl, sizeof(float));
    vec_t *delay = (vec_t *)calloc(nl, sizeof(vec_t));
    uint64_t st[5] = {0, 0, 0, 0, 0};

    /* init with the point sitting at POSITION start after the build (wrapper.hpp:54-55) */
    size_t ref = start;
    out[0] = t->perm[ref];
    for (size_t L = 0; L < nl; ++L) {
        leaf_scan(t, dis, t->leaf_lo[L], t->leaf_hi[L], t->q + ref * d, &mpos[L], &mdis[L]);
        st[0] += t->leaf_hi[L] - t->leaf_lo[L];
    }
    for (size_t s = 1; s < k; ++s) {
        /* KDLineTree.h:56-67: strict '>' over buckets in leaf order */
        float best = -FLT_MAX;
        size_t bp = 0;
        for (size_t L = 0; L < nl; ++L)
            if (mdis[L] > best) {
                best = mdis[L];
                bp = mpos[L];
            }
        ref = bp;
        out[s] = t->perm[ref];
        const float *r = t->q + ref * d;
        /* KDLineTree.h:69-75 + KDNode.h:120-166 (leaf branch) */
        for (size_t L = 0; L < nl; ++L) {
            float lastmax = mdis[L];
            float cur = sqdist(t->q + mpos[L] * d, r, d);
            st[1]++;
            if (cur > lastmax) {
                /* KDNode.h:105-118 */
                const float *low = t->leaf_box + L * 2 * d, *high = low + d;
                float bound = 0.0f;
                for (size_t j = 0; j < d; ++j) {
                    float e = 0.0f;
                    if (r[j] > high[j])
                        e = r[j] - high[j];
                    else if (r[j] < low[j])
                        e = low[j] - r[j];
                    bound = bound + e * e;
                }
                if (bound < lastmax) {
                    vec_push(&delay[L], ref);
                    st[3]++;
                } else {
                    st[4]++;
                }
            } else {
                vec_push(&delay[L], ref);
                for (size_t x = 0; x < delay[L].len; ++x) {
                    leaf_scan(t, dis, t->leaf_lo[L], t->leaf_hi[L], t->q + delay[L].v[x] * d, &mpos[L],
                              &mdis[L]);
                    st[0] += t->leaf_hi[L] - t->leaf_lo[L];
                }
                delay[L].len = 0;
                st[2]++;
            }
        }
    }
    if (stats) memcpy(stats, st, sizeof(st));
    for (size_t L = 0; L < nl; ++L) free(delay[L].v);
    free(delay);
    free(mpos);
    free(mdis);
    free(dis);
    kdl_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * kd-line sampling, EAGER form: exact FPS over the permuted array started at
 * position `start`, dis init FLT_MAX, ties -> lowest position.  O(n*k); used to show that the lazy
 * bucket bookkeeping above is observationally an accelerator only.
 * ------------------------------------------------------------------------------------------------ */
int oracle_kdline_sample_eager(const float *pts, size_t n, size_t d, size_t k, size_t start, size_t h,
                               size_t *out) {
    if (d == 0 || d > ORACLE_MAX_DIM) return 1;
    if (start >= n) return 2;
    if (n == 0 || k == 0 || k > n || h == 0) return 3;
    kdl_t *t = kdl_build(pts, n, d, h);
    float *dis = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) dis[i] = FLT_MAX;
    size_t ref = start;
    out[0] = t->perm[ref];
    for (size_t s = 1; s < k; ++s) {
        const float *r = t->q + ref * d;
        float best = -FLT_MAX;
        size_t bp = 0;
        for (size_t i = 0; i < n; ++i) {
            float v = sqdist(t->q + i * d, r, d);
            float cur = dis[i];
            cur = (v < cur) ? v : cur;
            dis[i] = cur;
            if (cur > best) {
                best = cur;
                bp = i;
            }
        }
        ref = bp;
        out[s] = t->perm[ref];
    }
    free(dis);
    kdl_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Full-size CERTIFIER (test infrastructure).  Given a claimed pick sequence out[k], replays the
 * recurrence of src/lib.cpp:211-238 (mode 0: dm = +inf, ties -> highest index) or of SURVEY.md A.4
 * (mode 1: rows are already in permuted order, dm = FLT_MAX, ties -> lowest position) along the
 * CLAIMED picks and checks that every free pick is the arg-max of the running min-distances.  By
 * induction over t this holds iff out[] equals what the sequential reference loop produces, but the
 * point range can be cut into independent chunks (the picks are known), so it runs on all host cores:
 * each worker keeps its chunk's dm[] cache-resident for all k rounds and records its chunk-local
 * winner per round; the winners are merged at the end.
 * returns 0 = certified, 1 = mismatch (first bad round in *bad_round), 3 = bad sizes.
 * ------------------------------------------------------------------------------------------------ */
#include <pthread.h>

typedef struct {
    const float *pts;
    size_t n, d, k, lo, hi;
    const size_t *picks;
    int mode;
    float *best_val;  /* [k] chunk-local winner value per round (round 0 unused) */
    size_t *best_idx; /* [k] */
} cert_job_t;

static void *cert_worker(void *arg) {
    cert_job_t *j = (cert_job_t *)arg;
    const size_t d = j->d, cnt = j->hi - j->lo;
    float *dm = (float *)malloc((cnt ? cnt : 1) * sizeof(float));
    for (size_t i = 0; i < cnt; ++i) dm[i] = j->mode == 0 ? INFINITY : FLT_MAX;
    for (size_t t = 1; t < j->k; ++t) {
        const float *q = j->pts + j->picks[t - 1] * d;
        float best = -1.0f;
        size_t bi = j->lo;
        for (size_t i = 0; i < cnt; ++i) {
            float v = sqdist(j->pts + (j->lo + i) * d, q, d);
            float cur = dm[i];
            cur = (v < cur) ? v : cur;
            dm[i] = cur;
            if (j->mode == 0 ? (cur >= best) : (cur > best)) {
                best = cur;
                bi = j->lo + i;
            }
        }
        j->best_val[t] = best;
        j->best_idx[t] = bi;
    }
    free(dm);
    return NULL;
}

int oracle_certify_fps(const float *pts, size_t n, size_t d, size_t k, const size_t *picks,
                       size_t n_forced, int mode, int n_threads, size_t *bad_round) {
    if (n == 0 || d == 0 || k == 0 || k > n || n_forced == 0) return 3;
    for (size_t t = 0; t < k; ++t)
        if (picks[t] >= n) {
            if (bad_round) *bad_round = t;
            return 1;
        }
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n) n_threads = (int)n;
    cert_job_t *jobs = (cert_job_t *)calloc((size_t)n_threads, sizeof(cert_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    size_t chunk = (n + (size_t)n_threads - 1) / (size_t)n_threads;
    for (int w = 0; w < n_threads; ++w) {
        cert_job_t *j = &jobs[w];
        j->pts = pts;
        j->n = n;
        j->d = d;
        j->k = k;
        j->lo = (size_t)w * chunk < n ? (size_t)w * chunk : n;
        j->hi = j->lo + chunk < n ? j->lo + chunk : n;
        j->picks = picks;
        j->mode = mode;
        j->best_val = (float *)malloc(k * sizeof(float));
        j->best_idx = (size_t *)malloc(k * sizeof(size_t));
        pthread_create(&th[w], NULL, cert_worker, j);
    }
    for (int w = 0; w < n_threads; ++w) pthread_join(th[w], NULL);
    int rc = 0;
    for (size_t t = n_forced; t < k && rc == 0; ++t) {
        float best = -1.0f;
        size_t bi = 0;
        for (int w = 0; w < n_threads; ++w) { /* chunks in ascending index order */
            if (jobs[w].hi == jobs[w].lo) continue;
            float v = jobs[w].best_val[t];
            if (mode == 0 ? (v >= best) : (v > best)) {
                best = v;
                bi = jobs[w].best_idx[t];
            }
        }
        if (bi != picks[t]) {
            rc = 1;
            if (bad_round) *bad_round = t;
        }
    }
    for (int w = 0; w < n_threads; ++w) {
        free(jobs[w].best_val);
        free(jobs[w].best_idx);
    }
    free(jobs);
    free(th);
    return rc;
}
