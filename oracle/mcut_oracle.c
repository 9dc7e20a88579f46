/* oracle/mcut_oracle.c — see mcut_oracle.h.  TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline
 * "port"); the product under mcut_b200/ never links or calls this file.
 *
 * Written from the behaviour of the reference (cutdigital/mcut) as described in SURVEY.md §8; the
 * expansion arithmetic follows J. R. Shewchuk, "Adaptive Precision Floating-Point Arithmetic and Fast
 * Robust Geometric Predicates" (1997), whose constants the reference hard-codes at
 * source/shewchuk.c:420-433.  Pinned bit-for-bit against oracle/_ref (tests/golden/).
 */
#include "mcut_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================================================
 * a1  coordinate re-centring — source/preproc.cpp:2124-2290 and :91-185
 * ============================================================================================== */

static void bbox_and_mean(int is_float, const void* verts, uint32_t n, double mn[3], double mx[3], double com[3])
{
    /* preproc.cpp:2143-2196: the mean is a plain left-to-right sum (order matters for the bits);
     * for float input the running min/max are squeezed through float (static_cast<float>(bbox)). */
    for (int j = 0; j < 3; ++j) {
        mn[j] = DBL_MAX;
        mx[j] = -DBL_MAX; /* numeric_limits<double>::lowest() */
        com[j] = 0.0;
    }
    if (is_float) {
        const float* p = (const float*)verts;
        for (uint32_t v = 0; v < n; ++v)
            for (int j = 0; j < 3; ++j) {
                const float c = p[3 * (size_t)v + j];
                const float a = (float)mx[j];
                const float b = (float)mn[j];
                mx[j] = (double)((a < c) ? c : a); /* std::max(a, c) */
                mn[j] = (double)((c < b) ? c : b); /* std::min(b, c) */
                com[j] += (double)c;
            }
    } else {
        const double* p = (const double*)verts;
        for (uint32_t v = 0; v < n; ++v)
            for (int j = 0; j < 3; ++j) {
                const double c = p[3 * (size_t)v + j];
                mx[j] = (mx[j] < c) ? c : mx[j];
                mn[j] = (c < mn[j]) ? c : mn[j];
                com[j] += c;
            }
    }
    for (int j = 0; j < 3; ++j) com[j] = com[j] / (double)n;
}

void mco_vertex_parameters(int is_float, const void* src, uint32_t nsv, const void* cut, uint32_t ncv, double com[3],
    double shift[3], double src_bbox[6], double cut_bbox[6])
{
    double smn[3], smx[3], scom[3], cmn[3], cmx[3], ccom[3];
    bbox_and_mean(is_float, src, nsv, smn, smx, scom);
    bbox_and_mean(is_float, cut, ncv, cmn, cmx, ccom);
    double tpq[3];
    for (int j = 0; j < 3; ++j) {
        com[j] = (scom[j] + ccom[j]) / 2.0; /* :2215 */
        const double allmin = (cmn[j] < smn[j]) ? cmn[j] : smn[j]; /* compwise_min(src, cut) */
        tpq[j] = com[j] - allmin; /* :2221 */
    }
    /* :2222-2225 offset_from_origin = normalize(to_positive_quadrant); shift = tpq + offset */
    double len2 = 0.0;
    for (int j = 0; j < 3; ++j) len2 += tpq[j] * tpq[j];
    const double len = sqrt(len2);
    for (int j = 0; j < 3; ++j) shift[j] = tpq[j] + (tpq[j] / len);
    for (int j = 0; j < 3; ++j) {
        src_bbox[j] = smn[j] + shift[j];
        src_bbox[3 + j] = smx[j] + shift[j];
        cut_bbox[j] = cmn[j] + shift[j];
        cut_bbox[3 + j] = cmx[j] + shift[j];
    }
}

void mco_transform_vertices(int is_float, const void* in, uint32_t nv, const double com[3], const double shift[3],
    const double* perturbation, double* out)
{
    if (is_float) {
        /* preproc.cpp:124-134: subtraction and addition happen in float with (float)com, (float)shift */
        const float* p = (const float*)in;
        const float fc[3] = { (float)com[0], (float)com[1], (float)com[2] };
        const float fs[3] = { (float)shift[0], (float)shift[1], (float)shift[2] };
        for (uint32_t v = 0; v < nv; ++v)
            for (int j = 0; j < 3; ++j) {
                const float d = p[3 * (size_t)v + j] - fc[j];
                const float x = d + fs[j];
                out[3 * (size_t)v + j] = (double)x + (perturbation ? perturbation[j] : 0.0);
            }
    } else {
        const double* p = (const double*)in;
        for (uint32_t v = 0; v < nv; ++v)
            for (int j = 0; j < 3; ++j) {
                const double x = (p[3 * (size_t)v + j] - com[j]) + shift[j]; /* :166-168 */
                out[3 * (size_t)v + j] = x + (perturbation ? perturbation[j] : 0.0);
            }
    }
}

double mco_cut_bbox_eps(const double cut_bbox[6], double gp_constant, int absolute)
{
    /* preproc.cpp:2518 length(cutmesh_bboxmax - cutmesh_bboxmin), :2667-2675 */
    double s = 0.0;
    for (int j = 0; j < 3; ++j) {
        const double d = cut_bbox[3 + j] - cut_bbox[j];
        s += d * d;
    }
    double scalar = sqrt(s);
    if (absolute) scalar = 1.0;
    return scalar * gp_constant;
}

/* ================================================================================================
 * a2/a3  face AABBs, mesh AABB, Morton codes — source/bvh.cpp:196-217, :242-433; math.h:866-928
 * ============================================================================================== */

void mco_face_bboxes(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, double eps,
    double* bboxes, double root[6])
{
    mco_face_bboxes_prior(xyz, face_off, face_vtx, nf, eps, NULL, 0, bboxes, root);
}

/* build_oibvh() resizes the caller's face_bboxes and expands what it finds there (bvh.cpp:242, :253-264); preproc.cpp keeps
 * one such vector per mesh for a whole mcDispatch (:2453, never cleared), so the rebuild after a floating-polygon repartition
 * (:2733-2760) starts faces [0, n_prior) from the box they had, and enlarges it again. */
void mco_face_bboxes_prior(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, double eps,
    const double* prior, uint32_t n_prior, double* bboxes, double root[6])
{
    for (int j = 0; j < 3; ++j) {
        root[j] = DBL_MAX;
        root[3 + j] = -DBL_MAX;
    }
    for (uint32_t f = 0; f < nf; ++f) {
        double* b = bboxes + 6 * (size_t)f;
        for (int j = 0; j < 3; ++j) {
            b[j] = f < n_prior ? prior[6 * (size_t)f + j] : DBL_MAX;
            b[3 + j] = f < n_prior ? prior[6 * (size_t)f + 3 + j] : -DBL_MAX;
        }
        for (uint32_t h = face_off[f]; h < face_off[f + 1]; ++h) {
            const double* p = xyz + 3 * (size_t)face_vtx[h];
            for (int j = 0; j < 3; ++j) {
                b[3 + j] = (b[3 + j] < p[j]) ? p[j] : b[3 + j];
                b[j] = (p[j] < b[j]) ? p[j] : b[j];
            }
        }
        if (eps > 0.0) /* bvh.cpp:268-272, math.h:906-910 */
            for (int j = 0; j < 3; ++j) {
                b[3 + j] = b[3 + j] + eps;
                b[j] = b[j] - eps;
            }
        for (int j = 0; j < 3; ++j) { /* bvh.cpp:318-343: root = union of the (enlarged) face boxes */
            root[3 + j] = (root[3 + j] < b[3 + j]) ? b[3 + j] : root[3 + j];
            root[j] = (b[j] < root[j]) ? b[j] : root[j];
        }
    }
}

static uint32_t spread10(uint32_t v)
{
    /* 10 bits -> every third bit (bvh.cpp:196-203) */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

uint32_t mco_morton3D(float x, float y, float z)
{
    /* bvh.cpp:206-217 */
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    return spread10((uint32_t)x) * 4u + spread10((uint32_t)y) * 2u + spread10((uint32_t)z);
}

void mco_morton_codes(const double* bboxes, uint32_t nf, const double root[6], uint32_t* codes)
{
    /* bvh.cpp:275 centre = (min + max) / 2; :382-402 offset = centre - root.min; code of (float)(offset/dims) */
    double dims[3];
    for (int j = 0; j < 3; ++j) dims[j] = root[3 + j] - root[j];
    for (uint32_t f = 0; f < nf; ++f) {
        const double* b = bboxes + 6 * (size_t)f;
        float n[3];
        for (int j = 0; j < 3; ++j) {
            const double c = (b[j] + b[3 + j]) / 2;
            const double off = c - root[j];
            n[j] = (float)(off / dims[j]);
        }
        codes[f] = mco_morton3D(n[0], n[1], n[2]);
    }
}

/* ================================================================================================
 * a5/a6  OIBVH — source/bvh.cpp:71-193 (index helpers), :437-636 (build), :638-783 (traversal)
 * ============================================================================================== */

static int np2(int x)
{
    x--;
    x |= x >> 1;
    x |= x >> 2;
    x |= x >> 4;
    x |= x >> 8;
    x |= x >> 16;
    x++;
    return x;
}
static int ilog2u(unsigned x) { return 31 - __builtin_clz(x); }

int mco_oibvh_size(int t) { return 2 * t - 1 + __builtin_popcount((unsigned)(np2(t) - t)); } /* bvh.cpp:131-134 */

static int level_leftmost(int level) { return (1 << level) - 1; }
/* bvh.cpp:163-172: ancestor of the right-most real leaf on `level`; the reference evaluates this in float */
static int level_rightmost_real(int rightmost_leaf, int leaf_level, int level)
{
    const int dist = leaf_level - level;
    return (int)((1.0f / (1 << dist)) + ((float)rightmost_leaf / (1 << dist)) - 1);
}

typedef struct {
    int nleaves, leaf_level, rightmost_leaf;
    double* box; /* [(2^(leaf_level+1)-1) * 6], indexed by implicit index (layout is not observable) */
    uint32_t* leaf_face; /* [nleaves] */
} oibvh_t;

static int overlap6(const double* a, const double* b)
{
    /* math.h:931-941: closed intervals */
    return (a[0] <= b[3] && a[3] >= b[0]) && (a[1] <= b[4] && a[4] >= b[1]) && (a[2] <= b[5] && a[5] >= b[2]);
}

typedef struct {
    uint32_t code, face;
} leafkey_t;
static int cmp_leafkey(const void* a, const void* b)
{
    const leafkey_t* x = (const leafkey_t*)a;
    const leafkey_t* y = (const leafkey_t*)b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return (x->face > y->face) - (x->face < y->face);
}

static void oibvh_build(oibvh_t* t, const double* bboxes, uint32_t nf)
{
    double root[6];
    for (int j = 0; j < 3; ++j) {
        root[j] = DBL_MAX;
        root[3 + j] = -DBL_MAX;
    }
    for (uint32_t f = 0; f < nf; ++f)
        for (int j = 0; j < 3; ++j) {
            const double* b = bboxes + 6 * (size_t)f;
            root[3 + j] = (root[3 + j] < b[3 + j]) ? b[3 + j] : root[3 + j];
            root[j] = (b[j] < root[j]) ? b[j] : root[j];
        }
    uint32_t* codes = (uint32_t*)malloc(sizeof(uint32_t) * nf);
    mco_morton_codes(bboxes, nf, root, codes);
    leafkey_t* keys = (leafkey_t*)malloc(sizeof(leafkey_t) * nf);
    for (uint32_t f = 0; f < nf; ++f) {
        keys[f].code = codes[f];
        keys[f].face = f;
    }
    /* bvh.cpp:437-442 uses an unstable std::sort by code; ties cannot change the emitted pair set */
    qsort(keys, nf, sizeof(leafkey_t), cmp_leafkey);
    t->nleaves = (int)nf;
    t->leaf_level = ilog2u((unsigned)np2((int)nf));
    t->rightmost_leaf = level_leftmost(t->leaf_level) + (int)nf - 1;
    const size_t nodes = ((size_t)1 << (t->leaf_level + 1)) - 1;
    t->box = (double*)malloc(sizeof(double) * 6 * nodes);
    t->leaf_face = (uint32_t*)malloc(sizeof(uint32_t) * nf);
    const int ll = level_leftmost(t->leaf_level);
    for (uint32_t i = 0; i < nf; ++i) {
        t->leaf_face[i] = keys[i].face;
        memcpy(t->box + 6 * (size_t)(ll + (int)i), bboxes + 6 * (size_t)keys[i].face, sizeof(double) * 6);
    }
    for (int level = t->leaf_level - 1; level >= 0; --level) { /* bvh.cpp:498-635 */
        const int right = level_rightmost_real(t->rightmost_leaf, t->leaf_level, level);
        const int left = level_leftmost(level);
        const int child_right = level_rightmost_real(t->rightmost_leaf, t->leaf_level, level + 1);
        for (int n = left; n <= right; ++n) {
            double* b = t->box + 6 * (size_t)n;
            const double* l = t->box + 6 * (size_t)(2 * n + 1);
            for (int j = 0; j < 3; ++j) {
                b[j] = DBL_MAX;
                b[3 + j] = -DBL_MAX;
            }
            for (int j = 0; j < 3; ++j) {
                b[3 + j] = (b[3 + j] < l[3 + j]) ? l[3 + j] : b[3 + j];
                b[j] = (l[j] < b[j]) ? l[j] : b[j];
            }
            if (2 * n + 2 <= child_right) {
                const double* r = t->box + 6 * (size_t)(2 * n + 2);
                for (int j = 0; j < 3; ++j) {
                    b[3 + j] = (b[3 + j] < r[3 + j]) ? r[3 + j] : b[3 + j];
                    b[j] = (r[j] < b[j]) ? r[j] : b[j];
                }
            }
        }
    }
    free(keys);
    free(codes);
}

static void oibvh_free(oibvh_t* t)
{
    free(t->box);
    free(t->leaf_face);
}

static int cmp_u64(const void* a, const void* b)
{
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

size_t mco_oibvh_pairs(const double* src_bboxes, uint32_t nsf, const double* cut_bboxes, uint32_t ncf,
    uint64_t** pairs_out, uint64_t* n_tests_out)
{
    oibvh_t S, C;
    oibvh_build(&S, src_bboxes, nsf);
    oibvh_build(&C, cut_bboxes, ncf);
    /* BFS queue of (src node, cut node) implicit indices — bvh.cpp:647-781 */
    size_t qcap = 1024, qhead = 0, qtail = 0;
    uint64_t* q = (uint64_t*)malloc(sizeof(uint64_t) * qcap);
    size_t pcap = 1024, np = 0;
    uint64_t* pairs = (uint64_t*)malloc(sizeof(uint64_t) * pcap);
    uint64_t n_tests = 0;
#define QPUSH(a, b)                                                      \
    do {                                                                 \
        if (qtail == qcap) {                                             \
            if (qhead > qcap / 2) {                                      \
                memmove(q, q + qhead, sizeof(uint64_t) * (qtail - qhead)); \
                qtail -= qhead;                                          \
                qhead = 0;                                               \
            } else {                                                     \
                qcap *= 2;                                               \
                q = (uint64_t*)realloc(q, sizeof(uint64_t) * qcap);      \
            }                                                            \
        }                                                                \
        q[qtail++] = ((uint64_t)(uint32_t)(a) << 32) | (uint32_t)(b);    \
    } while (0)
    QPUSH(0, 0);
    while (qhead < qtail) {
        const uint64_t front = q[qhead++];
        const int a = (int)(front >> 32), b = (int)(front & 0xFFFFFFFFu);
        const int la = ilog2u((unsigned)a + 1u), lb = ilog2u((unsigned)b + 1u);
        const int a_leaf = (la == S.leaf_level), b_leaf = (lb == C.leaf_level);
        ++n_tests;
        if (!overlap6(S.box + 6 * (size_t)a, C.box + 6 * (size_t)b)) continue;
        if (a_leaf && b_leaf) {
            const uint32_t sf = S.leaf_face[a - level_leftmost(la)];
            const uint32_t cf = C.leaf_face[b - level_leftmost(lb)];
            if (np == pcap) {
                pcap *= 2;
                pairs = (uint64_t*)realloc(pairs, sizeof(uint64_t) * pcap);
            }
            pairs[np++] = ((uint64_t)sf << 32) | cf;
        } else if (a_leaf) {
            const int cr = level_rightmost_real(C.rightmost_leaf, C.leaf_level, lb + 1);
            QPUSH(a, 2 * b + 1);
            if (2 * b + 2 <= cr) QPUSH(a, 2 * b + 2);
        } else if (b_leaf) {
            const int sr = level_rightmost_real(S.rightmost_leaf, S.leaf_level, la + 1);
            QPUSH(2 * a + 1, b);
            if (2 * a + 2 <= sr) QPUSH(2 * a + 2, b);
        } else {
            const int sr = level_rightmost_real(S.rightmost_leaf, S.leaf_level, la + 1);
            const int cr = level_rightmost_real(C.rightmost_leaf, C.leaf_level, lb + 1);
            const int s_right = (2 * a + 2 <= sr), c_right = (2 * b + 2 <= cr);
            QPUSH(2 * a + 1, 2 * b + 1);
            if (c_right) QPUSH(2 * a + 1, 2 * b + 2);
            if (s_right) {
                QPUSH(2 * a + 2, 2 * b + 1);
                if (c_right) QPUSH(2 * a + 2, 2 * b + 2);
            }
        }
    }
#undef QPUSH
    free(q);
    oibvh_free(&S);
    oibvh_free(&C);
    qsort(pairs, np, sizeof(uint64_t), cmp_u64);
    *pairs_out = pairs;
    if (n_tests_out) *n_tests_out = n_tests;
    return np;
}

size_t mco_grid_pairs(const double* src_bboxes, uint32_t nsf, const double* cut_bboxes, uint32_t ncf, uint64_t** pairs_out)
{
    /* uniform grid over the cut boxes; each source box visits the cells it touches */
    double lo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, hi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX }, mean[3] = { 0, 0, 0 };
    for (uint32_t c = 0; c < ncf; ++c)
        for (int j = 0; j < 3; ++j) {
            const double* b = cut_bboxes + 6 * (size_t)c;
            if (b[j] < lo[j]) lo[j] = b[j];
            if (b[3 + j] > hi[j]) hi[j] = b[3 + j];
            mean[j] += (b[3 + j] - b[j]) / (double)ncf;
        }
    int dim[3];
    double cell[3];
    for (int j = 0; j < 3; ++j) {
        double ext = hi[j] - lo[j];
        double cs = mean[j] * 2.0;
        int d = (cs > 0.0 && ext > 0.0) ? (int)(ext / cs) : 1;
        if (d < 1) d = 1;
        if (d > 128) d = 128;
        dim[j] = d;
        cell[j] = ext > 0.0 ? ext / d : 1.0;
    }
#define CELL_OF(x, j) ((int)(((x) - lo[j]) / cell[j]) < 0 ? 0 : ((int)(((x) - lo[j]) / cell[j]) >= dim[j] ? dim[j] - 1 : (int)(((x) - lo[j]) / cell[j])))
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    uint32_t* start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    for (int pass = 0; pass < 2; ++pass) {
        static uint32_t* items;
        if (pass == 1) {
            uint32_t acc = 0;
            for (size_t i = 0; i <= ncell; ++i) {
                uint32_t t = start[i];
                start[i] = acc;
                acc += t;
            }
            items = (uint32_t*)malloc(sizeof(uint32_t) * (acc ? acc : 1));
        }
        for (uint32_t c = 0; c < ncf; ++c) {
            const double* b = cut_bboxes + 6 * (size_t)c;
            int c0[3], c1[3];
            for (int j = 0; j < 3; ++j) {
                c0[j] = CELL_OF(b[j], j);
                c1[j] = CELL_OF(b[3 + j], j);
            }
            for (int x = c0[0]; x <= c1[0]; ++x)
                for (int y = c0[1]; y <= c1[1]; ++y)
                    for (int z = c0[2]; z <= c1[2]; ++z) {
                        const size_t id = ((size_t)x * dim[1] + y) * dim[2] + z;
                        if (pass == 0) start[id]++;
                        else items[start[id]++] = c;
                    }
        }
        if (pass == 1) {
            /* start[] was advanced; shift back */
            for (size_t i = ncell; i > 0; --i) start[i] = start[i - 1];
            start[0] = 0;
            size_t pcap = 1024, np = 0;
            uint64_t* pairs = (uint64_t*)malloc(sizeof(uint64_t) * pcap);
            for (uint32_t s = 0; s < nsf; ++s) {
                const double* a = src_bboxes + 6 * (size_t)s;
                int outside = 0;
                for (int j = 0; j < 3; ++j)
                    if (a[3 + j] < lo[j] || a[j] > hi[j]) outside = 1;
                if (outside) continue;
                int c0[3], c1[3];
                for (int j = 0; j < 3; ++j) {
                    c0[j] = CELL_OF(a[j], j);
                    c1[j] = CELL_OF(a[3 + j], j);
                }
                const size_t first = np;
                for (int x = c0[0]; x <= c1[0]; ++x)
                    for (int y = c0[1]; y <= c1[1]; ++y)
                        for (int z = c0[2]; z <= c1[2]; ++z) {
                            const size_t id = ((size_t)x * dim[1] + y) * dim[2] + z;
                            for (uint32_t k = start[id]; k < start[id + 1]; ++k) {
                                const uint32_t c = items[k];
                                if (!overlap6(a, cut_bboxes + 6 * (size_t)c)) continue;
                                if (np == pcap) {
                                    pcap *= 2;
                                    pairs = (uint64_t*)realloc(pairs, sizeof(uint64_t) * pcap);
                                }
                                pairs[np++] = ((uint64_t)s << 32) | c;
                            }
                        }
                /* de-duplicate this source face's hits (a cut box can live in several visited cells) */
                if (np - first > 1) {
                    qsort(pairs + first, np - first, sizeof(uint64_t), cmp_u64);
                    size_t w = first + 1;
                    for (size_t r = first + 1; r < np; ++r)
                        if (pairs[r] != pairs[w - 1]) pairs[w++] = pairs[r];
                    np = w;
                }
            }
            free(items);
            free(start);
            *pairs_out = pairs;
            return np;
        }
    }
#undef CELL_OF
    return 0; /* unreachable */
}

void mco_free(void* p) { free(p); }

/* ================================================================================================
 * polygon soup ids — source/kernel.cpp:1593-1732, source/hmesh.cpp:406-651 (add_edge/add_face),
 * :705-733 (get_vertices_around_face returns halfedge TARGETS => the user's list rotated by one)
 * ============================================================================================== */

typedef struct {
    uint64_t key; /* (min << 32 | max) + 1, 0 = empty */
    uint32_t edge;
} eslot_t;

static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

int mco_soup_build(const double* src_xyz, uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf,
    const double* cut_xyz, uint32_t ncv, const uint32_t* cut_off, const uint32_t* cut_vtx, uint32_t ncf, mco_soup_t* out)
{
    memset(out, 0, sizeof(*out));
    const uint32_t nhs = src_off[nsf], nhc = cut_off[ncf];
    out->nv = nsv + ncv;
    out->nf = nsf + ncf;
    out->nh = nhs + nhc;
    out->src_nv = nsv;
    out->src_nf = nsf;
    out->xyz = (double*)malloc(sizeof(double) * 3 * (size_t)out->nv);
    memcpy(out->xyz, src_xyz, sizeof(double) * 3 * (size_t)nsv);
    memcpy(out->xyz + 3 * (size_t)nsv, cut_xyz, sizeof(double) * 3 * (size_t)ncv);
    out->face_off = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)out->nf + 1));
    out->face_vtx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)out->nh);
    out->face_edge = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)out->nh);
    out->edge_v = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (size_t)out->nh);
    out->edge_f = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (size_t)out->nh);
    size_t cap = 16;
    while (cap < 2 * (size_t)out->nh) cap <<= 1;
    eslot_t* tab = (eslot_t*)calloc(cap, sizeof(eslot_t));
    uint32_t ne = 0, h = 0;
    int rc = 0;
    for (uint32_t f = 0; f < out->nf && rc == 0; ++f) {
        const int is_cut = f >= nsf;
        const uint32_t* vtx = is_cut ? cut_vtx + cut_off[f - nsf] : src_vtx + src_off[f];
        const uint32_t n = is_cut ? cut_off[f - nsf + 1] - cut_off[f - nsf] : src_off[f + 1] - src_off[f];
        const uint32_t voff = is_cut ? nsv : 0;
        /* list handed to add_face: source faces in the user's order (ps starts as a copy of the source
         * hmesh); cut faces as cs.get_vertices_around_face() returns them, i.e. already rotated by one
         * (kernel.cpp:1678, 1699, 1720-1723) */
        const uint32_t rot_in = is_cut ? 1u : 0u;
        out->face_off[f] = h;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t v0 = vtx[(i + rot_in) % n] + voff;
            const uint32_t v1 = vtx[(i + rot_in + 1) % n] + voff;
            const uint32_t lo = v0 < v1 ? v0 : v1, hi = v0 < v1 ? v1 : v0;
            const uint64_t key = (((uint64_t)lo << 32) | hi) + 1;
            size_t s = (size_t)mix64(key) & (cap - 1);
            while (tab[s].key != 0 && tab[s].key != key) s = (s + 1) & (cap - 1);
            uint32_t e;
            if (tab[s].key == 0) { /* hmesh.cpp:406-521: new edge, h0 runs v0 -> v1 */
                e = ne++;
                tab[s].key = key;
                tab[s].edge = e;
                out->edge_v[2 * (size_t)e] = v0;
                out->edge_v[2 * (size_t)e + 1] = v1;
                out->edge_f[2 * (size_t)e] = f;
                out->edge_f[2 * (size_t)e + 1] = MCO_NULL;
            } else {
                e = tab[s].edge;
                if (out->edge_v[2 * (size_t)e] == v0 || out->edge_f[2 * (size_t)e + 1] != MCO_NULL) {
                    rc = -1; /* hmesh.cpp:612-628: halfedge already owned by a face */
                    break;
                }
                out->edge_f[2 * (size_t)e + 1] = f;
            }
            out->face_vtx[h] = v1; /* target of halfedge i (hmesh.cpp:705-733) */
            out->face_edge[h] = e;
            ++h;
        }
    }
    out->face_off[out->nf] = h;
    out->ne = ne;
    free(tab);
    if (rc != 0) mco_soup_free(out);
    return rc;
}

void mco_soup_free(mco_soup_t* s)
{
    free(s->xyz);
    free(s->face_off);
    free(s->face_vtx);
    free(s->face_edge);
    free(s->edge_v);
    free(s->edge_f);
    memset(s, 0, sizeof(*s));
}

/* ================================================================================================
 * a9  plane of a face — source/math.cpp:130-239
 * ============================================================================================== */

static double dot3(const double a[3], const double b[3])
{
    /* math.h:634-642: accumulates from 0.0, x then y then z */
    double out = 0.0;
    for (int i = 0; i < 3; ++i) out += a[i] * b[i];
    return out;
}

int mco_plane_coefficients(const double* v, int n, double normal[3], double* d)
{
    normal[0] = normal[1] = normal[2] = 0.0;
    *d = 0.0;
    for (int i = 0; i < n; ++i) { /* Newell, math.cpp:140-149 */
        const double* c = v + 3 * (size_t)i;
        const double* x = v + 3 * (size_t)((i + 1) % n);
        normal[0] += (c[1] - x[1]) * (c[2] + x[2]);
        normal[1] += (c[2] - x[2]) * (c[0] + x[0]);
        normal[2] += (c[0] - x[0]) * (c[1] + x[1]);
    }
    if (isnan(normal[0]) || isnan(normal[1]) || isnan(normal[2]) || dot3(normal, normal) < 1e-9) {
        normal[0] = normal[1] = normal[2] = 0.0; /* :151-174 */
        return 0;
    }
    if (dot3(normal, normal) == 0.0) return 0; /* :207-210 */
    const double len = sqrt(dot3(normal, normal)); /* normalize: v / length(v), math.h:717-721 */
    for (int i = 0; i < 3; ++i) normal[i] = normal[i] / len;
    *d = dot3(v, normal); /* :224 */
    double largest = 0.0;
    int idx = 0;
    for (int i = 0; i < 3; ++i) { /* :226-236 */
        const double t = fabs(normal[i]);
        if (t > largest) {
            largest = t;
            idx = i;
        }
    }
    return idx;
}

/* ================================================================================================
 * a10/a11  Shewchuk's adaptive orientation predicates — source/shewchuk.c
 * constants :420-433; Two_Sum/Two_Diff/Split/Two_Product :175-335; expansion ops :1096,:1368,:1468;
 * orient2d :1695-1729; orient2dadapt :1611-1693; orient3d :2367-2410; orient3dadapt :1962-2365
 * ============================================================================================== */

static const double k_splitter = 134217729.0; /* 2^27 + 1 */
static const double k_resulterrbound = 3.3306690738754706e-16;
static const double k_ccwerrboundA = 3.3306690738754716e-16;
static const double k_ccwerrboundB = 2.2204460492503146e-16;
static const double k_ccwerrboundC = 1.1093356479670487e-31;
static const double k_o3derrboundA = 7.7715611723761027e-16;
static const double k_o3derrboundB = 3.3306690738754731e-16;
static const double k_o3derrboundC = 3.2047474274603644e-31;

typedef struct {
    double hi, lo;
} dd_t;

static inline dd_t fast_two_sum(double a, double b)
{
    dd_t r;
    r.hi = a + b;
    const double bv = r.hi - a;
    r.lo = b - bv;
    return r;
}
static inline dd_t two_sum(double a, double b)
{
    dd_t r;
    r.hi = a + b;
    const double bv = r.hi - a;
    const double av = r.hi - bv;
    const double br = b - bv;
    const double ar = a - av;
    r.lo = ar + br;
    return r;
}
static inline double two_diff_tail(double a, double b, double x)
{
    const double bv = a - x;
    const double av = x + bv;
    const double br = bv - b;
    const double ar = a - av;
    return ar + br;
}
static inline dd_t two_diff(double a, double b)
{
    dd_t r;
    r.hi = a - b;
    r.lo = two_diff_tail(a, b, r.hi);
    return r;
}
static inline dd_t split(double a)
{
    dd_t r;
    const double c = k_splitter * a;
    const double abig = c - a;
    r.hi = c - abig;
    r.lo = a - r.hi;
    return r;
}
static inline dd_t two_product_presplit(double a, double b, dd_t bs)
{
    dd_t r;
    r.hi = a * b;
    const dd_t as = split(a);
    const double e1 = r.hi - (as.hi * bs.hi);
    const double e2 = e1 - (as.lo * bs.hi);
    const double e3 = e2 - (as.hi * bs.lo);
    r.lo = (as.lo * bs.lo) - e3;
    return r;
}
static inline dd_t two_product(double a, double b) { return two_product_presplit(a, b, split(b)); }

/* (a1 + a0) - b  ->  x[2] + x[1] + x[0] */
static inline void two_one_diff(double a1, double a0, double b, double* x2, double* x1, double* x0)
{
    const dd_t i = two_diff(a0, b);
    *x0 = i.lo;
    const dd_t j = two_sum(a1, i.hi);
    *x2 = j.hi;
    *x1 = j.lo;
}
/* (a1 + a0) - (b1 + b0)  ->  x[3..0] */
static inline void two_two_diff(dd_t a, dd_t b, double x[4])
{
    double j, z;
    two_one_diff(a.hi, a.lo, b.lo, &j, &z, &x[0]);
    two_one_diff(j, z, b.hi, &x[3], &x[2], &x[1]);
}
/* (a1 + a0) * b  ->  x[3..0] */
static inline void two_one_product(dd_t a, double b, double x[4])
{
    const dd_t bs = split(b);
    const dd_t i = two_product_presplit(a.lo, b, bs);
    x[0] = i.lo;
    const dd_t j = two_product_presplit(a.hi, b, bs);
    const dd_t k = two_sum(i.hi, j.lo);
    x[1] = k.lo;
    const dd_t l = fast_two_sum(j.hi, k.hi);
    x[3] = l.hi;
    x[2] = l.lo;
}

static int expansion_sum(int elen, const double* e, int flen, const double* f, double* h)
{
    /* fast_expansion_sum_zeroelim: merge by magnitude, accumulate with (fast-)two-sum, drop zeros */
    double Q;
    int ei = 0, fi = 0, hi = 0;
    double enow = e[0], fnow = f[0];
    if ((fnow > enow) == (fnow > -enow)) {
        Q = enow;
        ++ei;
        if (ei < elen) enow = e[ei];
    } else {
        Q = fnow;
        ++fi;
        if (fi < flen) fnow = f[fi];
    }
    if (ei < elen && fi < flen) {
        dd_t s;
        if ((fnow > enow) == (fnow > -enow)) {
            s = fast_two_sum(enow, Q);
            ++ei;
            if (ei < elen) enow = e[ei];
        } else {
            s = fast_two_sum(fnow, Q);
            ++fi;
            if (fi < flen) fnow = f[fi];
        }
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
        while (ei < elen && fi < flen) {
            if ((fnow > enow) == (fnow > -enow)) {
                s = two_sum(Q, enow);
                ++ei;
                if (ei < elen) enow = e[ei];
            } else {
                s = two_sum(Q, fnow);
                ++fi;
                if (fi < flen) fnow = f[fi];
            }
            Q = s.hi;
            if (s.lo != 0.0) h[hi++] = s.lo;
        }
    }
    while (ei < elen) {
        const dd_t s = two_sum(Q, enow);
        ++ei;
        if (ei < elen) enow = e[ei];
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
    }
    while (fi < flen) {
        const dd_t s = two_sum(Q, fnow);
        ++fi;
        if (fi < flen) fnow = f[fi];
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
    }
    if (Q != 0.0 || hi == 0) h[hi++] = Q;
    return hi;
}

static int expansion_scale(int elen, const double* e, double b, double* h)
{
    /* scale_expansion_zeroelim */
    const dd_t bs = split(b);
    dd_t p = two_product_presplit(e[0], b, bs);
    double Q = p.hi;
    int hi = 0;
    if (p.lo != 0.0) h[hi++] = p.lo;
    for (int i = 1; i < elen; ++i) {
        p = two_product_presplit(e[i], b, bs);
        const dd_t s = two_sum(Q, p.lo);
        if (s.lo != 0.0) h[hi++] = s.lo;
        const dd_t t = fast_two_sum(p.hi, s.hi);
        Q = t.hi;
        if (t.lo != 0.0) h[hi++] = t.lo;
    }
    if (Q != 0.0 || hi == 0) h[hi++] = Q;
    return hi;
}

static double expansion_estimate(int n, const double* e)
{
    double q = e[0];
    for (int i = 1; i < n; ++i) q += e[i];
    return q;
}

static double orient2d_adapt(const double* pa, const double* pb, const double* pc, double detsum)
{
    const double acx = pa[0] - pc[0], bcx = pb[0] - pc[0];
    const double acy = pa[1] - pc[1], bcy = pb[1] - pc[1];
    const dd_t left = two_product(acx, bcy);
    const dd_t right = two_product(acy, bcx);
    double B[4];
    two_two_diff(left, right, B);
    double det = expansion_estimate(4, B);
    double errbound = k_ccwerrboundB * detsum;
    if (det >= errbound || -det >= errbound) return det;

    const double acxt = two_diff_tail(pa[0], pc[0], acx);
    const double bcxt = two_diff_tail(pb[0], pc[0], bcx);
    const double acyt = two_diff_tail(pa[1], pc[1], acy);
    const double bcyt = two_diff_tail(pb[1], pc[1], bcy);
    if (acxt == 0.0 && acyt == 0.0 && bcxt == 0.0 && bcyt == 0.0) return det;

    errbound = k_ccwerrboundC * detsum + k_resulterrbound * fabs(det);
    det += (acx * bcyt + bcy * acxt) - (acy * bcxt + bcx * acyt);
    if (det >= errbound || -det >= errbound) return det;

    double u[4], C1[8], C2[12], D[16];
    two_two_diff(two_product(acxt, bcy), two_product(acyt, bcx), u);
    const int c1 = expansion_sum(4, B, 4, u, C1);
    two_two_diff(two_product(acx, bcyt), two_product(acy, bcxt), u);
    const int c2 = expansion_sum(c1, C1, 4, u, C2);
    two_two_diff(two_product(acxt, bcyt), two_product(acyt, bcxt), u);
    const int dl = expansion_sum(c2, C2, 4, u, D);
    return D[dl - 1];
}

double mco_orient2d(const double pa[2], const double pb[2], const double pc[2])
{
    const double detleft = (pa[0] - pc[0]) * (pb[1] - pc[1]);
    const double detright = (pa[1] - pc[1]) * (pb[0] - pc[0]);
    const double det = detleft - detright;
    double detsum;
    if (detleft > 0.0) {
        if (detright <= 0.0) return det;
        detsum = detleft + detright;
    } else if (detleft < 0.0) {
        if (detright >= 0.0) return det;
        detsum = -detleft - detright;
    } else {
        return det;
    }
    const double errbound = k_ccwerrboundA * detsum;
    if (det >= errbound || -det >= errbound) return det;
    return orient2d_adapt(pa, pb, pc, detsum);
}

/* one of the three "tail x other" groups of stage D: given the tails (xt, yt) of one point and the
 * rounded (x, y) differences of the other two points m and n, produce the expansions
 *   t_m = xt*my - yt*mx   and   t_n = yt*nx - xt*ny   (lengths 1, 2 or 4) */
static void tail_cross(double xt, double yt, double mx, double my, double nx, double ny, double* tm, int* tmlen, double* tn,
    int* tnlen)
{
    if (xt == 0.0) {
        if (yt == 0.0) {
            tm[0] = 0.0;
            *tmlen = 1;
            tn[0] = 0.0;
            *tnlen = 1;
        } else {
            dd_t p = two_product(-yt, mx);
            tm[0] = p.lo;
            tm[1] = p.hi;
            *tmlen = 2;
            p = two_product(yt, nx);
            tn[0] = p.lo;
            tn[1] = p.hi;
            *tnlen = 2;
        }
    } else if (yt == 0.0) {
        dd_t p = two_product(xt, my);
        tm[0] = p.lo;
        tm[1] = p.hi;
        *tmlen = 2;
        p = two_product(-xt, ny);
        tn[0] = p.lo;
        tn[1] = p.hi;
        *tnlen = 2;
    } else {
        two_two_diff(two_product(xt, my), two_product(yt, mx), tm);
        *tmlen = 4;
        two_two_diff(two_product(yt, nx), two_product(xt, ny), tn);
        *tnlen = 4;
    }
}

typedef struct {
    double buf[2][192];
    int cur, len;
} fin_t;

static inline void fin_add(fin_t* f, int n, const double* e)
{
    f->len = expansion_sum(f->len, f->buf[f->cur], n, e, f->buf[f->cur ^ 1]);
    f->cur ^= 1;
}

static double orient3d_adapt(const double* pa, const double* pb, const double* pc, const double* pd, double permanent)
{
    const double adx = pa[0] - pd[0], bdx = pb[0] - pd[0], cdx = pc[0] - pd[0];
    const double ady = pa[1] - pd[1], bdy = pb[1] - pd[1], cdy = pc[1] - pd[1];
    const double adz = pa[2] - pd[2], bdz = pb[2] - pd[2], cdz = pc[2] - pd[2];

    double bc[4], ca[4], ab[4], adet[8], bdet[8], cdet[8], abdet[16];
    two_two_diff(two_product(bdx, cdy), two_product(cdx, bdy), bc);
    const int alen = expansion_scale(4, bc, adz, adet);
    two_two_diff(two_product(cdx, ady), two_product(adx, cdy), ca);
    const int blen = expansion_scale(4, ca, bdz, bdet);
    two_two_diff(two_product(adx, bdy), two_product(bdx, ady), ab);
    const int clen = expansion_scale(4, ab, cdz, cdet);

    fin_t fin;
    fin.cur = 0;
    const int ablen = expansion_sum(alen, adet, blen, bdet, abdet);
    fin.len = expansion_sum(ablen, abdet, clen, cdet, fin.buf[0]);

    double det = expansion_estimate(fin.len, fin.buf[0]);
    double errbound = k_o3derrboundB * permanent;
    if (det >= errbound || -det >= errbound) return det;

    const double adxt = two_diff_tail(pa[0], pd[0], adx), bdxt = two_diff_tail(pb[0], pd[0], bdx),
                 cdxt = two_diff_tail(pc[0], pd[0], cdx);
    const double adyt = two_diff_tail(pa[1], pd[1], ady), bdyt = two_diff_tail(pb[1], pd[1], bdy),
                 cdyt = two_diff_tail(pc[1], pd[1], cdy);
    const double adzt = two_diff_tail(pa[2], pd[2], adz), bdzt = two_diff_tail(pb[2], pd[2], bdz),
                 cdzt = two_diff_tail(pc[2], pd[2], cdz);
    if (adxt == 0.0 && bdxt == 0.0 && cdxt == 0.0 && adyt == 0.0 && bdyt == 0.0 && cdyt == 0.0 && adzt == 0.0
        && bdzt == 0.0 && cdzt == 0.0)
        return det;

    errbound = k_o3derrboundC * permanent + k_resulterrbound * fabs(det);
    det += (adz * ((bdx * cdyt + cdy * bdxt) - (bdy * cdxt + cdx * bdyt)) + adzt * (bdx * cdy - bdy * cdx))
        + (bdz * ((cdx * adyt + ady * cdxt) - (cdy * adxt + adx * cdyt)) + bdzt * (cdx * ady - cdy * adx))
        + (cdz * ((adx * bdyt + bdy * adxt) - (ady * bdxt + bdx * adyt)) + cdzt * (adx * bdy - ady * bdx));
    if (det >= errbound || -det >= errbound) return det;

    /* stage D: exact.  at_b = adxt*bdy - adyt*bdx, at_c = adyt*cdx - adxt*cdy, and cyclically */
    double at_b[4], at_c[4], bt_c[4], bt_a[4], ct_a[4], ct_b[4];
    int at_bl, at_cl, bt_cl, bt_al, ct_al, ct_bl;
    tail_cross(adxt, adyt, bdx, bdy, cdx, cdy, at_b, &at_bl, at_c, &at_cl);
    tail_cross(bdxt, bdyt, cdx, cdy, adx, ady, bt_c, &bt_cl, bt_a, &bt_al);
    tail_cross(cdxt, cdyt, adx, ady, bdx, bdy, ct_a, &ct_al, ct_b, &ct_bl);

    double bct[8], cat[8], abt[8], w[16], v[12], u[4];
    const int bctl = expansion_sum(bt_cl, bt_c, ct_bl, ct_b, bct);
    fin_add(&fin, expansion_scale(bctl, bct, adz, w), w);
    const int catl = expansion_sum(ct_al, ct_a, at_cl, at_c, cat);
    fin_add(&fin, expansion_scale(catl, cat, bdz, w), w);
    const int abtl = expansion_sum(at_bl, at_b, bt_al, bt_a, abt);
    fin_add(&fin, expansion_scale(abtl, abt, cdz, w), w);

    if (adzt != 0.0) fin_add(&fin, expansion_scale(4, bc, adzt, v), v);
    if (bdzt != 0.0) fin_add(&fin, expansion_scale(4, ca, bdzt, v), v);
    if (cdzt != 0.0) fin_add(&fin, expansion_scale(4, ab, cdzt, v), v);

    /* tail x tail terms: (xt_i * yt_j) * z_k and, when z_k has a tail, * zt_k */
#define TT(xt, yt, z, zt)                          \
    do {                                           \
        const dd_t p_ = two_product((xt), (yt));   \
        two_one_product(p_, (z), u);               \
        fin_add(&fin, 4, u);                       \
        if ((zt) != 0.0) {                         \
            two_one_product(p_, (zt), u);          \
            fin_add(&fin, 4, u);                   \
        }                                          \
    } while (0)
    if (adxt != 0.0) {
        if (bdyt != 0.0) TT(adxt, bdyt, cdz, cdzt);
        if (cdyt != 0.0) TT(-adxt, cdyt, bdz, bdzt);
    }
    if (bdxt != 0.0) {
        if (cdyt != 0.0) TT(bdxt, cdyt, adz, adzt);
        if (adyt != 0.0) TT(-bdxt, adyt, cdz, cdzt);
    }
    if (cdxt != 0.0) {
        if (adyt != 0.0) TT(cdxt, adyt, bdz, bdzt);
        if (bdyt != 0.0) TT(-cdxt, bdyt, adz, adzt);
    }
#undef TT
    if (adzt != 0.0) fin_add(&fin, expansion_scale(bctl, bct, adzt, w), w);
    if (bdzt != 0.0) fin_add(&fin, expansion_scale(catl, cat, bdzt, w), w);
    if (cdzt != 0.0) fin_add(&fin, expansion_scale(abtl, abt, cdzt, w), w);

    return fin.buf[fin.cur][fin.len - 1];
}

double mco_orient3d_stageA(const double pa[3], const double pb[3], const double pc[3], const double pd[3], int* certain)
{
    const double adx = pa[0] - pd[0], bdx = pb[0] - pd[0], cdx = pc[0] - pd[0];
    const double ady = pa[1] - pd[1], bdy = pb[1] - pd[1], cdy = pc[1] - pd[1];
    const double adz = pa[2] - pd[2], bdz = pb[2] - pd[2], cdz = pc[2] - pd[2];
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    const double cdxady = cdx * ady, adxcdy = adx * cdy;
    const double adxbdy = adx * bdy, bdxady = bdx * ady;
    const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz)
        + (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    const double errbound = k_o3derrboundA * permanent;
    *certain = (det > errbound) || (-det > errbound);
    return *certain ? det : permanent;
}

double mco_orient3d(const double pa[3], const double pb[3], const double pc[3], const double pd[3])
{
    int certain;
    const double r = mco_orient3d_stageA(pa, pb, pc, pd, &certain);
    if (certain) return r;
    return orient3d_adapt(pa, pb, pc, pd, r /* permanent */);
}

/* ================================================================================================
 * a14  projection to 2D + point in polygon — source/math.cpp:710-793, :851-902, :553-704
 * ============================================================================================== */

void mco_projection_matrix(const double normal[3], int max_comp, double P[6])
{
    const double len = sqrt(dot3(normal, normal));
    const double a[3] = { normal[0] / len, normal[1] / len, normal[2] / len }; /* :718 */
    double b[3] = { 0.0, 0.0, 0.0 };
    const int s = (0.0 < normal[max_comp]) - (normal[max_comp] < 0.0); /* math.cpp:66-77 */
    b[max_comp] = 1.0 * (double)s; /* :726 */
    double I[3][3] = { { 1.0, 0.0, 0.0 }, { 0.0, -1.0, 0.0 }, { 0.0, 0.0, 1.0 } }; /* :730-733 */
    double R[3][3];
    memcpy(R, I, sizeof(R));
    if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) { /* :740-754 */
        const double apb[3] = { a[0] + b[0], a[1] + b[1], a[2] + b[2] };
        const double adb = dot3(a, b);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double outer = apb[i] * apb[j]; /* math.h:646-675: column j = a * b[j] */
                R[i][j] = ((outer / adb) * 2.0) - I[i][j];
            }
    }
    double K[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }; /* :758-790 */
    if (max_comp == 0) {
        K[0][1] = 1.0;
        K[1][2] = 1.0;
    } else if (max_comp == 1) {
        K[0][0] = 1.0;
        K[1][2] = 1.0;
    } else {
        K[0][0] = 1.0;
        K[1][1] = 1.0;
    }
    for (int i = 0; i < 2; ++i) /* math.h:450-463 generic product accumulating from 0.0 */
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += K[i][k] * R[k][j];
            P[3 * i + j] = acc;
        }
}

static void project2(const double P[6], const double v[3], double out[2])
{
    /* math.h:508-533: column-major accumulation */
    out[0] = 0.0;
    out[1] = 0.0;
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 2; ++row) out[row] = out[row] + (P[3 * row + col] * v[col]);
}

static char pip2d(const double q[2], const double* poly /* [n*2] */, int n, double* scratch /* [n*2] */)
{
    for (int i = 0; i < n; ++i) { /* :648-655 */
        scratch[2 * i] = poly[2 * i] - q[0];
        scratch[2 * i + 1] = poly[2 * i + 1] - q[1];
    }
    int rcross = 0, lcross = 0;
    for (int i = 0; i < n; ++i) { /* :661-690 */
        const double xi = scratch[2 * i], yi = scratch[2 * i + 1];
        if (xi == 0.0 && yi == 0.0) return 'v';
        const int il = (i + n - 1) % n;
        const double xl = scratch[2 * il], yl = scratch[2 * il + 1];
        const int rstrad = (yi > 0.0) != (yl > 0.0);
        const int lstrad = (yi < 0.0) != (yl < 0.0);
        if (rstrad || lstrad) {
            const double x = (xi * yl - xl * yi) / (yl - yi);
            if (rstrad && x > 0.0) rcross++;
            if (lstrad && x < 0.0) lcross++;
        }
    }
    if ((rcross % 2) != (lcross % 2)) return 'e';
    return (rcross % 2) == 1 ? 'i' : 'o';
}

char mco_point_in_polygon(const double p[3], const double* verts, int n, const double normal[3], int max_comp)
{
    double P[6], pp[2];
    mco_projection_matrix(normal, max_comp, P);
    project2(P, p, pp);
    double stack_buf[64];
    double* buf = (n <= 16) ? stack_buf : (double*)malloc(sizeof(double) * 4 * (size_t)n);
    for (int i = 0; i < n; ++i) project2(P, verts + 3 * (size_t)i, buf + 2 * i);
    const char r = pip2d(pp, buf, n, buf + 2 * n);
    if (buf != stack_buf) free(buf);
    return r;
}

/* ================================================================================================
 * a10/a13  segment vs plane — source/math.cpp:391-427 (+ :289-389), :249-287
 * ============================================================================================== */

static int best_noncollinear_triple(const double* verts, int n, const double normal[3], int max_comp, int ijk[3])
{
    /* math.cpp:289-389: project, orient2d on every i<j<k, keep non-zero ones, sort descending by |value| and take
     * the front.  libstdc++ sorts <= 16 elements by straight insertion (stable), so for n <= 5 the front is the
     * FIRST maximal triple in enumeration order; the same rule is applied for larger n (exact ties for the maximum
     * among > 16 triples would depend on introsort internals). */
    double P[6];
    mco_projection_matrix(normal, max_comp, P);
    double* x = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) project2(P, verts + 3 * (size_t)i, x + 2 * i);
    double best = -1.0;
    int found = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            for (int k = j + 1; k < n; ++k) {
                const double r = mco_orient2d(x + 2 * i, x + 2 * j, x + 2 * k);
                if (r == 0.0) continue;
                if (fabs(r) > best) {
                    best = fabs(r);
                    ijk[0] = i;
                    ijk[1] = j;
                    ijk[2] = k;
                    found = 1;
                }
            }
    free(x);
    return found;
}

char mco_segment_plane_type(const double q[3], const double r[3], const double* verts, int n, const double normal[3],
    int max_comp, double* q_res, double* r_res)
{
    int ijk[3] = { 0, 1, 2 };
    if (n > 3 && !best_noncollinear_triple(verts, n, normal, max_comp, ijk)) {
        if (q_res) *q_res = 0.0;
        if (r_res) *r_res = 0.0;
        return '0';
    }
    const double* a = verts + 3 * (size_t)ijk[0];
    const double* b = verts + 3 * (size_t)ijk[1];
    const double* c = verts + 3 * (size_t)ijk[2];
    const double qr = mco_orient3d(a, b, c, q);
    const double rr = mco_orient3d(a, b, c, r);
    if (q_res) *q_res = qr;
    if (r_res) *r_res = rr;
    if (qr == 0.0 && rr == 0.0) return 'p';
    if (qr == 0.0) return 'q';
    if (rr == 0.0) return 'r';
    if ((rr < 0.0 && qr < 0.0) || (rr > 0.0 && qr > 0.0)) return '0';
    return '1';
}

char mco_segment_plane_intersection(double p[3], const double normal[3], double d, const double q[3], const double r[3])
{
    const double num = d - dot3(q, normal);
    const double rq[3] = { r[0] - q[0], r[1] - q[1], r[2] - q[2] };
    const double denom = dot3(rq, normal);
    if (denom == 0.0) return (num == 0.0) ? 'p' : '0';
    const double t = num / denom;
    for (int i = 0; i < 3; ++i) p[i] = q[i] + t * (r[i] - q[i]);
    if (0.0 < t && t < 1.0) return '1';
    if (num == 0.0) return 'q';
    if (num == denom) return 'r';
    return '0';
}

/* ================================================================================================
 * a7..a15  the narrowphase — source/kernel.cpp:1779-3231
 * ============================================================================================== */

static int sgn(double x) { return (x > 0.0) - (x < 0.0); }

void mco_narrow_free(mco_narrow_out_t* o)
{
    free(o->tests);
    free(o->records);
    free(o->cand_faces);
    free(o->cand_normal);
    free(o->cand_d);
    free(o->cand_maxcomp);
    memset(o, 0, sizeof(*o));
}

int mco_narrowphase(const mco_soup_t* ps, const uint64_t* pairs, size_t npairs, const double* src_bboxes,
    const double* cut_bboxes, int stop_on_gp, mco_narrow_out_t* out)
{
    memset(out, 0, sizeof(*out));
    const uint32_t Fs = ps->src_nf;
    /* ---- candidate faces = keys of ps_face_to_potentially_intersecting_others (bvh.cpp:713-716) ---- */
    uint8_t* is_cand = (uint8_t*)calloc(ps->nf ? ps->nf : 1, 1);
    for (size_t i = 0; i < npairs; ++i) {
        is_cand[(uint32_t)(pairs[i] >> 32)] = 1;
        is_cand[Fs + (uint32_t)(pairs[i] & 0xFFFFFFFFu)] = 1;
    }
    size_t ncand = 0;
    for (uint32_t f = 0; f < ps->nf; ++f) ncand += is_cand[f];
    out->n_cand_faces = ncand;
    out->cand_faces = (uint32_t*)malloc(sizeof(uint32_t) * (ncand ? ncand : 1));
    out->cand_normal = (double*)malloc(sizeof(double) * 3 * (ncand ? ncand : 1));
    out->cand_d = (double*)malloc(sizeof(double) * (ncand ? ncand : 1));
    out->cand_maxcomp = (int32_t*)malloc(sizeof(int32_t) * (ncand ? ncand : 1));
    uint32_t* cand_slot = (uint32_t*)malloc(sizeof(uint32_t) * (ps->nf ? ps->nf : 1));

    /* ---- per-face plane data (kernel.cpp:2184-2356) ---- */
    {
        size_t k = 0;
        int bad = 0;
        double stackv[3 * 16];
        for (uint32_t f = 0; f < ps->nf; ++f) {
            if (!is_cand[f]) continue;
            const uint32_t n = ps->face_off[f + 1] - ps->face_off[f];
            double* v = n <= 16 ? stackv : (double*)malloc(sizeof(double) * 3 * n);
            for (uint32_t i = 0; i < n; ++i) memcpy(v + 3 * i, ps->xyz + 3 * (size_t)ps->face_vtx[ps->face_off[f] + i], 24);
            out->cand_faces[k] = f;
            cand_slot[f] = (uint32_t)k;
            out->cand_maxcomp[k] = mco_plane_coefficients(v, (int)n, out->cand_normal + 3 * k, out->cand_d + k);
            const double* nn = out->cand_normal + 3 * k;
            if (!bad && (dot3(nn, nn) == 0.0 || isnan(nn[0]) || isnan(nn[1]) || isnan(nn[2]))) {
                bad = 1; /* kernel.cpp:2237-2244, :2301-2312 (note the reference's `>` when classifying the mesh) */
                out->bad_face = f;
                out->status = (f > Fs) ? MCO_INVALID_CUT_MESH : MCO_INVALID_SRC_MESH;
            }
            if (v != stackv) free(v);
            ++k;
        }
        if (bad) {
            free(is_cand);
            free(cand_slot);
            return out->status;
        }
    }

    /* ---- edge -> sorted unique face list (kernel.cpp:1781-1983) ---- */
    /* tuples (edge, other face) for every candidate face, every halfedge slot, every paired face */
    size_t ntup = 0;
    for (size_t i = 0; i < npairs; ++i) {
        const uint32_t s = (uint32_t)(pairs[i] >> 32), c = Fs + (uint32_t)(pairs[i] & 0xFFFFFFFFu);
        ntup += (ps->face_off[s + 1] - ps->face_off[s]) + (ps->face_off[c + 1] - ps->face_off[c]);
    }
    uint64_t* tup = (uint64_t*)malloc(sizeof(uint64_t) * (ntup ? ntup : 1));
    size_t t = 0;
    for (size_t i = 0; i < npairs; ++i) {
        const uint32_t s = (uint32_t)(pairs[i] >> 32), c = Fs + (uint32_t)(pairs[i] & 0xFFFFFFFFu);
        for (uint32_t h = ps->face_off[s]; h < ps->face_off[s + 1]; ++h) tup[t++] = ((uint64_t)ps->face_edge[h] << 32) | c;
        for (uint32_t h = ps->face_off[c]; h < ps->face_off[c + 1]; ++h) tup[t++] = ((uint64_t)ps->face_edge[h] << 32) | s;
    }
    qsort(tup, ntup, sizeof(uint64_t), cmp_u64);
    size_t nuniq = 0;
    for (size_t i = 0; i < ntup; ++i)
        if (i == 0 || tup[i] != tup[i - 1]) tup[nuniq++] = tup[i];
    out->n_edge_face_before_cull = nuniq;

    /* ---- edge AABB cull (kernel.cpp:1989-2177) + tests (kernel.cpp:2415-2658) ---- */
    out->tests = (mco_test_t*)calloc(nuniq ? nuniq : 1, sizeof(mco_test_t));
    out->records = (mco_record_t*)calloc(nuniq ? nuniq : 1, sizeof(mco_record_t));
    int violated = 0;
    double stackv[3 * 16];
    for (size_t i = 0; i < nuniq; ++i) {
        const uint32_t e = (uint32_t)(tup[i] >> 32), g = (uint32_t)(tup[i] & 0xFFFFFFFFu);
        const double* q = ps->xyz + 3 * (size_t)ps->edge_v[2 * (size_t)e]; /* source(h0) */
        const double* r = ps->xyz + 3 * (size_t)ps->edge_v[2 * (size_t)e + 1]; /* target(h0) */
        double eb[6];
        for (int j = 0; j < 3; ++j) {
            eb[j] = q[j] < r[j] ? q[j] : r[j];
            eb[3 + j] = q[j] < r[j] ? r[j] : q[j];
        }
        const double* fb = g < Fs ? src_bboxes + 6 * (size_t)g : cut_bboxes + 6 * (size_t)(g - Fs);
        if (!overlap6(eb, fb)) continue;

        const uint32_t n = ps->face_off[g + 1] - ps->face_off[g];
        double* v = n <= 16 ? stackv : (double*)malloc(sizeof(double) * 3 * n);
        for (uint32_t k = 0; k < n; ++k) memcpy(v + 3 * k, ps->xyz + 3 * (size_t)ps->face_vtx[ps->face_off[g] + k], 24);
        const uint32_t slot = cand_slot[g];
        const double* normal = out->cand_normal + 3 * (size_t)slot;
        const double d = out->cand_d[slot];
        const int mc = out->cand_maxcomp[slot];

        mco_test_t* T = &out->tests[out->n_tests++];
        T->edge = e;
        T->face = g;
        double qres, rres;
        /* which of the two orient3d calls needed the exact stage (for the hardness statistics) */
        if (n == 3) {
            int cq, cr;
            (void)mco_orient3d_stageA(v, v + 3, v + 6, q, &cq);
            (void)mco_orient3d_stageA(v, v + 3, v + 6, r, &cr);
            T->exact_q = (uint8_t)!cq;
            T->exact_r = (uint8_t)!cr;
        }
        T->type = mco_segment_plane_type(q, r, v, (int)n, normal, mc, &qres, &rres);
        T->sign_q = (int8_t)sgn(qres);
        T->sign_r = (int8_t)sgn(rres);
        if (T->type == '1') {
            (void)mco_segment_plane_intersection(T->point, normal, d, q, r);
            T->pip = mco_point_in_polygon(T->point, v, (int)n, normal, mc);
            if (T->pip == 'v' || T->pip == 'e') violated = 1; /* kernel.cpp:2588-2597 */
            else if (T->pip == 'i') {
                mco_record_t* R = &out->records[out->n_records++];
                R->edge = e;
                R->face = g;
                memcpy(R->point, T->point, 24);
            }
        } else if (T->type != '0') { /* kernel.cpp:2518-2557 */
            const double* pts[2];
            int np = 0;
            if (T->type == 'q') pts[np++] = q;
            else if (T->type == 'r') pts[np++] = r;
            else {
                pts[np++] = q;
                pts[np++] = r;
            }
            for (int k = 0; k < np; ++k) {
                T->pip = mco_point_in_polygon(pts[k], v, (int)n, normal, mc);
                if (T->pip == 'i' || T->pip == 'v' || T->pip == 'e') {
                    violated = 1;
                    break;
                }
            }
        }
        if (v != stackv) free(v);
    }
    free(tup);
    free(is_cand);
    free(cand_slot);
    if (violated) {
        out->status = MCO_GENERAL_POSITION_VIOLATION;
        if (stop_on_gp) out->n_records = 0;
    }
    return out->status;
}


/* ======================================================================================================================
 * Input validation passes
 * ==================================================================================================================== */

/* source/kernel.cpp:235-364.  The reference walks the vertices in index order and floods each unvisited one breadth-first
 * through get_vertices_around_vertex; only the partition and the discovery order of the components escape, so a plain
 * adjacency-list flood in the same vertex order restates it. */
int mco_connected_components(uint32_t nv, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, int32_t* fccmap,
    int32_t* cc_vertex_count, int32_t* cc_face_count)
{
    const uint32_t nh = face_off[nf];
    uint32_t* deg = (uint32_t*)calloc((size_t)nv + 1, sizeof(uint32_t));
    uint32_t* adj = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (size_t)(nh ? nh : 1));
    int32_t* visited = (int32_t*)malloc(sizeof(int32_t) * (size_t)nv);
    uint32_t* queue = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nv);
    uint32_t f, i, v;
    int ncc = 0;
    for (f = 0; f < nf; ++f) {
        const uint32_t n = face_off[f + 1] - face_off[f];
        for (i = 0; i < n; ++i) {
            deg[face_vtx[face_off[f] + i] + 1] += 1;
            deg[face_vtx[face_off[f] + (i + 1) % n] + 1] += 1;
        }
    }
    for (v = 0; v < nv; ++v) deg[v + 1] += deg[v]; /* offsets */
    {
        uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nv);
        for (v = 0; v < nv; ++v) fill[v] = deg[v];
        for (f = 0; f < nf; ++f) {
            const uint32_t n = face_off[f + 1] - face_off[f];
            for (i = 0; i < n; ++i) {
                const uint32_t a = face_vtx[face_off[f] + i], b = face_vtx[face_off[f] + (i + 1) % n];
                adj[fill[a]++] = b;
                adj[fill[b]++] = a;
            }
        }
        free(fill);
    }
    for (v = 0; v < nv; ++v) visited[v] = -1;
    for (v = 0; v < nv; ++v) {
        uint32_t head = 0, tail = 0;
        if (visited[v] != -1) continue;
        visited[v] = ncc;
        cc_vertex_count[ncc] = 1;
        cc_face_count[ncc] = 0;
        queue[tail++] = v;
        while (head < tail) {
            const uint32_t u = queue[head++];
            uint32_t k;
            for (k = deg[u]; k < deg[u + 1]; ++k) {
                const uint32_t w = adj[k];
                if (visited[w] == -1) {
                    visited[w] = ncc;
                    cc_vertex_count[ncc] += 1;
                    queue[tail++] = w;
                }
            }
        }
        ++ncc;
    }
    for (f = 0; f < nf; ++f) { /* kernel.cpp:330-360: the component of the face's first vertex */
        const int32_t c = visited[face_vtx[face_off[f]]];
        fccmap[f] = c;
        cc_face_count[c] += 1;
    }
    free(deg);
    free(adj);
    free(visited);
    free(queue);
    return ncc;
}

/* source/preproc.cpp:1957-1990: a halfedge without a face exists exactly when an edge is used by one face only. */
uint32_t mco_border_edges(uint32_t nv, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf)
{
    /* count the uses of every unordered vertex pair with a sort of the pairs */
    const uint32_t nh = face_off[nf];
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nh ? nh : 1));
    uint32_t f, i, h = 0, border = 0;
    (void)nv;
    for (f = 0; f < nf; ++f) {
        const uint32_t n = face_off[f + 1] - face_off[f];
        for (i = 0; i < n; ++i) {
            const uint32_t a = face_vtx[face_off[f] + i], b = face_vtx[face_off[f] + (i + 1) % n];
            keys[h++] = ((uint64_t)(a < b ? a : b) << 32) | (a < b ? b : a);
        }
    }
    qsort(keys, nh, sizeof(uint64_t), cmp_u64);
    for (i = 0; i < nh;) {
        uint32_t j = i + 1;
        while (j < nh && keys[j] == keys[i]) ++j;
        if (j - i == 1) ++border;
        i = j;
    }
    free(keys);
    return border;
}


/* ======================================================================================================================
 * Winding number
 * ==================================================================================================================== */
static void w_sub(double* o, const double* a, const double* b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double w_dot(const double* a, const double* b)
{
    double r = 0.0; /* dot_product accumulates from 0.0 (math.h:634-642) */
    r += a[0] * b[0];
    r += a[1] * b[1];
    r += a[2] * b[2];
    return r;
}
static double w_len(const double* a) { return sqrt(w_dot(a, a)); } /* length() = sqrt(squared_length()) (math.h) */
static void w_cross(double* o, const double* a, const double* b)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void w_div(double* o, const double* a, double s) { o[0] = a[0] / s; o[1] = a[1] / s; o[2] = a[2] / s; }

static const double W_PI = 3.14159265358979323846;

/* source/preproc.cpp:1650-1698 */
double mco_solid_angle_tri(const double a[3], const double b[3], const double c[3], const double q[3])
{
    double qa[3], qb[3], qc[3], na[3], nb[3], nc[3], e1[3], e2[3], cr[3];
    double al, bl, cl, numerator, denominator;
    w_sub(qa, a, q);
    w_sub(qb, b, q);
    w_sub(qc, c, q);
    al = w_len(qa);
    bl = w_len(qb);
    cl = w_len(qc);
    if (al == 0.0 || bl == 0.0 || cl == 0.0) return 0.0;
    w_div(na, qa, al);
    w_div(nb, qb, bl);
    w_div(nc, qc, cl);
    w_sub(e1, nb, na);
    w_sub(e2, nc, na);
    w_cross(cr, e1, e2);
    numerator = w_dot(na, cr);
    if (numerator == 0.0) return 0.0;
    denominator = 1.0 + w_dot(na, nb) + w_dot(na, nc) + w_dot(nb, nc);
    return atan2(numerator, denominator) / (2. * W_PI);
}

/* source/preproc.cpp:1700-1810 */
double mco_solid_angle_quad(const double a[3], const double b[3], const double c[3], const double d[3], const double q[3])
{
    double v[4][3], len[4], diag02[3], diag13[3], v01[3], v23[3], cr[3], bary[4];
    double dot01, dot12, dot23, dot30, omega = 0.0;
    int i;
    w_sub(v[0], a, q);
    w_sub(v[1], b, q);
    w_sub(v[2], c, q);
    w_sub(v[3], d, q);
    for (i = 0; i < 4; ++i) len[i] = w_len(v[i]);
    if (len[0] == 0.0 || len[1] == 0.0 || len[2] == 0.0 || len[3] == 0.0) return 0.0;
    for (i = 0; i < 4; ++i) w_div(v[i], v[i], len[i]);
    w_sub(diag02, v[2], v[0]);
    w_sub(diag13, v[3], v[1]);
    w_sub(v01, v[1], v[0]);
    w_sub(v23, v[3], v[2]);
    w_cross(cr, v23, diag13);
    bary[0] = w_dot(v[3], cr);
    w_cross(cr, v23, diag02);
    bary[1] = -w_dot(v[2], cr);
    w_cross(cr, v01, diag13);
    bary[2] = -w_dot(v[1], cr);
    w_cross(cr, v01, diag02);
    bary[3] = w_dot(v[0], cr);
    dot01 = w_dot(v[0], v[1]);
    dot12 = w_dot(v[1], v[2]);
    dot23 = w_dot(v[2], v[3]);
    dot30 = w_dot(v[3], v[0]);
    if (bary[0] * bary[2] < bary[1] * bary[3]) { /* split 0-2 */
        const double n012 = bary[3], n023 = bary[1], dot02 = w_dot(v[0], v[2]);
        if (n012 != 0.0) omega = atan2(n012, 1.0 + dot01 + dot12 + dot02);
        if (n023 != 0.0) omega += atan2(n023, 1.0 + dot02 + dot23 + dot30);
    } else { /* split 1-3 */
        const double n013 = -bary[2], n123 = -bary[0], dot13 = w_dot(v[1], v[3]);
        if (n013 != 0.0) omega = atan2(n013, 1.0 + dot01 + dot13 + dot30);
        if (n123 != 0.0) omega += atan2(n123, 1.0 + dot12 + dot23 + dot13);
    }
    return omega / (2. * W_PI);
}

/* source/preproc.cpp:1812-1955, sequential form */
double mco_winding_number(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, const double q[3])
{
    double wn = 0.0;
    uint32_t f;
    for (f = 0; f < nf; ++f) {
        const uint32_t h = face_off[f], n = face_off[f + 1] - h;
        const double *a = xyz + 3 * (size_t)face_vtx[h], *b = xyz + 3 * (size_t)face_vtx[h + 1], *c = xyz + 3 * (size_t)face_vtx[h + 2];
        if (n == 3) wn += mco_solid_angle_tri(a, b, c, q);
        else if (n == 4) wn += mco_solid_angle_quad(a, b, c, xyz + 3 * (size_t)face_vtx[h + 3], q);
        else return NAN;
    }
    return wn;
}


/* ---- cut-path segment table ------------------------------------------------------------------------------ */
typedef struct {
    uint64_t key;
    uint32_t v;
} cp_entry_t;

static int cp_entry_cmp(const void* a, const void* b)
{
    const cp_entry_t *x = (const cp_entry_t*)a, *y = (const cp_entry_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->v < y->v ? -1 : (x->v > y->v ? 1 : 0);
}

/* source/kernel.cpp:1496-1531 with the arithmetic of math.h:635-642 (dot product accumulated from 0.0 in x, y, z order),
 * :718-721 (normalize = v / sqrt(dot(v, v))); std::sort on a handful of elements is an insertion sort */
static void cp_linear_projection_sort(const mco_record_t* rec, uint32_t* v, uint32_t n)
{
    const double* o = rec[v[0]].point;
    const double* d = rec[v[1]].point;
    double dir[3], len2 = 0.0;
    for (int k = 0; k < 3; ++k) dir[k] = o[k] - d[k];
    for (int k = 0; k < 3; ++k) len2 += dir[k] * dir[k];
    const double len = sqrt(len2);
    for (int k = 0; k < 3; ++k) dir[k] = dir[k] / len;
    double* proj = (double*)malloc(sizeof(double) * n);
    for (uint32_t i = 0; i < n; ++i) {
        const double* p = rec[v[i]].point;
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += (o[k] - p[k]) * dir[k];
        proj[i] = acc;
    }
    for (uint32_t i = 1; i < n; ++i) { /* stable insertion sort, ascending projection */
        const double x = proj[i];
        const uint32_t xv = v[i];
        uint32_t j = i;
        while (j > 0 && x < proj[j - 1]) {
            proj[j] = proj[j - 1];
            v[j] = v[j - 1];
            --j;
        }
        proj[j] = x;
        v[j] = xv;
    }
    free(proj);
}

int mco_cutpath_segments(const uint32_t* edge_f, uint32_t src_nf, const mco_record_t* rec, size_t n, mco_cutpath_t* out)
{
    memset(out, 0, sizeof(*out));
    cp_entry_t* e = (cp_entry_t*)malloc(sizeof(cp_entry_t) * (2 * n + 1));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        /* kernel.cpp:2610-2640: the tested edge's faces; the face of h0 unless h0 is a border halfedge */
        const uint32_t h0f = edge_f[2 * (size_t)rec[i].edge], h1f = edge_f[2 * (size_t)rec[i].edge + 1];
        const uint32_t tested = rec[i].face;
        const uint32_t own = h0f != MCO_NULL ? h0f : h1f;
        const uint32_t other = own == h0f ? h1f : MCO_NULL; /* (an edge whose h0 has no face has no second face to add) */
        if (own == MCO_NULL) {
            free(e);
            return -1;
        }
        const int edge_is_cut = own >= src_nf;
        /* key = {source-mesh face, cut-mesh face} */
        e[m].key = edge_is_cut ? ((uint64_t)tested << 32 | own) : ((uint64_t)own << 32 | tested);
        e[m++].v = (uint32_t)i;
        if (other != MCO_NULL) {
            e[m].key = edge_is_cut ? ((uint64_t)tested << 32 | other) : ((uint64_t)other << 32 | tested);
            e[m++].v = (uint32_t)i;
        }
    }
    qsort(e, m, sizeof(cp_entry_t), cp_entry_cmp);
    out->key = (uint64_t*)malloc(sizeof(uint64_t) * (m + 1));
    out->off = (uint32_t*)malloc(sizeof(uint32_t) * (m + 2));
    out->vtx = (uint32_t*)malloc(sizeof(uint32_t) * (m + 1));
    size_t g = 0;
    for (size_t i = 0; i < m; ++i) {
        if (i == 0 || e[i].key != e[i - 1].key) {
            out->key[g] = e[i].key;
            out->off[g++] = (uint32_t)i;
        }
        out->vtx[i] = e[i].v;
    }
    out->off[g] = (uint32_t)m;
    out->n_groups = g;
    out->n_entries = m;
    for (size_t k = 0; k < g; ++k) {
        const uint32_t c = out->off[k + 1] - out->off[k];
        if (c == 1) out->n_single++;
        if (c > 2) cp_linear_projection_sort(rec, out->vtx + out->off[k], c);
    }
    free(e);
    return 0;
}

void mco_cutpath_free(mco_cutpath_t* o)
{
    free(o->key);
    free(o->off);
    free(o->vtx);
    memset(o, 0, sizeof(*o));
}
