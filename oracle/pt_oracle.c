/*
 * pt_oracle.c — plain-C restatement of the reference's render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; nothing under
 * cpupathtrace_b200/ or include/ links, includes or calls it.
 *
 * What it restates (reference = johannesschaeufele/CPUPathTrace, paths relative to /root/reference):
 *   xorshift / RandomEngine                include/PathTrace/base.h:24-58
 *   libstdc++ 13 uniform_real / bernoulli  bits/random.tcc generate_canonical, bits/random.h (toolchain, not vendored)
 *   AABB::getIntersection                  src/scene/bounding_box.cpp:38-73
 *   Triangle / Sphere primitives           src/scene/object.cpp:72-207
 *   impl::constructBVH                     src/scene/scene.cpp:12-102
 *   impl::getChildIntersection             src/scene/scene.cpp:104-150      (kept RECURSIVE, as the reference)
 *   Scene::Scene, registerEmissiveObjects  src/scene/scene.cpp:153-208
 *   Scene::getIntersection, sampleLights   src/scene/scene.cpp:210-289
 *   BSDFs                                  src/scene/propagation.cpp:11-217
 *   Camera::shootRay, aperture samplers    src/camera.cpp:7-113
 *   impl::getSample, processItem           src/worker.cpp:26-326
 *
 * Pinning: tests/test_oracle.py checks this restatement against oracle/_ref (the unmodified reference compiled from
 * its own sources) — closest hits, per-sample radiance and whole processItem tiles must agree BIT FOR BIT, since both
 * run the same libm on the same host — and against the reference's own known-answer tests.
 *
 * It deliberately shares no code with the CUDA implementation: the scene arrives as the C-ABI's POD description
 * (include/ptb.h) and everything else is rebuilt here with the reference's data-structure shapes (a node tree, a
 * recursive traversal, one sequential engine per tile).  Compile with -ffp-contract=off and no -march.
 */
#include "ptb.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ vectors */

typedef struct {
    float x, y, z;
} v3;

typedef struct {
    float c[4];
} rgba;

static v3 V(float x, float y, float z) {
    v3 r = {x, y, z};
    return r;
}
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* util/vector.h:196-205: sum starts at 0 and adds products in index order */
static float vdot(v3 a, v3 b) {
    float s = 0.0F;
    s += a.x * b.x;
    s += a.y * b.y;
    s += a.z * b.z;
    return s;
}
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float vlen2(v3 a) { return vdot(a, a); }
static float vlen(v3 a) { return sqrtf(vlen2(a)); }
/* util/vector.h:161-167 */
static v3 vnorm(v3 a) {
    float inv = 1.0F / vlen(a);
    return vmul(a, inv);
}
/* util/vector.h:250-255: v - n * 2 * d */
static v3 vreflect(v3 v, v3 n) {
    float d = vdot(v, n);
    return vsub(v, vmul(vmul(n, 2.0F), d));
}
static float fminstd(float a, float b) { return b < a ? b : a; } /* std::min */
static float fmaxstd(float a, float b) { return a < b ? b : a; } /* std::max */

static rgba cadd(rgba a, rgba b) {
    rgba r;
    for(int i = 0; i < 4; i++) r.c[i] = a.c[i] + b.c[i];
    return r;
}
static rgba csub(rgba a, rgba b) {
    rgba r;
    for(int i = 0; i < 4; i++) r.c[i] = a.c[i] - b.c[i];
    return r;
}
static rgba cmul(rgba a, rgba b) {
    rgba r;
    for(int i = 0; i < 4; i++) r.c[i] = a.c[i] * b.c[i];
    return r;
}
static rgba cscale(rgba a, float s) {
    rgba r;
    for(int i = 0; i < 4; i++) r.c[i] = a.c[i] * s;
    return r;
}
static rgba cdiv(rgba a, float s) {
    rgba r;
    for(int i = 0; i < 4; i++) r.c[i] = a.c[i] / s;
    return r;
}
static rgba czero(void) {
    rgba r = {{0.0F, 0.0F, 0.0F, 0.0F}};
    return r;
}

/* ------------------------------------------------------------------------------------------------ random numbers */

typedef struct {
    uint64_t s;
} engine;

/* base.h:26 */
static engine engine_seed(uint64_t seed) {
    engine e;
    e.s = seed ^ (~seed << 32);
    return e;
}
/* base.h:28-35 */
static uint32_t engine_next(engine *e) {
    uint64_t result = e->s * 0xD989BCACC137DCD5ULL;
    e->s ^= e->s >> 11;
    e->s ^= e->s << 31;
    e->s ^= e->s >> 18;
    return (uint32_t)(result >> 32);
}
/* generate_canonical<float, 24> over a 32-bit engine: one draw */
static float canonical_f(engine *e) {
    float sum = (float)engine_next(e) * 1.0F;
    float ret = sum / 4294967296.0F;
    if(ret >= 1.0F) ret = nextafterf(1.0F, 0.0F);
    return ret;
}
/* uniform_real_distribution<float>(a, b) */
static float uniform_f(engine *e, float a, float b) { return canonical_f(e) * (b - a) + a; }
static float uniform01(engine *e) { return uniform_f(e, 0.0F, 1.0F); }
/* bernoulli_distribution(p): generate_canonical<double, 53> = two draws */
static int bernoulli(engine *e, double p) {
    double sum = 0.0;
    double tmp = 1.0;
    sum += (double)engine_next(e) * tmp;
    tmp *= 4294967296.0;
    sum += (double)engine_next(e) * tmp;
    tmp *= 4294967296.0;
    double ret = sum / tmp;
    if(ret >= 1.0) ret = nextafter(1.0, 0.0);
    return (ret - 0.0) < p * (1.0 - 0.0);
}

/* ------------------------------------------------------------------------------------------------ scene */

typedef struct {
    float lo[3], hi[3];
    int left, right; /* child node indices, -1 for leaves */
    int prim;        /* leaf: primitive index, else -1 */
} onode;

struct pto_scene {
    ptb_prim *prims;
    uint64_t n_prims;
    ptb_material *materials;
    uint32_t n_materials;
    ptb_point_light *lights;
    uint32_t n_lights;
    onode *nodes;
    int n_nodes;
    int root; /* -1: empty scene (NullObject leaf) */
    int *emissive;
    float *cdf;
    int n_emissive;
    int depth;
};
typedef struct pto_scene pto_scene;

static void prim_bounds(const ptb_prim *p, float lo[3], float hi[3]) {
    if(p->kind == PTB_PRIM_TRIANGLE) { /* object.cpp:184-186 */
        for(int c = 0; c < 3; c++) {
            lo[c] = fminstd(fminstd(p->p[c], p->p[3 + c]), p->p[6 + c]);
            hi[c] = fmaxstd(fmaxstd(p->p[c], p->p[3 + c]), p->p[6 + c]);
        }
    }
    else if(p->kind == PTB_PRIM_SPHERE) { /* object.cpp:90-93 */
        for(int c = 0; c < 3; c++) {
            lo[c] = p->p[c] - p->p[3];
            hi[c] = p->p[c] + p->p[3];
        }
    }
    else {
        for(int c = 0; c < 3; c++) lo[c] = hi[c] = 0.0F;
    }
}

static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* scene.cpp:12-102; `items` are leaf node indices in list order; returns the index of the subtree's root node */
static int build_bvh(pto_scene *s, int *items, int n, int depth) {
    if(depth > s->depth) s->depth = depth;
    if(n == 1) return items[0];

    float medians[3];
    float *coords = (float *)malloc(sizeof(float) * (size_t)n);
    for(int dim = 0; dim < 3; dim++) {
        for(int i = 0; i < n; i++) coords[i] = s->nodes[items[i]].lo[dim];
        qsort(coords, (size_t)n, sizeof(float), cmp_float); /* nth_element only fixes the value at that rank */
        medians[dim] = coords[n / 2 - 1];
    }
    free(coords);

    float areas[3];
    for(int dim = 0; dim < 3; dim++) {
        float lo[2][3], hi[2][3];
        for(int g = 0; g < 2; g++)
            for(int c = 0; c < 3; c++) {
                lo[g][c] = INFINITY;
                hi[g][c] = -INFINITY;
            }
        for(int i = 0; i < n; i++) {
            const onode *b = &s->nodes[items[i]];
            int g = b->lo[dim] <= medians[dim] ? 0 : 1;
            for(int c = 0; c < 3; c++) {
                lo[g][c] = fminstd(lo[g][c], b->lo[c]);
                hi[g][c] = fmaxstd(hi[g][c], b->hi[c]);
            }
        }
        float area = 0.0F;
        for(int g = 0; g < 2; g++) {
            float d0 = hi[g][0] - lo[g][0], d1 = hi[g][1] - lo[g][1], d2 = hi[g][2] - lo[g][2];
            area += 2 * (d0 * d1 + d1 * d2 + d0 * d2);
        }
        areas[dim] = area;
    }
    int axis = 0;
    float best = areas[0];
    for(int dim = 1; dim < 3; dim++)
        if(areas[dim] < best) {
            best = areas[dim];
            axis = dim;
        }

    int *left = (int *)malloc(sizeof(int) * (size_t)n);
    int *right = (int *)malloc(sizeof(int) * (size_t)n);
    int nl = 0, nr = 0;
    for(int i = 0; i < n; i++) {
        if(s->nodes[items[i]].lo[axis] <= medians[axis])
            left[nl++] = items[i];
        else
            right[nr++] = items[i];
    }
    while(nl > 1 && nl > 2 * nr) { /* scene.cpp:90-94 */
        right[nr++] = left[nl - 1];
        nl--;
    }

    int l = build_bvh(s, left, nl, depth + 1);
    int r = build_bvh(s, right, nr, depth + 1);
    free(left);
    free(right);

    int id = s->n_nodes++;
    onode *node = &s->nodes[id];
    for(int c = 0; c < 3; c++) { /* bounding_box.cpp:8-12 */
        node->lo[c] = fminstd(s->nodes[l].lo[c], s->nodes[r].lo[c]);
        node->hi[c] = fmaxstd(s->nodes[l].hi[c], s->nodes[r].hi[c]);
    }
    node->left = l;
    node->right = r;
    node->prim = -1;
    return id;
}

static float prim_area(const ptb_prim *p) {
    if(p->kind == PTB_PRIM_TRIANGLE) { /* object.cpp:188-190 */
        v3 a = V(p->p[0], p->p[1], p->p[2]), b = V(p->p[3], p->p[4], p->p[5]), c = V(p->p[6], p->p[7], p->p[8]);
        return vlen(vcross(vsub(b, a), vsub(c, a))) / 2.0F;
    }
    if(p->kind == PTB_PRIM_SPHERE) { /* object.cpp:95-99 */
        const float pi = (float)M_PI;
        float radius2 = p->p[3] * p->p[3];
        return 4.0F * pi * radius2;
    }
    return 0.0F;
}

/* scene.cpp:183-208 */
static void register_emissive(pto_scene *s, int node) {
    const onode *n = &s->nodes[node];
    if(n->prim >= 0) {
        const ptb_prim *p = &s->prims[n->prim];
        const float *e = s->materials[p->material].emission;
        float power = (e[0] + e[1] + e[2]) * e[3];
        if(power <= 0.0F) return;
        float probability = power * prim_area(p);
        if(probability <= 0.0F) return;
        s->emissive[s->n_emissive] = n->prim;
        s->cdf[s->n_emissive] = probability;
        s->n_emissive++;
    }
    else {
        register_emissive(s, n->left);
        register_emissive(s, n->right);
    }
}

pto_scene *pto_scene_create(const ptb_scene_desc *desc) {
    pto_scene *s = (pto_scene *)calloc(1, sizeof(pto_scene));
    s->n_prims = desc->n_prims;
    s->n_materials = desc->n_materials;
    s->n_lights = desc->n_lights;
    s->prims = (ptb_prim *)malloc(sizeof(ptb_prim) * (size_t)(desc->n_prims + 1));
    s->materials = (ptb_material *)malloc(sizeof(ptb_material) * (size_t)(desc->n_materials + 1));
    s->lights = (ptb_point_light *)malloc(sizeof(ptb_point_light) * (size_t)(desc->n_lights + 1));
    if(desc->n_prims) memcpy(s->prims, desc->prims, sizeof(ptb_prim) * (size_t)desc->n_prims);
    if(desc->n_materials) memcpy(s->materials, desc->materials, sizeof(ptb_material) * (size_t)desc->n_materials);
    if(desc->n_lights) memcpy(s->lights, desc->lights, sizeof(ptb_point_light) * (size_t)desc->n_lights);

    int n = (int)desc->n_prims;
    s->nodes = (onode *)malloc(sizeof(onode) * (size_t)(2 * n + 1));
    s->emissive = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    s->cdf = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    s->root = -1;
    if(n > 0) {
        int *items = (int *)malloc(sizeof(int) * (size_t)n);
        for(int i = 0; i < n; i++) { /* scene.cpp:156-160: one leaf per object */
            prim_bounds(&s->prims[i], s->nodes[i].lo, s->nodes[i].hi);
            s->nodes[i].left = s->nodes[i].right = -1;
            s->nodes[i].prim = i;
            items[i] = i;
        }
        s->n_nodes = n;
        s->root = build_bvh(s, items, n, 1);
        free(items);
        register_emissive(s, s->root);
    }
    /* scene.cpp:165-180 */
    float cumulative = 0.0F;
    for(int i = 0; i < s->n_emissive; i++) {
        float probability = s->cdf[i];
        s->cdf[i] += cumulative;
        cumulative += probability;
    }
    for(int i = 0; i < s->n_emissive; i++) s->cdf[i] /= cumulative;
    return s;
}

void pto_scene_destroy(pto_scene *s) {
    if(!s) return;
    free(s->prims);
    free(s->materials);
    free(s->lights);
    free(s->nodes);
    free(s->emissive);
    free(s->cdf);
    free(s);
}

int pto_scene_depth(const pto_scene *s) { return s->depth; }
int pto_scene_emissive_count(const pto_scene *s) { return s->n_emissive; }

/* ------------------------------------------------------------------------------------------------ intersection */

typedef struct {
    v3 o, d;
} ray;

/* bounding_box.cpp:38-73 */
static float box_hit(const float lo[3], const float hi[3], const ray *r) {
    float ix = fabsf(r->d.x) > 0.0F ? 1.0F / r->d.x : FLT_MAX;
    float iy = fabsf(r->d.y) > 0.0F ? 1.0F / r->d.y : FLT_MAX;
    float iz = fabsf(r->d.z) > 0.0F ? 1.0F / r->d.z : FLT_MAX;
    float t1 = (lo[0] - r->o.x) * ix, t2 = (hi[0] - r->o.x) * ix;
    float t3 = (lo[1] - r->o.y) * iy, t4 = (hi[1] - r->o.y) * iy;
    float t5 = (lo[2] - r->o.z) * iz, t6 = (hi[2] - r->o.z) * iz;
    float t_min = fmaxstd(fmaxstd(fminstd(t1, t2), fminstd(t3, t4)), fminstd(t5, t6));
    float t_max = fminstd(fminstd(fmaxstd(t1, t2), fmaxstd(t3, t4)), fmaxstd(t5, t6));
    float t = t_min;
    if(t_max < 0.0F || t_min > t_max) return -1.0F;
    if(t_min < 0.0F && t_min <= t_max && t_max >= 0.0F) t = 0.0F;
    return t;
}

/* object.cpp:146-182 and :72-84 */
static float prim_hit(const ptb_prim *p, const ray *r) {
    if(p->kind == PTB_PRIM_TRIANGLE) {
        const float epsilon = 1E-6F;
        v3 a = V(p->p[0], p->p[1], p->p[2]), b = V(p->p[3], p->p[4], p->p[5]), c = V(p->p[6], p->p[7], p->p[8]);
        v3 ab = vsub(b, a), ac = vsub(c, a);
        v3 pvec = vcross(r->d, ac);
        float det = vdot(ab, pvec);
        if(p->cull_backface) {
            if(det <= epsilon) return -1.0F;
        }
        else {
            if(fabsf(det) <= epsilon) return -1.0F;
        }
        float inv_det = 1.0F / det;
        v3 tvec = vsub(r->o, a);
        float u = vdot(tvec, pvec) * inv_det;
        if(u < 0 || u > 1) return -1.0F;
        v3 qvec = vcross(tvec, ab);
        float v = vdot(r->d, qvec) * inv_det;
        if(v < 0 || u + v > 1) return -1.0F;
        return vdot(ac, qvec) * inv_det;
    }
    if(p->kind == PTB_PRIM_SPHERE) {
        v3 co = vsub(r->o, V(p->p[0], p->p[1], p->p[2]));
        float d = vdot(r->d, co);
        float radius2 = p->p[3] * p->p[3];
        float discriminant = d * d - vlen2(co) + radius2;
        if(discriminant >= 0) return -(d + sqrtf(discriminant));
        return -1.0F;
    }
    return -1.0F;
}

typedef struct {
    float t;
    int prim; /* -1: none */
} hit;

/* scene.cpp:104-150, recursion kept */
static hit child_hit(const pto_scene *s, int node, const ray *r, float t_max) {
    const onode *n = &s->nodes[node];
    hit none = {-1.0F, -1};
    if(n->prim >= 0) {
        hit h = {prim_hit(&s->prims[n->prim], r), n->prim};
        return h;
    }
    float left_t = box_hit(s->nodes[n->left].lo, s->nodes[n->left].hi, r);
    float right_t = box_hit(s->nodes[n->right].lo, s->nodes[n->right].hi, r);
    float close_t = fminstd(left_t, right_t);
    float far_t = fmaxstd(left_t, right_t);
    int close_node = left_t < right_t ? n->left : n->right;
    int far_node = left_t < right_t ? n->right : n->left;

    hit close_hit = none;
    if(close_t >= 0.0F && close_t < t_max) close_hit = child_hit(s, close_node, r, t_max);
    if(close_hit.t >= 0.0F) {
        if(close_hit.t < far_t) return close_hit;
        t_max = fminstd(t_max, close_hit.t);
    }
    if(far_t >= 0.0F && far_t < t_max) {
        hit far_hit = child_hit(s, far_node, r, t_max);
        if(far_hit.t < 0.0F || (close_hit.t >= 0.0F && close_hit.t < far_hit.t)) return close_hit;
        return far_hit;
    }
    return close_hit;
}

/* scene.cpp:210-220 */
static hit scene_hit(const pto_scene *s, const ray *r) {
    hit none = {-1.0F, -1};
    if(s->root < 0) return none; /* NullObject leaf: whatever the box says, the primitive reports -1 */
    float t = box_hit(s->nodes[s->root].lo, s->nodes[s->root].hi, r);
    if(t >= 0.0F) return child_hit(s, s->root, r, FLT_MAX);
    hit miss = {t, -1};
    return miss;
}

void pto_intersect(const pto_scene *s, const float *rays, uint64_t n, float *t_out, int32_t *prim_out) {
    for(uint64_t i = 0; i < n; i++) {
        ray r = {V(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), V(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
        hit h = scene_hit(s, &r);
        t_out[i] = h.t;
        prim_out[i] = (h.t >= 0.0F) ? h.prim : -1;
    }
}

/* ------------------------------------------------------------------------------------------------ certificate
 *
 * CPU statement of the "certified closest hit" the CUDA kernels use (cpupathtrace_b200/csrc/traverse.cuh): walk ANY
 * hierarchy whose inner boxes are the exact unions of their children's boxes, nearest child first, pruning subtrees
 * whose entry exceeds t_best (1 + 2^-7), and keep the result only if it is strictly nearer than every other hit found,
 * no other hit lies at or before its own leaf-box entry, that entry does not exceed t_best (1 + 2^-9), and t_best > 0;
 * a ray that meets a primitive whose test reports a hit more than 2^-8 in front of the primitive's own box is abandoned
 * (certain = 0) on the spot.
 * The claim under test: whenever the walk returns certain = 1, (t, primitive) equals what the reference walk
 * (scene_hit above, scene.cpp:104-150) returns on the reference tree -- whatever the shape of the hierarchy walked here.
 * The hierarchy is built from `tree_seed`: a random permutation of the primitives split at random positions
 * (shape 0), at position 1 (shape 1: a chain, the most unbalanced tree) or in the middle (shape 2).
 */
typedef struct {
    float lo[3], hi[3];
    int left, right; /* >= 0: inner node, < 0: ~primitive */
} cnode;

static uint64_t cert_rand(uint64_t *state) { /* splitmix64 */
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static void cert_ref_box(const pto_scene *s, const cnode *nodes, int ref, float lo[3], float hi[3]) {
    if(ref < 0) {
        prim_bounds(&s->prims[~ref], lo, hi);
    }
    else {
        for(int c = 0; c < 3; c++) {
            lo[c] = nodes[ref].lo[c];
            hi[c] = nodes[ref].hi[c];
        }
    }
}

static int cert_build(const pto_scene *s, cnode *nodes, int *n_nodes, int *items, int n, int shape, uint64_t *state) {
    if(n == 1) return ~items[0];
    int split = shape == 1 ? 1 : (shape == 2 ? n / 2 : 1 + (int)(cert_rand(state) % (uint64_t)(n - 1)));
    int left = cert_build(s, nodes, n_nodes, items, split, shape, state);
    int right = cert_build(s, nodes, n_nodes, items + split, n - split, shape, state);
    int id = (*n_nodes)++;
    float llo[3], lhi[3], rlo[3], rhi[3];
    cert_ref_box(s, nodes, left, llo, lhi);
    cert_ref_box(s, nodes, right, rlo, rhi);
    for(int c = 0; c < 3; c++) {
        nodes[id].lo[c] = fminstd(llo[c], rlo[c]);
        nodes[id].hi[c] = fmaxstd(lhi[c], rhi[c]);
    }
    nodes[id].left = left;
    nodes[id].right = right;
    return id;
}


/* ------------------------------------------------------------------------------------------------ certificate guard
 *
 * Restatement of cpupathtrace_b200/csrc/cert_guard.cpp (the product builds the same table on the host and the kernels
 * apply it when a ray is fetched).  The certificate of the walk below is exact about everything the walk TESTS; about a
 * primitive R it never reaches (box entry e_R > t_Q (1 + s)) it assumes that R's own intersection routine cannot report
 * a distance below e_R (1 - s/2).  That is a statement about fp32 rounding in Triangle / Sphere::getIntersection
 * (object.cpp:72-84, 146-182), and it is false for large triangles met at grazing incidence and for spheres grazed
 * from nearby.  Error model (eps = 2^-24, safety factor K = 2):
 *   triangle  |t_computed - t_true| <= 9 eps K |ab||ac| (|o - a| + 2 t_true) / |det|,  only for |det| > 1e-6
 *   sphere    |t_computed - t_true| <= 3 eps K |co| + min(Ddisc / (2 sqrt(disc)), sqrt(Ddisc)), Ddisc = 4 eps K max(|co|^2, r^2)
 * A triangle whose worst case (|det| = 1e-6) still keeps the relative part below s/12 is SAFE: only the absolute term
 * dmax diam remains, which the certificate covers by demanding t_Q >= tau_safe.  Every other triangle contributes a
 * PLANE (unit normal n, offset h, half thickness w of its box across the plane, cone, band): a ray is handed straight
 * to the reference-order walk when it is nearly parallel to the plane (|n.d| < cone), starts inside the slab
 * (|n.o - h| <= w), or approaches the plane from less than w + band.  Spheres: rays starting closer than 1.8 r to the
 * centre, or grazing (|disc| < r^2 / 16) from less than 3 r.  More than PTO_GUARD_PLANES distinct planes,
 * PTO_GUARD_SPHERES spheres, or a sliver whose cone would exceed 0.25, switch certification off for the scene.
 */
#define PTO_GUARD_PLANES 24
#define PTO_GUARD_SPHERES 8

typedef struct {
    int enabled;
    float slack;    /* s */
    float tau_safe; /* certificate needs t_Q >= tau_safe */
    int n_planes;
    float plane_n[PTO_GUARD_PLANES][3];
    float plane_h[PTO_GUARD_PLANES];
    float plane_w[PTO_GUARD_PLANES];
    float plane_cone[PTO_GUARD_PLANES];
    float plane_k[PTO_GUARD_PLANES]; /* band = k * (|o - box centre| + half diagonal) */
    float plane_lo[PTO_GUARD_PLANES][3]; /* box of the triangles behind the plane */
    float plane_hi[PTO_GUARD_PLANES][3];
    float plane_r[PTO_GUARD_PLANES];
    int n_spheres;
    float sphere[PTO_GUARD_SPHERES][4];
} pto_guard;

static void guard_build(const pto_scene *s, pto_guard *g) {
    const double eps = 5.9604644775390625e-8, K = 2.0, slack = 0.0078125;
    memset(g, 0, sizeof(*g));
    g->enabled = 1;
    g->slack = (float)slack;
    double tau_safe = 0.0;
    for(uint64_t i = 0; i < s->n_prims && g->enabled; i++) {
        const ptb_prim *p = &s->prims[i];
        if(p->kind == PTB_PRIM_SPHERE) {
            if(g->n_spheres == PTO_GUARD_SPHERES) {
                g->enabled = 0;
                break;
            }
            for(int c = 0; c < 4; c++) g->sphere[g->n_spheres][c] = p->p[c];
            g->n_spheres++;
            continue;
        }
        if(p->kind != PTB_PRIM_TRIANGLE) continue;
        double a[3], ab[3], ac[3], bc[3];
        for(int c = 0; c < 3; c++) {
            a[c] = p->p[c];
            ab[c] = (double)p->p[3 + c] - a[c];
            ac[c] = (double)p->p[6 + c] - a[c];
            bc[c] = ac[c] - ab[c];
        }
        double lab = sqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]), lac = sqrt(ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2]);
        double lbc = sqrt(bc[0] * bc[0] + bc[1] * bc[1] + bc[2] * bc[2]);
        double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
        double area2 = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double m = lab * lac, diam = fmax(lab, fmax(lac, lbc));
        if(!(area2 > 0.0) || !(m > 0.0)) continue; /* degenerate: |det| = 0 for every ray, always rejected */
        double dmax = 9.0 * eps * K * m / 1e-6;
        if(dmax <= slack / 12.0) {
            tau_safe = fmax(tau_safe, dmax * diam / (slack / 4.0));
            continue;
        }
        double c_u = 9.0 * eps * K * m / area2; /* relative error = c_u / |cos| */
        double cone = 12.0 * c_u / slack;
        if(cone > 0.25) { /* a sliver: dangerous from almost every direction */
            g->enabled = 0;
            break;
        }
        for(int c = 0; c < 3; c++) n[c] /= area2;
        double h = n[0] * a[0] + n[1] * a[1] + n[2] * a[2];
        /* half thickness across the plane of the fp32 box: the box of a triangle is flat only for axis-aligned ones */
        float lo[3], hi[3];
        prim_bounds(p, lo, hi);
        double w = 0.0;
        for(int corner = 0; corner < 8; corner++) {
            double x = (corner & 1) ? hi[0] : lo[0], y = (corner & 2) ? hi[1] : lo[1], z = (corner & 4) ? hi[2] : lo[2];
            w = fmax(w, fabs(n[0] * x + n[1] * y + n[2] * z - h));
        }
        double k_u = 4.0 * c_u / slack;
        /* merge with an existing plane: same normal up to sign, same offset */
        int merged = 0;
        for(int j = 0; j < g->n_planes && !merged; j++) {
            double dotn = n[0] * g->plane_n[j][0] + n[1] * g->plane_n[j][1] + n[2] * g->plane_n[j][2];
            double hj = dotn >= 0.0 ? h : -h;
            if(fabs(fabs(dotn) - 1.0) < 1e-8 && fabs(hj - g->plane_h[j]) <= 1e-6 * (1.0 + fabs(hj))) {
                for(int c = 0; c < 3; c++) {
                    g->plane_lo[j][c] = fminf(g->plane_lo[j][c], lo[c]);
                    g->plane_hi[j][c] = fmaxf(g->plane_hi[j][c], hi[c]);
                }
                g->plane_w[j] = (float)fmax(g->plane_w[j], w * 1.0001 + 1e-7 * (1.0 + fabs(h)));
                g->plane_cone[j] = (float)fmax(g->plane_cone[j], cone * 1.0001);
                g->plane_k[j] = (float)fmax(g->plane_k[j], k_u * 1.0001);
                merged = 1;
            }
        }
        if(merged) continue;
        if(g->n_planes == PTO_GUARD_PLANES) {
            g->enabled = 0;
            break;
        }
        int j = g->n_planes++;
        for(int c = 0; c < 3; c++) {
            g->plane_n[j][c] = (float)n[c];
            g->plane_lo[j][c] = lo[c];
            g->plane_hi[j][c] = hi[c];
        }
        g->plane_h[j] = (float)h;
        g->plane_w[j] = (float)(w * 1.0001 + 1e-7 * (1.0 + fabs(h)));
        g->plane_cone[j] = (float)(cone * 1.0001);
        g->plane_k[j] = (float)(k_u * 1.0001);
    }
    for(int j = 0; j < g->n_planes; j++) {
        double dx = (double)g->plane_hi[j][0] - g->plane_lo[j][0], dy = (double)g->plane_hi[j][1] - g->plane_lo[j][1], dz = (double)g->plane_hi[j][2] - g->plane_lo[j][2];
        g->plane_r[j] = (float)(0.5 * sqrt(dx * dx + dy * dy + dz * dz) * 1.0001);
    }
    g->tau_safe = (float)(tau_safe * 1.0001);
}

/* 1: the ray must be traced by the reference-order walk (same arithmetic as guardFlagsRay in csrc/traverse.cuh) */
static int guard_flags_ray(const pto_guard *g, const ray *r) {
    if(!g->enabled) return 1;
    for(int j = 0; j < g->n_planes; j++) {
        float hd = ((g->plane_n[j][0] * r->o.x + g->plane_n[j][1] * r->o.y) + g->plane_n[j][2] * r->o.z) - g->plane_h[j];
        float cd = (g->plane_n[j][0] * r->d.x + g->plane_n[j][1] * r->d.y) + g->plane_n[j][2] * r->d.z;
        float ahd = fabsf(hd);
        float dx = r->o.x - 0.5F * (g->plane_lo[j][0] + g->plane_hi[j][0]);
        float dy = r->o.y - 0.5F * (g->plane_lo[j][1] + g->plane_hi[j][1]);
        float dz = r->o.z - 0.5F * (g->plane_lo[j][2] + g->plane_hi[j][2]);
        float band = g->plane_k[j] * (sqrtf((dx * dx + dy * dy) + dz * dz) + g->plane_r[j]);
        if(fabsf(cd) < g->plane_cone[j] || ahd <= g->plane_w[j] + band) {
            float entry = box_hit(g->plane_lo[j], g->plane_hi[j], r);
            if(entry >= 0.0F && (fabsf(cd) < g->plane_cone[j] || entry * fabsf(cd) < band)) return 1;
        }
    }
    for(int j = 0; j < g->n_spheres; j++) {
        float cx = r->o.x - g->sphere[j][0], cy = r->o.y - g->sphere[j][1], cz = r->o.z - g->sphere[j][2];
        float radius = g->sphere[j][3];
        float r2 = radius * radius;
        float reach = fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz));
        if(reach > radius) { /* inside the sphere's box the leaf is always reached (entry 0) and tested exactly */
            if(reach <= 1.01F * radius) return 1;
            float co2 = (cx * cx + cy * cy) + cz * cz;
            if(co2 < 9.0F * r2) {
                float dd = (r->d.x * cx + r->d.y * cy) + r->d.z * cz;
                float disc = dd * dd - co2 + r2;
                if(fabsf(disc) < r2 * 0.0625F) return 1;
            }
        }
    }
    return 0;
}

typedef struct {
    int ref;
    float entry;
} cert_entry;

static int g_guarded = 0; /* pto_set_guard: 1 = apply the guard table (PTB_FLAG_CERTIFIED_CLOSEST), 0 = relaxed */

void pto_set_guard(int guarded) { g_guarded = guarded; }

/* 1 when the guard table covers the scene (ptb_scene_info.certifiable) */
int pto_scene_certifiable(const pto_scene *s) {
    pto_guard guard;
    guard_build(s, &guard);
    return guard.enabled;
}

void pto_intersect_certified(const pto_scene *s, const float *rays, uint64_t n, uint64_t tree_seed, int shape, float *t_out, int32_t *prim_out,
                             uint8_t *certain_out) {
    const float prune_slack = 1.0078125F;   /* 1 + 2^-7 */
    const float entry_slack = 1.001953125F; /* 1 + 2^-9 */
    const float suspect_factor = 0.99609375F; /* 1 - 2^-8 */
    if(s->n_prims == 0) {
        for(uint64_t i = 0; i < n; i++) {
            t_out[i] = -1.0F;
            prim_out[i] = -1;
            certain_out[i] = 1;
        }
        return;
    }
    int np = (int)s->n_prims;
    int *items = (int *)malloc(sizeof(int) * (size_t)np);
    uint64_t state = tree_seed;
    for(int i = 0; i < np; i++) items[i] = i;
    for(int i = np - 1; i > 0; i--) {
        int j = (int)(cert_rand(&state) % (uint64_t)(i + 1));
        int tmp = items[i];
        items[i] = items[j];
        items[j] = tmp;
    }
    cnode *nodes = (cnode *)malloc(sizeof(cnode) * (size_t)(np > 1 ? np - 1 : 1));
    int n_nodes = 0;
    int root = cert_build(s, nodes, &n_nodes, items, np, shape, &state);
    free(items);
    cert_entry *stack = (cert_entry *)malloc(sizeof(cert_entry) * (size_t)(np + 1));
    pto_guard guard;
    guard_build(s, &guard);

    for(uint64_t i = 0; i < n; i++) {
        ray r = {V(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), V(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
        if(g_guarded && guard_flags_ray(&guard, &r)) { /* handed to the reference-order walk without a certified walk */
            t_out[i] = -1.0F;
            prim_out[i] = -1;
            certain_out[i] = 0;
            continue;
        }
        float rlo[3], rhi[3];
        cert_ref_box(s, nodes, root, rlo, rhi);
        float root_t = box_hit(rlo, rhi, &r);
        float best_t = FLT_MAX, prune_t = FLT_MAX, rival_t = 0.0F, hit_t = -1.0F;
        int best = -1, certain = 1, sp = 0;
        if(!(root_t >= 0.0F)) {
            t_out[i] = root_t;
            prim_out[i] = -1;
            certain_out[i] = 1;
            continue;
        }
        stack[sp].ref = root;
        stack[sp].entry = root_t;
        sp++;
        int abandoned = 0;
        while(sp > 0 && !abandoned) {
            cert_entry e = stack[--sp];
            if(!(e.entry < prune_t)) continue; /* deferred sibling re-tested against the current bound */
            int ref = e.ref;
            float entry = e.entry;
            for(;;) {
                if(ref < 0) {
                    float t = prim_hit(&s->prims[~ref], &r);
                    if(t >= 0.0F && t < entry * suspect_factor) {
                        /* hit reported in front of the primitive's own box: the walk is abandoned, the ray re-traced */
                        certain = 0;
                        abandoned = 1;
                        break;
                    }
                    if(t >= 0.0F) {
                        if(t < best_t) {
                            float rival = fmaxstd(t, entry);
                            certain = best_t > rival && entry <= t * entry_slack && t > 0.0F;
                            rival_t = rival;
                            best_t = t;
                            prune_t = t * prune_slack;
                            hit_t = t;
                            best = ~ref;
                        }
                        else if(t <= rival_t) {
                            certain = 0;
                        }
                    }
                    break;
                }
                float llo[3], lhi[3], qlo[3], qhi[3];
                cert_ref_box(s, nodes, nodes[ref].left, llo, lhi);
                cert_ref_box(s, nodes, nodes[ref].right, qlo, qhi);
                float lt = box_hit(llo, lhi, &r), rt = box_hit(qlo, qhi, &r);
                int vl = lt >= 0.0F && lt < prune_t, vr = rt >= 0.0F && rt < prune_t;
                if(vl && vr) {
                    int left_first = lt < rt;
                    stack[sp].ref = left_first ? nodes[ref].right : nodes[ref].left;
                    stack[sp].entry = left_first ? rt : lt;
                    sp++;
                    entry = left_first ? lt : rt;
                    ref = left_first ? nodes[ref].left : nodes[ref].right;
                }
                else if(vl || vr) {
                    entry = vl ? lt : rt;
                    ref = vl ? nodes[ref].left : nodes[ref].right;
                }
                else {
                    break;
                }
            }
        }
        if(g_guarded && best >= 0 && !(hit_t >= guard.tau_safe)) certain = 0;
        t_out[i] = hit_t;
        prim_out[i] = best;
        certain_out[i] = (uint8_t)certain;
    }
    free(stack);
    free(nodes);
}

void pto_aabb_intersect(const float lo[3], const float hi[3], uint64_t n, const float *rays, float *t_out) {
    for(uint64_t i = 0; i < n; i++) {
        ray r = {V(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), V(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
        t_out[i] = box_hit(lo, hi, &r);
    }
}

/* ------------------------------------------------------------------------------------------------ surfaces, lights */

/* object.cpp:126-144 and :86-88 */
static v3 prim_normal(const ptb_prim *p, v3 pos) {
    if(p->kind == PTB_PRIM_SPHERE) return vnorm(vsub(pos, V(p->p[0], p->p[1], p->p[2])));
    if(p->kind != PTB_PRIM_TRIANGLE) return V(0.0F, 1.0F, 0.0F);
    v3 a = V(p->p[0], p->p[1], p->p[2]), b = V(p->p[3], p->p[4], p->p[5]), c = V(p->p[6], p->p[7], p->p[8]);
    v3 na = V(p->p[9], p->p[10], p->p[11]), nb = V(p->p[12], p->p[13], p->p[14]), nc = V(p->p[15], p->p[16], p->p[17]);
    v3 ab = vsub(b, a), ac = vsub(c, a), ap = vsub(pos, a);
    float d00 = vdot(ab, ab), d01 = vdot(ab, ac), d11 = vdot(ac, ac), d20 = vdot(ap, ab), d21 = vdot(ap, ac);
    float inv_d = 1.0F / (d00 * d11 - d01 * d01);
    float v = (d11 * d20 - d01 * d21) * inv_d;
    float w = (d00 * d21 - d01 * d20) * inv_d;
    float u = 1.0F - v - w;
    return vnorm(vadd(vadd(vmul(na, u), vmul(nb, v)), vmul(nc, w)));
}

typedef struct {
    v3 pos;
    rgba spectrum;
    float pd;
} light_sample;

/* object.cpp:192-207 and :101-116 */
static void prim_sample(const ptb_prim *p, engine *e, v3 *pos, float *pd, int *cull) {
    if(p->kind == PTB_PRIM_TRIANGLE) {
        v3 a = V(p->p[0], p->p[1], p->p[2]), b = V(p->p[3], p->p[4], p->p[5]), c = V(p->p[6], p->p[7], p->p[8]);
        float r1 = uniform01(e);
        float r2 = uniform01(e);
        float rr1 = sqrtf(r1);
        *pos = vadd(vadd(vmul(a, 1.0F - rr1), vmul(b, rr1 * (1.0F - r2))), vmul(c, rr1 * r2));
        float area = vlen(vcross(vsub(b, a), vsub(c, a))) / 2.0F;
        *pd = 1.0F / area;
        *cull = p->cull_backface != 0;
    }
    else if(p->kind == PTB_PRIM_SPHERE) {
        const float pi = (float)M_PI;
        float radius = p->p[3], radius2 = radius * radius;
        float theta = 2.0F * pi * uniform01(e);
        float phi = acosf(1.0F - 2.0F * uniform01(e));
        float x = sinf(phi) * cosf(theta), y = sinf(phi) * sinf(theta), z = cosf(phi);
        *pos = vadd(V(p->p[0], p->p[1], p->p[2]), vmul(V(x, y, z), radius));
        *pd = 1.0F / (4.0F * pi * radius2);
        *cull = 0;
    }
    else {
        *pos = V(0, 0, 0);
        *pd = 0.0F;
        *cull = 0;
    }
}

/* scene.cpp:222-289; returns the number of samples written to out (capacity n_lights + sample count) */
static int sample_lights(const pto_scene *s, v3 pos, engine *e, light_sample *out) {
    int n_out = 0;
    int emissive_count = s->n_emissive;
    int sample_count = 2 + (int)log10((double)(emissive_count + 1));
    if(emissive_count < sample_count) sample_count = emissive_count;

    for(uint32_t i = 0; i < s->n_lights; i++) {
        const ptb_point_light *l = &s->lights[i];
        out[n_out].pos = V(l->pos[0], l->pos[1], l->pos[2]);
        memcpy(out[n_out].spectrum.c, l->rgba, sizeof(float) * 4);
        out[n_out].pd = 1.0F;
        n_out++;
    }
    for(int i = 0; i < sample_count; i++) {
        float r = uniform01(e);
        int lo = 0, len = emissive_count; /* std::lower_bound */
        while(len > 0) {
            int half = len / 2;
            if(s->cdf[lo + half] < r) {
                lo += half + 1;
                len -= half + 1;
            }
            else
                len = half;
        }
        int index = lo;
        float selection_p = s->cdf[index];
        if(index > 0) selection_p -= s->cdf[index - 1];
        selection_p *= (float)sample_count;

        const ptb_prim *p = &s->prims[s->emissive[index]];
        v3 surface_pos;
        float surface_p;
        int surface_cull;
        prim_sample(p, e, &surface_pos, &surface_p, &surface_cull);
        v3 surface_n = prim_normal(p, surface_pos);
        v3 to_light = vsub(surface_pos, pos);
        v3 dir = vnorm(to_light);
        float abs_dot = fabsf(vdot(vneg(dir), surface_n));
        if(!(abs_dot > 0.0F)) continue;
        if(!(vlen2(to_light) > 0.0F)) continue;
        if(surface_cull) {
            if(!(vdot(dir, surface_n) < 0.0F)) continue;
        }
        float conversion = vlen2(to_light) / abs_dot;
        out[n_out].pos = surface_pos;
        memcpy(out[n_out].spectrum.c, s->materials[p->material].emission, sizeof(float) * 4);
        out[n_out].pd = selection_p * surface_p * conversion;
        n_out++;
    }
    return n_out;
}

int pto_sample_lights(const pto_scene *s, const float pos[3], uint64_t seed, int max_out, float *out) {
    engine e = engine_seed(seed);
    light_sample *tmp = (light_sample *)malloc(sizeof(light_sample) * (size_t)(s->n_lights + 64));
    int n = sample_lights(s, V(pos[0], pos[1], pos[2]), &e, tmp);
    for(int i = 0; i < n && i < max_out; i++) {
        float *o = out + 8 * i;
        o[0] = tmp[i].pos.x;
        o[1] = tmp[i].pos.y;
        o[2] = tmp[i].pos.z;
        memcpy(o + 3, tmp[i].spectrum.c, sizeof(float) * 4);
        o[7] = tmp[i].pd;
    }
    free(tmp);
    return n;
}

/* ------------------------------------------------------------------------------------------------ BSDFs */

/* propagation.cpp:24-62 */
static v3 local_to_global(v3 vec, v3 n) {
    v3 d;
    if(fabsf(n.x) > 0.0F) {
        if(fabsf(n.y) > 0.0F)
            d = V(0.0F, -n.x, n.y);
        else
            d = V(0.0F, -n.x, n.z);
    }
    else {
        if(fabsf(n.y) > 0.0F)
            d = V(-n.y, n.z, 0.0F);
        else
            d = V(1.0F, 0.0F, 0.0F);
    }
    d = vnorm(d);
    v3 b1 = vnorm(vcross(d, n));
    v3 b2 = vnorm(vcross(b1, n));
    return V(vdot(V(b1.x, b2.x, n.x), vec), vdot(V(b1.y, b2.y, n.y), vec), vdot(V(b1.z, b2.z, n.z), vec));
}

/* propagation.cpp:64-83 */
static void fresnel(float ray_dot, float ri_leaving, float ri_entering, float *reflectance, float *cos_theta_t) {
    float sin_theta_i = sqrtf(fmaxstd(1.0F - ray_dot * ray_dot, 0.0F));
    float sin_theta_t = ri_leaving / ri_entering * sin_theta_i;
    if(sin_theta_t >= 1.0F) {
        *reflectance = 1.0F;
        *cos_theta_t = 0.0F;
        return;
    }
    float ct = sqrtf(fmaxstd(1.0F - sin_theta_t * sin_theta_t, 0.0F));
    float r_parallel = ((ri_entering * ray_dot) - (ri_leaving * ct)) / ((ri_entering * ray_dot) + (ri_leaving * ct));
    float r_perpendicular = ((ri_leaving * ray_dot) - (ri_entering * ct)) / ((ri_leaving * ray_dot) + (ri_entering * ct));
    *reflectance = (r_parallel * r_parallel + r_perpendicular * r_perpendicular) / 2.0F;
    *cos_theta_t = ct;
}

/* propagation.cpp:89-99, 120-160, 180-204 */
static void propagate(const ptb_material *m, ray in, v3 pos, v3 normal, float epsilon, engine *e, ray *out, float *factor, float *pd) {
    const float pi = (float)M_PI;
    if(m->bsdf == PTB_BSDF_LAMBERT) {
        /* importanceSampleCosine(dist(re), dist(re), 1.0F): g++ evaluates the arguments right to left */
        float r2 = uniform01(e);
        float r1 = uniform01(e);
        float ex = 1.0F;
        float fac = sqrtf(1.0F - powf(r2, 2.0F / (ex + 1)));
        float cos_theta = powf(r2, 1.0F / (ex + 1));
        v3 local = V(fac * cosf(2.0F * pi * r1), fac * sinf(2.0F * pi * r1), cos_theta);
        float p = (ex + 1) * powf(cos_theta, ex) / (2.0F * pi);
        v3 dir = local_to_global(local, normal);
        out->o = vadd(pos, vmul(dir, epsilon));
        out->d = dir;
        *factor = 1.0F;
        *pd = p;
        return;
    }
    if(m->bsdf == PTB_BSDF_GLASS) {
        float ray_dot = -vdot(in.d, normal);
        float ri = m->refractive_index;
        float ri_leaving = ray_dot >= 0 ? 1.0F : ri;
        float ri_entering = ray_dot >= 0 ? ri : 1.0F;
        float rat, cos_t;
        fresnel(fabsf(ray_dot), ri_leaving, ri_entering, &rat, &cos_t);
        if(bernoulli(e, (double)rat)) {
            v3 dir = vreflect(in.d, vmul(normal, ray_dot < 0.0F ? -1.0F : 1.0F));
            out->o = vadd(pos, vmul(dir, epsilon));
            out->d = dir;
            *factor = rat;
            *pd = rat;
        }
        else {
            float ri_ratio = ri_leaving / ri_entering;
            v3 dir = vadd(vmul(in.d, ri_ratio), vmul(vmul(normal, ri_ratio * fabsf(ray_dot) - cos_t), ray_dot < 0.0F ? -1.0F : 1.0F));
            dir = vnorm(dir);
            float ri_fac = (ri_entering * ri_entering) / (ri_leaving * ri_leaving);
            out->o = vadd(pos, vmul(dir, epsilon));
            out->d = dir;
            *factor = ri_fac * (1.0F - rat);
            *pd = 1.0F - rat;
        }
        return;
    }
    int unaligned = vdot(in.d, normal) > 0.0F;
    if(m->one_way && unaligned) {
        out->o = vadd(pos, vmul(in.d, epsilon));
        out->d = in.d;
        *factor = 1.0F;
        *pd = 1.0F;
        return;
    }
    v3 normal_dir = normal;
    if(!m->one_way && unaligned) normal_dir = vmul(normal_dir, -1.0F);
    v3 dir = vreflect(in.d, normal_dir);
    out->o = vadd(pos, vmul(dir, epsilon));
    out->d = dir;
    *factor = 1.0F;
    *pd = 1.0F;
}

/* propagation.cpp:101-116, 162-176, 206-217; specular colour = white (material.cpp:15-17) */
static void bsdf_spectrum(const ptb_material *m, v3 from_d, v3 to_d, v3 normal, rgba light, int synthetic, rgba *spectrum, float *shade, float *pd) {
    const float pi = (float)M_PI;
    rgba white = {{1.0F, 1.0F, 1.0F, 1.0F}};
    rgba diffuse;
    memcpy(diffuse.c, m->diffuse, sizeof(float) * 4);
    if(m->bsdf == PTB_BSDF_LAMBERT) {
        *shade = fmaxstd(vdot(normal, to_d), 0.0F) / pi;
        *spectrum = cmul(diffuse, light);
        *pd = 1.0F;
        return;
    }
    *shade = 1.0F;
    *pd = synthetic ? 0.0F : 1.0F;
    if(m->bsdf == PTB_BSDF_GLASS) {
        *spectrum = vdot(from_d, to_d) <= 0.0F ? cmul(light, white) : cmul(light, diffuse);
        return;
    }
    *spectrum = light;
    if(!m->one_way || vdot(from_d, to_d) <= 0.0F) *spectrum = cmul(light, white);
}

/* ------------------------------------------------------------------------------------------------ camera */

/* camera.cpp:7-49 */
static void aperture_sample(const ptb_camera *c, engine *e, float *sx, float *sy) {
    if(c->aperture_kind == PTB_APERTURE_CIRCULAR) {
        const float pi = (float)M_PI;
        float r = sqrtf(uniform01(e));
        float theta = 2 * pi * uniform01(e);
        *sx = r * cosf(theta);
        *sy = r * sinf(theta);
        return;
    }
    float ratio = c->hexagon_horizontal_ratio;
    float x, y;
    int inside;
    do {
        x = uniform01(e);
        y = uniform01(e);
        float relative_x = x - ratio;
        inside = (relative_x <= 0.0F) || (relative_x / (1.0F - ratio)) >= y;
    } while(!inside);
    if(bernoulli(e, 0.5)) x = -x;
    if(bernoulli(e, 0.5)) y = -y;
    *sx = x;
    *sy = y;
}

/* camera.cpp:78-113 */
static ray shoot_ray(const ptb_camera *c, float x, float y, float pixel_width, float pixel_height, engine *e) {
    float offset_x = uniform_f(e, -pixel_width / 2.0F, pixel_width / 2.0F);
    float offset_y = uniform_f(e, -pixel_height / 2.0F, pixel_height / 2.0F);
    float sensor_x = x + offset_x;
    float sensor_y = y + offset_y;
    v3 origin = V(c->origin[0], c->origin[1], c->origin[2]);
    v3 forward = V(c->forward[0], c->forward[1], c->forward[2]);
    v3 up = V(c->up[0], c->up[1], c->up[2]);
    v3 right = V(c->right[0], c->right[1], c->right[2]);
    v3 sensor_pos = vsub(vsub(vsub(origin, forward), vmul(up, sensor_y)), vmul(right, sensor_x));
    float ax = 0.0F, ay = 0.0F;
    if(c->aperture_kind != PTB_APERTURE_NONE) {
        float sx, sy;
        aperture_sample(c, e, &sx, &sy);
        ax = sx * c->aperture_width_half;
        ay = sy * c->aperture_height_half;
    }
    ray r;
    r.o = vadd(vadd(origin, vmul(up, ax)), vmul(right, ay));
    if(c->focal_plane_dist > 0.0F) {
        v3 base_dir = vnorm(vsub(origin, sensor_pos));
        v3 target = vadd(origin, vmul(base_dir, c->focal_plane_dist / vdot(forward, base_dir)));
        r.d = vnorm(vsub(target, r.o));
    }
    else {
        r.d = vnorm(vsub(r.o, sensor_pos));
    }
    return r;
}

void pto_camera_shoot(const ptb_camera *c, uint64_t n, const float *xy, float pw, float ph, const uint64_t *seeds, float *out) {
    for(uint64_t i = 0; i < n; i++) {
        engine e = engine_seed(seeds[i]);
        ray r = shoot_ray(c, xy[2 * i], xy[2 * i + 1], pw, ph, &e);
        out[6 * i] = r.o.x;
        out[6 * i + 1] = r.o.y;
        out[6 * i + 2] = r.o.z;
        out[6 * i + 3] = r.d.x;
        out[6 * i + 4] = r.d.y;
        out[6 * i + 5] = r.d.z;
    }
}

/* ------------------------------------------------------------------------------------------------ getSample */

static float contribution(rgba c) { return (c.c[0] + c.c[1] + c.c[2]) / 3.0F; }

typedef struct {
    uint64_t closest_rays, shadow_rays, vertices, samples;
} pto_counters;

/* worker.cpp:26-146 */
static rgba get_sample(const pto_scene *s, const ptb_camera *cam, int width, int height, float epsilon, int max_depth, float x_camera, float y_camera,
                       engine *e, int *collected, pto_counters *counters) {
    float pixel_width = 1.0F / (float)width;
    float pixel_height = 1.0F / (float)height;
    ray r = shoot_ray(cam, x_camera, y_camera, pixel_width, pixel_height, e);
    int sample_collected = 0;
    float contribution_unweighted = 1.0F;
    double sample_divisor = 1.0F;
    double sample_bounce_pd = 1.0;
    rgba sample_spectrum = {{1.0F, 1.0F, 1.0F, 1.0F}};
    rgba out_spectrum = czero();
    int path_length = 0;
    light_sample *lights = (light_sample *)malloc(sizeof(light_sample) * (size_t)(s->n_lights + 64));
    if(counters) counters->samples++;

    for(;;) {
        hit h = scene_hit(s, &r);
        if(counters) counters->closest_rays++;
        if(h.t < 0.0F) break;
        path_length++;
        sample_collected = 1;
        if(counters) counters->vertices++;

        v3 pos = vadd(r.o, vmul(r.d, h.t));
        const ptb_prim *object = &s->prims[h.prim];
        v3 n = prim_normal(object, pos);
        const ptb_material *material = &s->materials[object->material];
        rgba emission;
        memcpy(emission.c, material->emission, sizeof(float) * 4);
        out_spectrum = cadd(out_spectrum, cdiv(cmul(sample_spectrum, emission), (float)(sample_divisor * sample_bounce_pd)));

        float bounce_probability =
          path_length <= 4 ? 1.0F : 0.1F + 0.1F * fminstd(contribution_unweighted * contribution(sample_spectrum), 1.0F);
        int do_bounce = uniform01(e) < bounce_probability;

        int n_lights = sample_lights(s, pos, e, lights);
        for(int i = 0; i < n_lights; i++) {
            v3 to_light = vsub(lights[i].pos, pos);
            v3 light_dir = vnorm(to_light);
            ray light_ray = {vadd(pos, vmul(light_dir, epsilon)), light_dir};
            float light_t = scene_hit(s, &light_ray).t;
            if(counters) counters->shadow_rays++;
            if(light_t < 0.0F || (light_t >= vlen(to_light) - epsilon)) {
                rgba base;
                float shading_factor, shadow_ray_pd;
                bsdf_spectrum(material, r.d, light_ray.d, n, lights[i].spectrum, 1, &base, &shading_factor, &shadow_ray_pd);
                if(shadow_ray_pd > 0.0F) {
                    rgba combined = cmul(cscale(base, shading_factor), sample_spectrum);
                    rgba weighed = cdiv(combined, (float)(sample_divisor * sample_bounce_pd * lights[i].pd * shadow_ray_pd));
                    out_spectrum = cadd(out_spectrum, weighed);
                }
            }
        }

        if(!do_bounce) {
            sample_bounce_pd *= 1.0F - bounce_probability;
            break;
        }
        sample_bounce_pd *= bounce_probability;
        if(sample_bounce_pd <= 1E-20) break;
        if(max_depth > 0 && path_length >= max_depth) break; /* extension, off (0) for reference behaviour */

        ray next;
        float ray_factor, ray_pd;
        propagate(material, r, pos, n, epsilon, e, &next, &ray_factor, &ray_pd);
        sample_divisor *= ray_pd;
        sample_divisor /= ray_factor;
        contribution_unweighted *= ray_factor;

        rgba shaded;
        float shading_factor, shading_pd;
        bsdf_spectrum(material, r.d, next.d, n, sample_spectrum, 0, &shaded, &shading_factor, &shading_pd);
        sample_divisor *= shading_pd;
        sample_divisor /= shading_factor;
        contribution_unweighted *= shading_factor;
        sample_spectrum = shaded;
        if(sample_divisor <= 1E-20) break;
        r = next;
    }
    free(lights);
    out_spectrum.c[3] = sample_collected ? 1.0F : 0.0F;
    *collected = sample_collected;
    return out_spectrum;
}

static void pixel_to_camera(int x, int y, int width, int height, float *xc, float *yc) {
    const float one_half = 1.0F / 2.0F;
    *xc = 2 * (((float)x + one_half) / (float)width - one_half);
    *yc = 2 * (((float)y + one_half) / (float)height - one_half);
    *yc = -*yc;
}

/* one getSample per (pixel, seed): what processItem computes for a 1x1 item at 1 spp with RandomEngine(seed) */
void pto_render_samples(const pto_scene *s, const ptb_camera *cam, int width, int height, float epsilon, int max_depth, uint64_t n, const int32_t *pixels,
                        const uint64_t *seeds, float *out_rgba, uint64_t *counters_out) {
    pto_counters counters = {0, 0, 0, 0};
    for(uint64_t i = 0; i < n; i++) {
        engine e = engine_seed(seeds[i]);
        float xc, yc;
        pixel_to_camera(pixels[2 * i], pixels[2 * i + 1], width, height, &xc, &yc);
        int collected;
        rgba c = get_sample(s, cam, width, height, epsilon, max_depth, xc, yc, &e, &collected, &counters);
        if(!collected) c = czero(); /* pixel_value stays 0 when nothing was collected (worker.cpp:196, 263) */
        memcpy(out_rgba + 4 * i, c.c, sizeof(float) * 4);
    }
    if(counters_out) {
        counters_out[0] = counters.samples;
        counters_out[1] = counters.closest_rays;
        counters_out[2] = counters.shadow_rays;
        counters_out[3] = counters.vertices;
    }
}

/* ------------------------------------------------------------------------------------------------ processItem */

#define PTO_MAX_CANDIDATES 64

typedef struct {
    int min_samples, max_samples;
    int stats_sample_count, candidate_batch_count, check_sample_count;
    rgba pixel_value;
    int collected_sample_count;
    rgba contribution_mean, contribution_m2;
    int contribution_count;
    int stats_sample_index;
    rgba sample_aggregate;
    rgba candidate_means[PTO_MAX_CANDIDATES], candidate_m2s[PTO_MAX_CANDIDATES];
    int candidate_counts[PTO_MAX_CANDIDATES];
    int n_candidates;
    rgba candidate_mean, candidate_m2;
    int candidate_count;
    int remaining_checks;
    int accepted_candidate;
} pixel_state;

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* worker.cpp:158-191 */
static void pixel_begin(pixel_state *p, int min_samples, int max_samples) {
    memset(p, 0, sizeof(*p));
    p->min_samples = min_samples;
    p->max_samples = max_samples;
    p->stats_sample_count = imin(imax(min_samples / 4, 1), 64);
    p->candidate_batch_count = imax(imax(min_samples, max_samples / 4) / p->stats_sample_count, 2);
    p->check_sample_count =
      imin(imax(imax(min_samples / 2, (max_samples - min_samples) / 8), imax(8, p->stats_sample_count)), 1024) / p->stats_sample_count;
    p->remaining_checks = p->check_sample_count;
}

/* worker.cpp:196-260; returns 1 when sampling of this pixel stops early */
static int pixel_add(pixel_state *p, rgba color) {
    p->contribution_count++;
    p->stats_sample_index++;
    p->sample_aggregate = cadd(p->sample_aggregate, color);
    if(p->stats_sample_index == p->stats_sample_count) {
        p->sample_aggregate = cdiv(p->sample_aggregate, (float)p->stats_sample_count);
        rgba delta = csub(p->sample_aggregate, p->contribution_mean);
        p->contribution_mean = cadd(p->contribution_mean, cdiv(delta, (float)(p->contribution_count / p->stats_sample_count)));
        rgba delta2 = csub(p->sample_aggregate, p->contribution_mean);
        p->contribution_m2 = cadd(p->contribution_m2, cmul(delta, delta2));

        if(p->candidate_count == p->candidate_batch_count) {
            if(p->n_candidates < PTO_MAX_CANDIDATES) {
                p->candidate_means[p->n_candidates] = p->candidate_mean;
                p->candidate_m2s[p->n_candidates] = p->candidate_m2;
                p->candidate_counts[p->n_candidates] = p->candidate_count;
                p->n_candidates++;
            }
            p->candidate_mean = czero();
            p->candidate_m2 = czero();
            p->candidate_count = 0;
        }
        p->candidate_count++;
        rgba cd = csub(p->sample_aggregate, p->candidate_mean);
        p->candidate_mean = cadd(p->candidate_mean, cdiv(cd, (float)p->candidate_count));
        rgba cd2 = csub(p->sample_aggregate, p->candidate_mean);
        p->candidate_m2 = cadd(p->candidate_m2, cmul(cd, cd2));

        p->stats_sample_index = 0;
        p->sample_aggregate = czero();
    }
    p->pixel_value = cadd(p->pixel_value, color);
    p->collected_sample_count++;

    if(p->stats_sample_index == 0 && p->collected_sample_count >= imax(p->min_samples, 2)) {
        int passed_check = 0;
        if(p->contribution_count / p->stats_sample_count >= 2) {
            rgba m2_weighted = cdiv(p->contribution_m2, (float)(p->contribution_count / p->stats_sample_count - 1));
            float stddev = sqrtf(m2_weighted.c[0] + m2_weighted.c[1] + m2_weighted.c[2]);
            if(stddev < 1E-4F || stddev / (3 * 3 * contribution(p->contribution_mean) + 1E-5) < 0.2F) {
                passed_check = 1;
                p->remaining_checks--;
                if(p->remaining_checks <= 0) {
                    p->accepted_candidate = 1;
                    return 1;
                }
            }
        }
        if(!passed_check) p->remaining_checks = p->check_sample_count;
    }
    return 0;
}

/* worker.cpp:263-317 */
static rgba pixel_finish(pixel_state *p) {
    rgba pixel_value = p->pixel_value;
    if(p->collected_sample_count > 0) pixel_value = cscale(pixel_value, 1.0F / (float)p->collected_sample_count);
    if(p->candidate_count > 0 && p->n_candidates < PTO_MAX_CANDIDATES) {
        p->candidate_means[p->n_candidates] = p->candidate_mean;
        p->candidate_m2s[p->n_candidates] = p->candidate_m2;
        p->candidate_counts[p->n_candidates] = p->candidate_count;
        p->n_candidates++;
    }
    if(!p->accepted_candidate) {
        rgba colors[PTO_MAX_CANDIDATES];
        float stddevs[PTO_MAX_CANDIDATES];
        int n = 0;
        for(int i = 0; i < p->n_candidates; i++) {
            if(p->candidate_counts[i] < imax((p->candidate_batch_count * 3) / 4, 2)) continue;
            rgba m2_weighted = cdiv(p->candidate_m2s[i], (float)p->candidate_counts[i]);
            colors[n] = p->candidate_means[i];
            stddevs[n] = sqrtf(m2_weighted.c[0] + m2_weighted.c[1] + m2_weighted.c[2]);
            n++;
        }
        if(n > 0) {
            /* std::sort by stddev; libstdc++ uses a (stable) insertion sort below 16 elements */
            for(int i = 1; i < n; i++) {
                rgba c = colors[i];
                float sd = stddevs[i];
                int j = i;
                while(j > 0 && sd < stddevs[j - 1]) {
                    colors[j] = colors[j - 1];
                    stddevs[j] = stddevs[j - 1];
                    j--;
                }
                colors[j] = c;
                stddevs[j] = sd;
            }
            pixel_value = colors[0];
            float stddev = stddevs[0];
            for(int i = 1; i < n; i++) {
                float other = stddevs[i];
                if(other < fmaxstd(stddev + 0.005F, stddev * 1.01F)) {
                    pixel_value = cadd(pixel_value, cdiv(csub(colors[i], pixel_value), (float)(i + 1)));
                    stddev = other;
                }
                else
                    break;
            }
        }
    }
    return pixel_value;
}

/* worker.cpp:149-326 with one sequential engine, exactly as the reference */
void pto_process_item(const pto_scene *s, const ptb_camera *cam, int width, int height, int min_samples, int max_samples, float epsilon, int offset_x,
                      int offset_y, int tile_width, int tile_height, uint64_t seed, float *out_rgba) {
    engine e = engine_seed(seed);
    for(int y = offset_y; y < offset_y + tile_height; y++) {
        for(int x = offset_x; x < offset_x + tile_width; x++) {
            float xc, yc;
            pixel_to_camera(x, y, width, height, &xc, &yc);
            pixel_state p;
            pixel_begin(&p, min_samples, max_samples);
            for(int k = 0; k < max_samples; k++) {
                int collected;
                rgba c = get_sample(s, cam, width, height, epsilon, 0, xc, yc, &e, &collected, NULL);
                if(collected) {
                    if(pixel_add(&p, c)) break;
                }
            }
            rgba v = pixel_finish(&p);
            memcpy(out_rgba + 4 * ((size_t)(y - offset_y) * (size_t)tile_width + (size_t)(x - offset_x)), v.c, sizeof(float) * 4);
        }
    }
}

/* processItem's per-pixel statistics over precomputed samples: samples[k * n_pixels + q], alpha = collected flag.
 * Used to check the device's resolve kernel on identical per-sample inputs. */
/* consumed_out (may be NULL): how many samples the loop of worker.cpp:172-260 drew for the pixel before it ended */
void pto_resolve_counts(int min_samples, int max_samples, uint32_t n_pixels, const float *samples, float *out_rgba, int32_t *consumed_out) {
    for(uint32_t q = 0; q < n_pixels; q++) {
        pixel_state p;
        pixel_begin(&p, min_samples, max_samples);
        int k = 0;
        while(k < max_samples) {
            const float *sm = samples + 4 * ((size_t)k * n_pixels + q);
            k++;
            if(sm[3] == 0.0F) continue;
            rgba c = {{sm[0], sm[1], sm[2], 1.0F}};
            if(pixel_add(&p, c)) break;
        }
        rgba v = pixel_finish(&p);
        memcpy(out_rgba + 4 * (size_t)q, v.c, sizeof(float) * 4);
        if(consumed_out != NULL) consumed_out[q] = k;
    }
}

void pto_resolve(int min_samples, int max_samples, uint32_t n_pixels, const float *samples, float *out_rgba) {
    pto_resolve_counts(min_samples, max_samples, n_pixels, samples, out_rgba, NULL);
}

/* ------------------------------------------------------------------------------------------------ unit entries
 * The same one-element-at-a-time functions the C-ABI exposes (ptb_prim_normal / ptb_prim_sample / ptb_bsdf_propagate /
 * ptb_bsdf_spectrum), with raw engine states in and out, for value parity of those entry points. */

void pto_prim_normal(const ptb_prim *prim, uint64_t n, const float *positions, float *out) {
    for(uint64_t i = 0; i < n; i++) {
        v3 nn = prim_normal(prim, V(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]));
        out[3 * i] = nn.x;
        out[3 * i + 1] = nn.y;
        out[3 * i + 2] = nn.z;
    }
}

void pto_prim_sample(const ptb_prim *prim, uint64_t n, uint64_t *states, float *out) {
    for(uint64_t i = 0; i < n; i++) {
        engine e;
        e.s = states[i];
        v3 pos;
        float pd;
        int cull;
        prim_sample(prim, &e, &pos, &pd, &cull);
        states[i] = e.s;
        out[5 * i] = pos.x;
        out[5 * i + 1] = pos.y;
        out[5 * i + 2] = pos.z;
        out[5 * i + 3] = pd;
        out[5 * i + 4] = cull ? 1.0F : 0.0F;
    }
}

/* in: 9 floats (incoming direction, position, normal); out: 8 floats (origin, direction, factor, density) */
void pto_bsdf_propagate(const ptb_material *m, float epsilon, uint64_t n, const float *in, uint64_t *states, float *out) {
    for(uint64_t i = 0; i < n; i++) {
        const float *p = in + 9 * i;
        engine e;
        e.s = states[i];
        ray rin = {V(0.0F, 0.0F, 0.0F), V(p[0], p[1], p[2])};
        ray rout;
        float factor, pd;
        propagate(m, rin, V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), epsilon, &e, &rout, &factor, &pd);
        states[i] = e.s;
        float *w = out + 8 * i;
        w[0] = rout.o.x;
        w[1] = rout.o.y;
        w[2] = rout.o.z;
        w[3] = rout.d.x;
        w[4] = rout.d.y;
        w[5] = rout.d.z;
        w[6] = factor;
        w[7] = pd;
    }
}

/* in: 13 floats (from-camera direction, to-light direction, normal, light rgba); out: 6 floats (rgba, shade, density) */
void pto_bsdf_spectrum(const ptb_material *m, int synthetic, uint64_t n, const float *in, float *out) {
    for(uint64_t i = 0; i < n; i++) {
        const float *p = in + 13 * i;
        rgba light = {{p[9], p[10], p[11], p[12]}};
        rgba spectrum;
        float shade, pd;
        bsdf_spectrum(m, V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), light, synthetic, &spectrum, &shade, &pd);
        float *w = out + 6 * i;
        for(int c = 0; c < 4; c++) w[c] = spectrum.c[c];
        w[4] = shade;
        w[5] = pd;
    }
}

int pto_version(void) { return 2; }
