/*
 * traverse.cuh -- device-side octree search for sm_100a.
 *
 * Restates, per thread, the reference's recursive first-hit ray search
 * (src/voxtrees/search.c:8-153) and any-hit ball search (search.c:155-183)
 * over the flattened device tree, with the reference's scalar geometry
 * (src/voxtrees/geom.c:13-154).  Every fp32 operation is performed in the
 * reference's order with round-to-nearest and without contraction (the
 * translation unit is compiled with -fmad=false; divisions are IEEE), so leaf
 * choice and hit points are bit-identical to the scalar reference build
 * compiled with -ffp-contract=off.
 *
 * Device tree layout (built by voxgpu_tree_upload):
 *   nodes[3*id+0] = { bb_min.xyz , meta }   meta: bits 0-1 kind (0 inner, 1 leaf, 2 dense leaf)
 *                                                  inner: bits 8-15 child mask
 *                                                  leaf : bits 2-31 number of dots
 *   nodes[3*id+1] = { bb_max.xyz , link }   inner: id of first non-NULL child (children are
 *                                                  consecutive, in subspace order)
 *                                            leaf : index of first dot in dots[]
 *   nodes[3*id+2] = { center.xyz , 0    }   inner only
 *   dots[i]       = { voxel min corner xyz, 0 }
 * Records are 48 B, 16-byte aligned: a node visit is two or three LDG.128.
 *
 * The recursion is turned into a stack of pending calls.  search.c returns as
 * soon as any recursive call hits, so a miss only ever means "go on with the
 * next pending call" and nothing has to be returned upwards.  An inner node
 * pushes one EXPAND entry (itself + its entry point) before descending into
 * the first child; only if that subtree misses is the entry popped and the
 * dividing planes evaluated (search.c:105-148), exactly when the reference
 * evaluates them.  The plane children are then pushed as plain calls
 * (child id, plane hit point) in reverse order.
 */
#ifndef VOXGPU_TRAVERSE_CUH
#define VOXGPU_TRAVERSE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define VOXGPU_KIND_INNER 0u
#define VOXGPU_KIND_LEAF  1u
#define VOXGPU_KIND_DENSE 2u

#define VOXGPU_MAX_DEPTH 32                      /* root alone = 1 */
#define VOXGPU_RAY_STACK (2*VOXGPU_MAX_DEPTH)    /* <= 2 pending entries per level */
#define VOXGPU_BALL_STACK (7*VOXGPU_MAX_DEPTH+8) /* <= 7 pending siblings per level */
#define VOXGPU_EXPAND_BIT 0x80000000u

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3 (float4 q) { V3 r; r.x = q.x; r.y = q.y; r.z = q.z; return r; }
__device__ __forceinline__ float comp (const V3 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct TreeRef
{
    const float4 *nodes;
    const float4 *dots;
    float vx, vy, vz;       /* vox_voxel captured at upload */
};

struct RayResult
{
    int leaf;               /* breadth-first id, -1 = miss */
    int voxel;              /* index into dots[], -1 = dense leaf / miss */
    V3 point;
};

struct RayCounters
{
    unsigned calls, inner, leaf, dense, dots, expansions, max_stack;
};

/* geom.c:13-31 fit_into_box + geom.c:69-99 hit_box */
__device__ __forceinline__ bool ray_box (const V3 mn, const V3 mx, const V3 o, const V3 d, V3 &res)
{
    const bool lx = o.x < mn.x, hx = o.x > mx.x;
    const bool ly = o.y < mn.y, hy = o.y > mx.y;
    const bool lz = o.z < mn.z, hz = o.z > mx.z;
    float cx = lx ? mn.x : (hx ? mx.x : o.x);
    float cy = ly ? mn.y : (hy ? mx.y : o.y);
    float cz = lz ? mn.z : (hz ? mx.z : o.z);
    if (!(lx | hx | ly | hy | lz | hz)) { res = o; return true; }

    int axis = 0;
    float tmax = -1.0f;
    float t = __fdiv_rn (__fsub_rn (cx, o.x), d.x);
    if (t > tmax) { axis = 0; tmax = t; }
    t = __fdiv_rn (__fsub_rn (cy, o.y), d.y);
    if (t > tmax) { axis = 1; tmax = t; }
    t = __fdiv_rn (__fsub_rn (cz, o.z), d.z);
    if (t > tmax) { axis = 2; tmax = t; }
    if (tmax <= 0.0f) return false;

    if (axis != 0)
    {
        const float c = __fadd_rn (o.x, __fmul_rn (tmax, d.x));
        if (c < mn.x || c > mx.x) return false;
        cx = c;
    }
    if (axis != 1)
    {
        const float c = __fadd_rn (o.y, __fmul_rn (tmax, d.y));
        if (c < mn.y || c > mx.y) return false;
        cy = c;
    }
    if (axis != 2)
    {
        const float c = __fadd_rn (o.z, __fmul_rn (tmax, d.z));
        if (c < mn.z || c > mx.z) return false;
        cz = c;
    }
    res.x = cx; res.y = cy; res.z = cz;
    return true;
}

/* geom.c:125-131 vox_abs_metric: ((0+|dx|)+|dy|)+|dz| */
__device__ __forceinline__ float l1_metric (const V3 a, const V3 b)
{
    float r = fabsf (__fsub_rn (a.x, b.x));
    r = __fadd_rn (r, fabsf (__fsub_rn (a.y, b.y)));
    r = __fadd_rn (r, fabsf (__fsub_rn (a.z, b.z)));
    return r;
}

/* geom.c:101-123 hit_plane_within_box */
__device__ __forceinline__ bool ray_plane_in_box (const V3 o, const V3 d, float plane, int axis,
                                                  const V3 mn, const V3 mx, V3 &res)
{
    const float oa = comp (o, axis), da = comp (d, axis);
    float k = __fsub_rn (plane, oa);
    if (k == 0.0f || ((da < 0.0f) != (k < 0.0f))) return false;
    k = __fdiv_rn (k, da);
    res.x = __fadd_rn (o.x, __fmul_rn (k, d.x));
    if (res.x < mn.x || res.x > mx.x) return false;
    res.y = __fadd_rn (o.y, __fmul_rn (k, d.y));
    if (res.y < mn.y || res.y > mx.y) return false;
    res.z = __fadd_rn (o.z, __fmul_rn (k, d.z));
    if (res.z < mn.z || res.z > mx.z) return false;
    return true;
}

/* geom.c:53-65 get_corrected_subspace_idx */
__device__ __forceinline__ unsigned subspace_corrected (const V3 c, const V3 p, const V3 d)
{
    const unsigned bx = (c.x == p.x) ? (d.x < 0.0f) : (c.x > p.x);
    const unsigned by = (c.y == p.y) ? (d.y < 0.0f) : (c.y > p.y);
    const unsigned bz = (c.z == p.z) ? (d.z < 0.0f) : (c.z > p.z);
    return bx | (by << 1) | (bz << 2);
}

__device__ __forceinline__ unsigned child_id (unsigned first, unsigned mask, unsigned sub)
{
    return first + __popc (mask & ((1u << sub) - 1u));
}

/*
 * vox_ray_tree_intersection (search.c:8-153) from node `start` (start < 0x80000000).
 * `stack` is VOXGPU_RAY_STACK float4 entries of per-thread scratch.
 */
template <bool STATS>
__device__ __forceinline__ void ray_tree (const TreeRef &tr, unsigned start, const V3 origin, const V3 d,
                                          float4 *stack, RayResult &out, RayCounters *cnt)
{
    int sp = 0;
    unsigned cur = start;
    V3 org = origin;
    bool have_call = true;

    out.leaf = -1;
    out.voxel = -1;

    for (;;)
    {
        if (!have_call)
        {
            /* the previous call missed: take the next pending entry */
            if (sp == 0) return;
            const float4 e = stack[--sp];
            const unsigned tag = __float_as_uint (e.w);
            org.x = e.x; org.y = e.y; org.z = e.z;
            if (!(tag & VOXGPU_EXPAND_BIT)) { cur = tag; have_call = true; }
            else
            {
                /* search.c:99-148: first child missed; evaluate the dividing planes */
                const unsigned id = tag & ~VOXGPU_EXPAND_BIT;
                const float4 q0 = __ldg (tr.nodes + 3*(size_t)id);
                const float4 q1 = __ldg (tr.nodes + 3*(size_t)id + 1);
                const float4 q2 = __ldg (tr.nodes + 3*(size_t)id + 2);
                const V3 mn = v3 (q0), mx = v3 (q1), c = v3 (q2);
                const unsigned mask = (__float_as_uint (q0.w) >> 8) & 0xffu;
                const unsigned first = __float_as_uint (q1.w);
                if (STATS) cnt->expansions++;

                V3 p[3];
                float m[3];
                int ax[3];
                int n = 0;
                {
                    V3 r;
                    if (ray_plane_in_box (org, d, c.x, 0, mn, mx, r)) { p[n] = r; ax[n] = 0; m[n] = l1_metric (org, r); n++; }
                    if (ray_plane_in_box (org, d, c.y, 1, mn, mx, r)) { p[n] = r; ax[n] = 1; m[n] = l1_metric (org, r); n++; }
                    if (ray_plane_in_box (org, d, c.z, 2, mn, mx, r)) { p[n] = r; ax[n] = 2; m[n] = l1_metric (org, r); n++; }
                }
                /* search.c:115-134: selection sort, strict <, swap semantics */
#define VOXGPU_SWAP(i, j) do { V3 tp = p[i]; p[i] = p[j]; p[j] = tp;            \
                               float tm = m[i]; m[i] = m[j]; m[j] = tm;        \
                               int ta = ax[i]; ax[i] = ax[j]; ax[j] = ta; } while (0)
                if (n >= 2 && m[1] < m[0]) VOXGPU_SWAP (0, 1);
                if (n == 3)
                {
                    if (m[2] < m[0]) VOXGPU_SWAP (0, 2);
                    if (m[2] < m[1]) VOXGPU_SWAP (1, 2);
                }
#undef VOXGPU_SWAP
                /* search.c:137: the subspace flips accumulate */
                unsigned sub = subspace_corrected (c, org, d);
                unsigned s[3];
                for (int i=0; i<3; i++) { if (i < n) sub ^= 1u << ax[i]; s[i] = sub; }
                /* push later planes first so that the nearest is called first */
                for (int i=2; i>=0; i--)
                {
                    if (i < n && (mask & (1u << s[i])))
                    {
                        float4 ne;
                        ne.x = p[i].x; ne.y = p[i].y; ne.z = p[i].z;
                        ne.w = __uint_as_float (child_id (first, mask, s[i]));
                        stack[sp++] = ne;
                    }
                }
                if (STATS) cnt->max_stack = max (cnt->max_stack, (unsigned) sp);
            }
            continue;
        }

        /* ---- one call: vox_ray_tree_intersection (cur, org, d) ---- */
        have_call = false;
        const float4 q0 = __ldg (tr.nodes + 3*(size_t)cur);
        const float4 q1 = __ldg (tr.nodes + 3*(size_t)cur + 1);
        const unsigned meta = __float_as_uint (q0.w);
        const unsigned link = __float_as_uint (q1.w);
        const unsigned kind = meta & 3u;
        if (STATS) cnt->calls++;

        V3 entry;
        if (!ray_box (v3 (q0), v3 (q1), org, d, entry)) continue;   /* search.c:22-27 */

        if (kind == VOXGPU_KIND_DENSE)
        {
            /* search.c:31-37 */
            if (STATS) cnt->dense++;
            out.leaf = (int) cur; out.voxel = -1; out.point = entry;
            return;
        }
        if (kind == VOXGPU_KIND_LEAF)
        {
            /* search.c:39-81 */
            const unsigned n = meta >> 2;
            float closest = __int_as_float (0x7f800000);
            if (STATS) { cnt->leaf++; }
            for (unsigned i=0; i<n; i++)
            {
                const float4 dq = __ldg (tr.dots + (size_t)link + i);
                V3 vmin = v3 (dq), vmax, far;
                vmax.x = __fadd_rn (vmin.x, tr.vx);
                vmax.y = __fadd_rn (vmin.y, tr.vy);
                vmax.z = __fadd_rn (vmin.z, tr.vz);
                if (STATS) cnt->dots++;
                if (ray_box (vmin, vmax, entry, d, far))
                {
                    const float dist = l1_metric (entry, far);
                    if (dist == 0.0f)
                    {
                        out.leaf = (int) cur; out.voxel = (int)(link + i); out.point = far;
                        return;
                    }
                    if (dist < closest)
                    {
                        closest = dist;
                        out.leaf = (int) cur; out.voxel = (int)(link + i); out.point = far;
                    }
                }
            }
            if (out.leaf >= 0) return;
            continue;
        }

        /* inner node, search.c:83-97 */
        {
            if (STATS) cnt->inner++;
            const float4 q2 = __ldg (tr.nodes + 3*(size_t)cur + 2);
            const unsigned mask = (meta >> 8) & 0xffu;
            const unsigned sub = subspace_corrected (v3 (q2), entry, d);
            float4 e;
            e.x = entry.x; e.y = entry.y; e.z = entry.z;
            e.w = __uint_as_float (cur | VOXGPU_EXPAND_BIT);
            stack[sp++] = e;
            if (STATS) cnt->max_stack = max (cnt->max_stack, (unsigned) sp);
            if (mask & (1u << sub))
            {
                cur = child_id (link, mask, sub);
                org = entry;
                have_call = true;
            }
        }
    }
}

/* geom.c:149-154 box_ball_interp with geom.c:133-139 vox_sqr_metric */
__device__ __forceinline__ bool box_ball (const V3 mn, const V3 mx, const V3 c, float rr)
{
    const float fx = c.x < mn.x ? mn.x : (c.x > mx.x ? mx.x : c.x);
    const float fy = c.y < mn.y ? mn.y : (c.y > mx.y ? mx.y : c.y);
    const float fz = c.z < mn.z ? mn.z : (c.z > mx.z ? mx.z : c.z);
    const float dx = __fsub_rn (fx, c.x), dy = __fsub_rn (fy, c.y), dz = __fsub_rn (fz, c.z);
    float s = __fmul_rn (dx, dx);
    s = __fadd_rn (s, __fmul_rn (dy, dy));
    s = __fadd_rn (s, __fmul_rn (dz, dz));
    return s < rr;
}

/*
 * vox_tree_ball_collidep (search.c:155-183).  The result is a pure "exists",
 * so the visiting order is free; children are tested before being pushed
 * (their records are consecutive) so only intersecting nodes reach the stack.
 */
__device__ __forceinline__ bool ball_tree (const TreeRef &tr, unsigned start, const V3 c, float radius,
                                           unsigned *stack)
{
    const float rr = __fmul_rn (radius, radius);
    int sp = 0;
    {
        const float4 q0 = __ldg (tr.nodes + 3*(size_t)start);
        const float4 q1 = __ldg (tr.nodes + 3*(size_t)start + 1);
        if (!box_ball (v3 (q0), v3 (q1), c, rr)) return false;
        stack[sp++] = start;
    }
    while (sp > 0)
    {
        const unsigned id = stack[--sp];
        const float4 q0 = __ldg (tr.nodes + 3*(size_t)id);
        const float4 q1 = __ldg (tr.nodes + 3*(size_t)id + 1);
        const unsigned meta = __float_as_uint (q0.w);
        const unsigned link = __float_as_uint (q1.w);
        const unsigned kind = meta & 3u;
        if (kind == VOXGPU_KIND_DENSE) return true;
        if (kind == VOXGPU_KIND_LEAF)
        {
            const unsigned n = meta >> 2;
            for (unsigned i=0; i<n; i++)
            {
                const float4 dq = __ldg (tr.dots + (size_t)link + i);
                V3 vmin = v3 (dq), vmax;
                vmax.x = __fadd_rn (vmin.x, tr.vx);
                vmax.y = __fadd_rn (vmin.y, tr.vy);
                vmax.z = __fadd_rn (vmin.z, tr.vz);
                if (box_ball (vmin, vmax, c, rr)) return true;
            }
            continue;
        }
        const int nchild = __popc ((meta >> 8) & 0xffu);
        for (int j=nchild-1; j>=0; j--)
        {
            const unsigned ch = link + j;
            const float4 c0 = __ldg (tr.nodes + 3*(size_t)ch);
            const float4 c1 = __ldg (tr.nodes + 3*(size_t)ch + 1);
            if (box_ball (v3 (c0), v3 (c1), c, rr))
            {
                if ((__float_as_uint (c0.w) & 3u) == VOXGPU_KIND_DENSE) return true;
                stack[sp++] = ch;
            }
        }
    }
    return false;
}

#endif

/* ======================================================================
 * Production traversal.  Same search, same fp32 operations in the same
 * order as ray_tree<> above (which stays as the instrumented / readable
 * form and as an on-device cross-check), reorganised after the first ncu
 * profile (profiles/r01_*):
 *
 *  - hit_box only divides on the axes where the origin was clamped.  On an
 *    unclamped axis the reference computes (o-o)/d = +-0 or NaN, which can
 *    never become the winning positive distance (geom.c:79-87), so the
 *    division -- which also fell into the IEEE slow path because of its zero
 *    numerator, 15% of all issued instructions -- is dropped without changing
 *    any result, NaN/inf inputs included.
 *  - the bound checks of hit_box / hit_plane_within_box are evaluated as
 *    predicates instead of early-exit branches.
 *  - one loop iteration = "take pending work until there is a call" followed
 *    by exactly one call, so the lanes of a warp meet at the (expensive) box
 *    test every iteration instead of drifting between call / pop / expand
 *    iterations.
 * ====================================================================== */

__device__ __forceinline__ bool ray_box_fast (const V3 mn, const V3 mx, const V3 o, const V3 d, V3 &res)
{
    const bool lx = o.x < mn.x, hx = o.x > mx.x;
    const bool ly = o.y < mn.y, hy = o.y > mx.y;
    const bool lz = o.z < mn.z, hz = o.z > mx.z;
    const bool kx = lx | hx, ky = ly | hy, kz = lz | hz;
    res = o;
    if (!(kx | ky | kz)) return true;                 /* origin inside: geom.c:73 */

    float cx = lx ? mn.x : (hx ? mx.x : o.x);
    float cy = ly ? mn.y : (hy ? mx.y : o.y);
    float cz = lz ? mn.z : (hz ? mx.z : o.z);

    int axis = 0;
    float tmax = -1.0f;
    if (kx) { const float t = __fdiv_rn (__fsub_rn (cx, o.x), d.x); if (t > tmax) { tmax = t; axis = 0; } }
    if (ky) { const float t = __fdiv_rn (__fsub_rn (cy, o.y), d.y); if (t > tmax) { tmax = t; axis = 1; } }
    if (kz) { const float t = __fdiv_rn (__fsub_rn (cz, o.z), d.z); if (t > tmax) { tmax = t; axis = 2; } }

    const float px = __fadd_rn (o.x, __fmul_rn (tmax, d.x));
    const float py = __fadd_rn (o.y, __fmul_rn (tmax, d.y));
    const float pz = __fadd_rn (o.z, __fmul_rn (tmax, d.z));
    bool ok = tmax > 0.0f;
    if (axis != 0) { ok &= !(px < mn.x || px > mx.x); cx = px; }
    if (axis != 1) { ok &= !(py < mn.y || py > mx.y); cy = py; }
    if (axis != 2) { ok &= !(pz < mn.z || pz > mx.z); cz = pz; }
    res.x = cx; res.y = cy; res.z = cz;
    return ok;
}

/* geom.c:101-123; returns validity, point in res (garbage when invalid) */
__device__ __forceinline__ bool ray_plane_fast (const V3 o, const V3 d, float plane, float oa, float da,
                                                const V3 mn, const V3 mx, V3 &res)
{
    float k = __fsub_rn (plane, oa);
    const bool cand = !(k == 0.0f || ((da < 0.0f) != (k < 0.0f)));
    if (!cand) return false;
    k = __fdiv_rn (k, da);
    res.x = __fadd_rn (o.x, __fmul_rn (k, d.x));
    res.y = __fadd_rn (o.y, __fmul_rn (k, d.y));
    res.z = __fadd_rn (o.z, __fmul_rn (k, d.z));
    return !(res.x < mn.x || res.x > mx.x) & !(res.y < mn.y || res.y > mx.y) & !(res.z < mn.z || res.z > mx.z);
}

__device__ __forceinline__ void ray_tree_fast (const TreeRef &tr, unsigned start, const V3 origin, const V3 d,
                                               float4 *stack, RayResult &out)
{
    int sp = 0;
    unsigned cur = start;
    V3 org = origin;
    bool have_call = true;

    out.leaf = -1;
    out.voxel = -1;

    for (;;)
    {
        /* ---- pending work until there is a call to make ---- */
        while (!have_call)
        {
            if (sp == 0) return;
            const float4 e = stack[--sp];
            const unsigned tag = __float_as_uint (e.w);
            org.x = e.x; org.y = e.y; org.z = e.z;
            if (!(tag & VOXGPU_EXPAND_BIT)) { cur = tag; have_call = true; break; }

            /* search.c:99-148: the first child missed; evaluate the dividing planes */
            const unsigned id = tag & ~VOXGPU_EXPAND_BIT;
            const float4 q0 = __ldg (tr.nodes + 3*(size_t)id);
            const float4 q1 = __ldg (tr.nodes + 3*(size_t)id + 1);
            const float4 q2 = __ldg (tr.nodes + 3*(size_t)id + 2);
            const V3 mn = v3 (q0), mx = v3 (q1), c = v3 (q2);
            const unsigned mask = (__float_as_uint (q0.w) >> 8) & 0xffu;
            const unsigned first = __float_as_uint (q1.w);

            V3 p0, p1, p2;
            const bool v0 = ray_plane_fast (org, d, c.x, org.x, d.x, mn, mx, p0);
            const bool v1 = ray_plane_fast (org, d, c.y, org.y, d.y, mn, mx, p1);
            const bool v2 = ray_plane_fast (org, d, c.z, org.z, d.z, mn, mx, p2);
            /* compact the valid ones in axis order 0,1,2 (search.c:108-113) into slots A,B,C
               (explicit selects: no dynamically indexed arrays, nothing spills to local memory) */
            const int n = (int) v0 + (int) v1 + (int) v2;
            if (n == 0) continue;
            const float m0 = l1_metric (org, p0), m1 = l1_metric (org, p1), m2 = l1_metric (org, p2);
            V3 pa, pb, pc = p2;
            float ma, mb, mc = m2;
            unsigned aa, ab, ac = 2u;
            if (v0) { pa = p0; ma = m0; aa = 0u; } else if (v1) { pa = p1; ma = m1; aa = 1u; } else { pa = p2; ma = m2; aa = 2u; }
            if (v0 & v1) { pb = p1; mb = m1; ab = 1u; } else { pb = p2; mb = m2; ab = 2u; }
            /* search.c:115-134: selection sort, strict <, swap semantics */
#define VOXGPU_SWAP(P, M, A, Q, N, B) do { V3 tp = P; P = Q; Q = tp; float tm = M; M = N; N = tm; \
                                           unsigned ta = A; A = B; B = ta; } while (0)
            if (n >= 2 && mb < ma) VOXGPU_SWAP (pa, ma, aa, pb, mb, ab);
            if (n == 3)
            {
                if (mc < ma) VOXGPU_SWAP (pa, ma, aa, pc, mc, ac);
                if (mc < mb) VOXGPU_SWAP (pb, mb, ab, pc, mc, ac);
            }
#undef VOXGPU_SWAP
            const unsigned sub = subspace_corrected (c, org, d);
            const unsigned s0 = sub ^ (1u << aa);
            const unsigned s1 = (n > 1) ? s0 ^ (1u << ab) : 8u;
            const unsigned s2 = (n > 2) ? s1 ^ (1u << ac) : 8u;
            const bool e0 = (mask >> s0) & 1u, e1 = (mask >> s1) & 1u, e2 = (mask >> s2) & 1u;   /* mask has 8 bits: s=8 -> 0 */
            /* nearest existing child is called now, the others wait on the stack (farthest deepest) */
            if (e2 && (e0 | e1))
                stack[sp++] = make_float4 (pc.x, pc.y, pc.z, __uint_as_float (child_id (first, mask, s2)));
            if (e1 && e0)
                stack[sp++] = make_float4 (pb.x, pb.y, pb.z, __uint_as_float (child_id (first, mask, s1)));
            if (e0)      { cur = child_id (first, mask, s0); org = pa; have_call = true; }
            else if (e1) { cur = child_id (first, mask, s1); org = pb; have_call = true; }
            else if (e2) { cur = child_id (first, mask, s2); org = pc; have_call = true; }
        }

        /* ---- one call: vox_ray_tree_intersection (cur, org, d) ---- */
        have_call = false;
        const float4 q0 = __ldg (tr.nodes + 3*(size_t)cur);
        const float4 q1 = __ldg (tr.nodes + 3*(size_t)cur + 1);
        const unsigned meta = __float_as_uint (q0.w);
        const unsigned link = __float_as_uint (q1.w);
        const unsigned kind = meta & 3u;

        V3 entry;
        if (!ray_box_fast (v3 (q0), v3 (q1), org, d, entry)) continue;

        if (kind == VOXGPU_KIND_INNER)
        {
            const float4 q2 = __ldg (tr.nodes + 3*(size_t)cur + 2);
            const unsigned mask = (meta >> 8) & 0xffu;
            const unsigned sub = subspace_corrected (v3 (q2), entry, d);
            stack[sp++] = make_float4 (entry.x, entry.y, entry.z, __uint_as_float (cur | VOXGPU_EXPAND_BIT));
            if (mask & (1u << sub))
            {
                cur = child_id (link, mask, sub);
                org = entry;
                have_call = true;
            }
            continue;
        }
        if (kind == VOXGPU_KIND_DENSE)
        {
            out.leaf = (int) cur; out.voxel = -1; out.point = entry;
            return;
        }
        /* leaf, search.c:39-81 */
        {
            const unsigned n = meta >> 2;
            float closest = __int_as_float (0x7f800000);
            for (unsigned i=0; i<n; i++)
            {
                const float4 dq = __ldg (tr.dots + (size_t)link + i);
                V3 vmin = v3 (dq), vmax, far;
                vmax.x = __fadd_rn (vmin.x, tr.vx);
                vmax.y = __fadd_rn (vmin.y, tr.vy);
                vmax.z = __fadd_rn (vmin.z, tr.vz);
                if (ray_box_fast (vmin, vmax, entry, d, far))
                {
                    const float dist = l1_metric (entry, far);
                    if (dist == 0.0f)
                    {
                        out.leaf = (int) cur; out.voxel = (int)(link + i); out.point = far;
                        return;
                    }
                    if (dist < closest)
                    {
                        closest = dist;
                        out.leaf = (int) cur; out.voxel = (int)(link + i); out.point = far;
                    }
                }
            }
            if (out.leaf >= 0) return;
        }
    }
}
