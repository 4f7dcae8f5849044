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
 * Device tree layout (built by voxgpu_tree_upload), node ids
 * are breadth-first positions:
 *   rec[id] (32 B, 32-byte aligned = one L2 sector, fetched by ONE 256-bit load,
 *            LDG.E.256 -- sm_100 only):
 *        lo = { bb_min.xyz , meta }   meta: bits 0-1 kind (0 inner, 1 leaf, 2 dense leaf)
 *                                           inner: bits 8-15 child mask
 *                                           leaf : bits 2-31 number of dots
 *        hi = { bb_max.xyz , link }   inner: id of first non-NULL child (children are
 *                                           consecutive, in subspace order)
 *                                     leaf : index of first dot in dots[]
 *   ctr[id]  = { center.xyz , 0 }     inner nodes only (16 B, one LDG.128)
 *   dots[i]  = { voxel min corner xyz, 0 }
 * Every node visit reads exactly one sector; an inner node adds half a sector.
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
#define VOXGPU_RAY_STACK (3*VOXGPU_MAX_DEPTH)    /* <= 3 pending entries per level (eager form; the lazy form needs 2) */
#define VOXGPU_BALL_STACK (7*VOXGPU_MAX_DEPTH+8) /* <= 7 pending siblings per level */
#define VOXGPU_EXPAND_BIT 0x80000000u

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3 (float4 q) { V3 r; r.x = q.x; r.y = q.y; r.z = q.z; return r; }
__device__ __forceinline__ float comp (const V3 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct __align__(32) NodeRec { float4 lo, hi; };

struct TreeRef
{
    const NodeRec *rec;
    const float4 *ctr;
    const float4 *dots;
    float vx, vy, vz;       /* vox_voxel captured at upload */
    /* the first `ntop` nodes (breadth-first order = the top levels of the tree) staged in shared
       memory by the kernel: 2 x float4 record + float4 centre per node; ntop = 0: nothing staged */
    const float4 *top_rec, *top_ctr;
    unsigned ntop;
    unsigned ntop_avail;    /* min (nodes in the tree, VOXGPU_SMEM_TOP): what a kernel may stage */
};

/* Build-time knobs for A/B runs (tools/ab_variants.sh). */
#ifndef VOXGPU_SMEM_TOP
#define VOXGPU_SMEM_TOP 0          /* nodes staged per CTA by the render kernel: 73 = levels 0-2, 585 = levels 0-3 */
#endif
#ifndef VOXGPU_EARLY_CTR
#define VOXGPU_EARLY_CTR 1         /* fetch the split centre together with the node record */
#endif
#ifndef VOXGPU_PREFETCH
#define VOXGPU_PREFETCH 0          /* prefetch.global.L1 of the child a lane descends into */
#endif

/* one 32-byte node record with a single 256-bit read-only load */
__device__ __forceinline__ void load_rec (const NodeRec *p, float4 &lo, float4 &hi)
{
    asm ("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
         : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
         : "l"(p));
}

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
                float4 q0, q1;
                load_rec (tr.rec + id, q0, q1);
                const float4 q2 = __ldg (tr.ctr + id);
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
        float4 q0, q1;
        load_rec (tr.rec + cur, q0, q1);
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
            const float4 q2 = __ldg (tr.ctr + cur);
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
        float4 q0, q1;
        load_rec (tr.rec + start, q0, q1);
        if (!box_ball (v3 (q0), v3 (q1), c, rr)) return false;
        stack[sp++] = start;
    }
    while (sp > 0)
    {
        const unsigned id = stack[--sp];
        float4 q0, q1;
        load_rec (tr.rec + id, q0, q1);
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
            float4 c0, c1;
            load_rec (tr.rec + ch, c0, c1);
            if (box_ball (v3 (c0), v3 (c1), c, rr))
            {
                if ((__float_as_uint (c0.w) & 3u) == VOXGPU_KIND_DENSE) return true;
                stack[sp++] = ch;
            }
        }
    }
    return false;
}

/* ======================================================================
 * Production traversal.  Same search, same fp32 results as ray_tree<> above
 * (which stays as the readable, instrumented form that mirrors the
 * reference's control flow one to one), reorganised around what the ncu
 * profiles and the per-tile trace showed (profiles/r01_*): the
 * kernel is bound by instruction issue at ~15 of 32 lanes active, and its
 * duration by the few heaviest tiles, i.e. by instructions per iteration.
 *
 *  - EAGER EXPANSION.  The reference evaluates an inner node's dividing
 *    planes only after the search of the first child has missed
 *    (search.c:99-148).  Whether and where the ray crosses them is a pure
 *    function of (node, entry point, direction), so it is computed when the
 *    node is visited -- while all lanes of the warp are at "a node" -- and the
 *    resulting calls (child, plane hit point) are parked on the stack in the
 *    reference's order.  A warp has ONE convergence point per iteration
 *    instead of alternating call / pop / expand phases; the node record is
 *    not re-read and the entry point is not stacked.
 *  - The selection sort of search.c:115-134 is evaluated as a rank
 *    computation on four comparisons and the pending calls are scattered to
 *    their stack slots; nothing is moved between registers.  (Checked
 *    exhaustively against the swap sequence, ties included.)
 *  - WARP-COOPERATIVE LEAVES.  A leaf holds up to 7 voxels that the reference
 *    tests one after the other (search.c:51-79); typically 5 of 32 lanes are
 *    at a leaf at the same time and the other 27 wait.  The tests are
 *    independent, so the lanes at a leaf hand their voxels out: each group of
 *    8 lanes serves one leaf (one voxel per lane), reduces to the reference's
 *    winner -- smallest L1 distance, earliest voxel on ties, which is also what
 *    the "distance == 0 ends the loop" shortcut of search.c:65-71 selects,
 *    distances being non-negative -- and returns it to the owning lane.
 *    The traversal is therefore warp-synchronous: all 32 lanes call it.
 *  - DIVISION.  (c - o) / d with d fixed per ray: r = RN(1/d) once per ray,
 *    then q = RN(x*r), q' = RN(q + (x - d*q)*r) with the residual exact by FMA
 *    is the correctly rounded quotient (Markstein's theorem; checked against
 *    x/d on 1.0e11 operand pairs, exhaustively in d's mantissa:
 *    tools/probes/divtest.c) as long as nothing under- or overflows, which an
 *    exponent-range test on x and (once per ray) on d guarantees; operands
 *    outside the range take IEEE division (__fdiv_rn).  3 FMA-pipe
 *    instructions instead of the ~11 of the generic sequence.
 *  - hit_box divides only where the origin was clamped: on an unclamped axis
 *    the reference computes (o-o)/d = +-0 or NaN, which can never become the
 *    winning positive distance (geom.c:79-87).  Divisions are issued
 *    unconditionally on a harmless numerator instead of under a branch, so
 *    the three of a box test (and of the three planes) overlap.
 *  - bound checks are predicates, not early-exit branches.
 * ====================================================================== */

struct RayK                 /* per-ray constants */
{
    V3 d;                   /* direction */
    V3 r;                   /* RN(1/d) */
    bool nx, ny, nz;        /* d < 0 */
    bool dok;               /* every |d| in [2^-60, 2^61): div3 may use the FMA form */
};

#define VOXGPU_DIV_LO 8.6736173798840355e-19f   /* 2^-60 */
#define VOXGPU_DIV_HI 2.305843009213694e+18f    /* 2^61 */

__device__ __forceinline__ bool div_in_range (float x) { const float a = fabsf (x); return a >= VOXGPU_DIV_LO && a < VOXGPU_DIV_HI; }

__device__ __forceinline__ RayK ray_constants (const V3 d)
{
    RayK k;
    k.d = d;
    k.nx = d.x < 0.0f; k.ny = d.y < 0.0f; k.nz = d.z < 0.0f;
    k.dok = div_in_range (d.x) && div_in_range (d.y) && div_in_range (d.z);
    k.r.x = __frcp_rn (d.x); k.r.y = __frcp_rn (d.y); k.r.z = __frcp_rn (d.z);
    return k;
}

/* (x/d.x, y/d.y, z/d.z), correctly rounded */
__device__ __forceinline__ void div3 (float x, float y, float z, const V3 d, const V3 r, bool dok,
                                      float &qx, float &qy, float &qz)
{
    const float lo = fminf (fminf (fabsf (x), fabsf (y)), fabsf (z));
    const float hi = fmaxf (fmaxf (fabsf (x), fabsf (y)), fabsf (z));
    if (dok && lo >= VOXGPU_DIV_LO && hi < VOXGPU_DIV_HI)
    {
        const float ax = __fmul_rn (x, r.x), ay = __fmul_rn (y, r.y), az = __fmul_rn (z, r.z);
        qx = __fmaf_rn (__fmaf_rn (-d.x, ax, x), r.x, ax);
        qy = __fmaf_rn (__fmaf_rn (-d.y, ay, y), r.y, ay);
        qz = __fmaf_rn (__fmaf_rn (-d.z, az, z), r.z, az);
    }
    else
    {
        qx = __fdiv_rn (x, d.x); qy = __fdiv_rn (y, d.y); qz = __fdiv_rn (z, d.z);
    }
}

__device__ __forceinline__ bool ray_box_fast (const V3 mn, const V3 mx, const V3 o, const V3 d, const V3 r, bool dok, V3 &res)
{
    const bool lx = o.x < mn.x, hx = o.x > mx.x;
    const bool ly = o.y < mn.y, hy = o.y > mx.y;
    const bool lz = o.z < mn.z, hz = o.z > mx.z;
    const bool kx = lx | hx, ky = ly | hy, kz = lz | hz;
    bool ok = true;
    res = o;                                            /* origin inside: geom.c:73 */
    if (kx | ky | kz)
    {
        float cx = lx ? mn.x : (hx ? mx.x : o.x);
        float cy = ly ? mn.y : (hy ? mx.y : o.y);
        float cz = lz ? mn.z : (hz ? mx.z : o.z);

        /* numerator 1 on unclamped axes: 0/d would take the slow path for nothing */
        float tx, ty, tz;
        div3 (kx ? __fsub_rn (cx, o.x) : 1.0f, ky ? __fsub_rn (cy, o.y) : 1.0f, kz ? __fsub_rn (cz, o.z) : 1.0f,
              d, r, dok, tx, ty, tz);
        int axis = 0;
        float tmax = -1.0f;
        if (kx && tx > tmax) { tmax = tx; }
        if (ky && ty > tmax) { tmax = ty; axis = 1; }
        if (kz && tz > tmax) { tmax = tz; axis = 2; }

        const float px = __fadd_rn (o.x, __fmul_rn (tmax, d.x));
        const float py = __fadd_rn (o.y, __fmul_rn (tmax, d.y));
        const float pz = __fadd_rn (o.z, __fmul_rn (tmax, d.z));
        ok = tmax > 0.0f;
        if (axis != 0) { ok &= !(px < mn.x || px > mx.x); cx = px; }
        if (axis != 1) { ok &= !(py < mn.y || py > mx.y); cy = py; }
        if (axis != 2) { ok &= !(pz < mn.z || pz > mx.z); cz = pz; }
        res.x = cx; res.y = cy; res.z = cz;
    }
    return ok;
}

/* geom.c:101-123 for the three dividing planes at once.  Validity in v*, points in p*. */
__device__ __forceinline__ void ray_planes_fast (const V3 o, const RayK &k, const V3 c, const V3 mn, const V3 mx,
                                                 bool &v0, bool &v1, bool &v2, V3 &p0, V3 &p1, V3 &p2)
{
    const float k0 = __fsub_rn (c.x, o.x), k1 = __fsub_rn (c.y, o.y), k2 = __fsub_rn (c.z, o.z);
    const bool c0 = !(k0 == 0.0f) && (k.nx == (k0 < 0.0f));
    const bool c1 = !(k1 == 0.0f) && (k.ny == (k1 < 0.0f));
    const bool c2 = !(k2 == 0.0f) && (k.nz == (k2 < 0.0f));
    float t0, t1, t2;
    div3 (c0 ? k0 : 1.0f, c1 ? k1 : 1.0f, c2 ? k2 : 1.0f, k.d, k.r, k.dok, t0, t1, t2);
    const V3 d = k.d;
    p0.x = __fadd_rn (o.x, __fmul_rn (t0, d.x)); p0.y = __fadd_rn (o.y, __fmul_rn (t0, d.y)); p0.z = __fadd_rn (o.z, __fmul_rn (t0, d.z));
    p1.x = __fadd_rn (o.x, __fmul_rn (t1, d.x)); p1.y = __fadd_rn (o.y, __fmul_rn (t1, d.y)); p1.z = __fadd_rn (o.z, __fmul_rn (t1, d.z));
    p2.x = __fadd_rn (o.x, __fmul_rn (t2, d.x)); p2.y = __fadd_rn (o.y, __fmul_rn (t2, d.y)); p2.z = __fadd_rn (o.z, __fmul_rn (t2, d.z));
    v0 = c0 & !(p0.x < mn.x || p0.x > mx.x) & !(p0.y < mn.y || p0.y > mx.y) & !(p0.z < mn.z || p0.z > mx.z);
    v1 = c1 & !(p1.x < mn.x || p1.x > mx.x) & !(p1.y < mn.y || p1.y > mx.y) & !(p1.z < mn.z || p1.z > mx.z);
    v2 = c2 & !(p2.x < mn.x || p2.x > mx.x) & !(p2.y < mn.y || p2.y > mx.y) & !(p2.z < mn.z || p2.z > mx.z);
}

/* geom.c:53-65 with the direction signs hoisted out of the loop */
__device__ __forceinline__ unsigned subspace_fast (const V3 c, const V3 p, bool nx, bool ny, bool nz)
{
    const bool bx = (c.x > p.x) || (nx && c.x == p.x);
    const bool by = (c.y > p.y) || (ny && c.y == p.y);
    const bool bz = (c.z > p.z) || (nz && c.z == p.z);
    return (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
}

#define VOXGPU_FULL 0xffffffffu
#define VOXGPU_DEAD 0xffffffffu

/*
 * The stack of pending calls of one lane.  Slots [0, VOXGPU_SMEM_STACK) live in shared memory, laid
 * out [slot][thread] as float4: whatever slot each lane of a warp is at, lane l touches banks 4l..4l+3,
 * so a 128-bit access is conflict-free (4 wavefronts, the minimum) -- in local memory the same access
 * with 32 different stack depths costs up to 32 x 4 sector accesses in L1 (the r01 profile had the stack
 * at 5.9 times the L1 sectors of the node fetches).  Deeper slots spill to per-thread local memory.
 * Shared memory is carved out of the same SRAM as L1, so the number of slots is a trade against the L1
 * hit rate of the node records: see DESIGN.md for the measured choice.  0 = everything in local memory.
 */
#ifndef VOXGPU_SMEM_STACK
#define VOXGPU_SMEM_STACK 0
#endif
#define VOXGPU_STACK_THREADS 128    /* block size of every kernel that searches (K1, K1b, K2) */

struct RayStack
{
    float4 *loc;            /* VOXGPU_RAY_STACK entries of per-thread scratch */
#if VOXGPU_SMEM_STACK
    float4 *sm;             /* this thread's column of the CTA's shared slots: slot i at sm[i * VOXGPU_STACK_THREADS] */
#endif
    __device__ __forceinline__ void put (int i, const float4 v) const
    {
#if VOXGPU_SMEM_STACK
        if (i < VOXGPU_SMEM_STACK) sm[i * VOXGPU_STACK_THREADS] = v; else loc[i - VOXGPU_SMEM_STACK] = v;
#else
        loc[i] = v;
#endif
    }
    __device__ __forceinline__ float4 get (int i) const
    {
#if VOXGPU_SMEM_STACK
        if (i < VOXGPU_SMEM_STACK) return sm[i * VOXGPU_STACK_THREADS];
        return loc[i - VOXGPU_SMEM_STACK];
#else
        return loc[i];
#endif
    }
};

/* in a kernel: `VOXGPU_DECLARE_STACK (stack);` declares the local (and shared) storage and a RayStack over it */
#if VOXGPU_SMEM_STACK
#define VOXGPU_DECLARE_STACK(name)                                                              \
    __shared__ float4 name##_shared[VOXGPU_SMEM_STACK * VOXGPU_STACK_THREADS];                  \
    float4 name##_local[VOXGPU_RAY_STACK];                                                      \
    RayStack name; name.loc = name##_local; name.sm = name##_shared + threadIdx.x
#else
#define VOXGPU_DECLARE_STACK(name)                                                              \
    float4 name##_local[VOXGPU_RAY_STACK];                                                      \
    RayStack name; name.loc = name##_local
#endif
#ifndef VOXGPU_COOP_MAX_LANES
#define VOXGPU_COOP_MAX_LANES 8      /* more lanes than this at a leaf: everyone loops over its own voxels
                                        (C2, tools/ab_variants.sh: 8 -> 12.54, 12 -> 12.48, 16 -> 12.40 Grays/s) */
#endif

/*
 * One lane's search state between node visits.  `cur` is the node of the pending call
 * (VOXGPU_DEAD: this lane's search is over), `org` the origin that call is made with.
 * A finished ray's stack is dead: the hit point and voxel wait there (stack[0]) instead
 * of in registers that every branch of the loop would have to carry.
 */
struct LaneSearch
{
    unsigned cur;
    V3 org;
    int sp;
    int leaf;               /* result: breadth-first id of the hit leaf, -1 = none */
    /* written only by a lane that is at a leaf; read (shuffled) only from such lanes */
    V3 lf_e;
    unsigned lf_link, lf_info;
};

__device__ __forceinline__ void lane_search_begin (LaneSearch &s, unsigned start, const V3 origin, bool has_ray)
{
    s.cur = has_ray ? start : VOXGPU_DEAD;
    s.org = origin;
    s.sp = 0;
    s.leaf = -1;
    s.lf_e = origin; s.lf_link = 0; s.lf_info = 0;
}

/*
 * ONE node visit of vox_ray_tree_intersection (search.c:8-153) for every live lane of
 * the warp, followed by the cooperative leaf tests and the pop of the next pending call.
 * WARP-SYNCHRONOUS: all 32 lanes must call it together.
 */
__device__ __forceinline__ void warp_search_step (const TreeRef &tr, const RayK &k, LaneSearch &s, const RayStack &stack, const int lane)
{
    /* what a lane does after the node visit of this iteration */
    enum { IDLE = 0, GO = 1, POP = 2, LEAF = 3 };
    const float inf = __int_as_float (0x7f800000);
    const V3 d = k.d;
    int st = IDLE;
    if (s.cur != VOXGPU_DEAD)
    {
        /* ---- one call: vox_ray_tree_intersection (cur, org, d) ---- */
        float4 q0, q1;
#if VOXGPU_SMEM_TOP
        const bool top = s.cur < tr.ntop;
        if (top) { q0 = tr.top_rec[2*s.cur]; q1 = tr.top_rec[2*s.cur + 1]; }
        else load_rec (tr.rec + s.cur, q0, q1);
#else
        load_rec (tr.rec + s.cur, q0, q1);
#endif
#if VOXGPU_EARLY_CTR
        /* the split centre is fetched with the record, not after the box test has passed: one round trip to
           L2 per node visit instead of two on the critical path of the longest rays (a wasted 16-byte load
           for leaves and for boxes the ray misses) */
        const float4 qc = __ldg (tr.ctr + s.cur);
#endif
        const unsigned meta = __float_as_uint (q0.w);
        const unsigned link = __float_as_uint (q1.w);
        const unsigned kind = meta & 3u;
        const V3 mn = v3 (q0), mx = v3 (q1);
        V3 entry;

        st = POP;                                                               /* search.c:22-27 */
        if (ray_box_fast (mn, mx, s.org, d, k.r, k.dok, entry))
        {
            if (kind == VOXGPU_KIND_INNER)
            {
#if VOXGPU_SMEM_TOP
                const V3 c = v3 (top ? tr.top_ctr[s.cur] : __ldg (tr.ctr + s.cur));
#elif VOXGPU_EARLY_CTR
                const V3 c = v3 (qc);
#else
                const V3 c = v3 (__ldg (tr.ctr + s.cur));
#endif
                const unsigned mask = (meta >> 8) & 0xffu;
                const unsigned sub = subspace_fast (c, entry, k.nx, k.ny, k.nz);
#if VOXGPU_PREFETCH
                /* the child this lane descends into is known now; its record is needed at the top of
                   the next iteration, after the pushes below and the leaf phase of the other lanes */
                if ((mask >> sub) & 1u)
                {
                    const unsigned nxt = child_id (link, mask, sub);
                    asm volatile ("prefetch.global.L1 [%0];" :: "l"(tr.rec + nxt));
                }
#endif

                /* search.c:105-113: the three dividing planes, from the entry point */
                V3 p0, p1, p2;
                bool v0, v1, v2;
                ray_planes_fast (entry, k, c, mn, mx, v0, v1, v2, p0, p1, p2);
                if (v0 | v1 | v2)
                {
                    /* search.c:115-134 as ranks.  Invalid planes get +inf and sort last. */
                    const float m0 = v0 ? l1_metric (entry, p0) : inf;
                    const float m1 = v1 ? l1_metric (entry, p1) : inf;
                    const float m2 = v2 ? l1_metric (entry, p2) : inf;
                    const bool A = m1 < m0, B = m2 < m1, C = m2 < m0, D = m0 < m1;
                    const bool b01 = !A && (!C || D);      /* plane 0 is tried before plane 1 */
                    const bool b02 = !C && (!A || !B);
                    const bool b12 = !B && (A || !C);
                    /* search.c:137: the subspace flips of the planes tried so far accumulate */
                    const unsigned s0 = sub ^ 1u ^ ((v1 && !b01) ? 2u : 0u) ^ ((v2 && !b02) ? 4u : 0u);
                    const unsigned s1 = sub ^ 2u ^ ((v0 &&  b01) ? 1u : 0u) ^ ((v2 && !b12) ? 4u : 0u);
                    const unsigned s2 = sub ^ 4u ^ ((v0 &&  b02) ? 1u : 0u) ^ ((v1 &&  b12) ? 2u : 0u);
                    const bool e0 = v0 && ((mask >> s0) & 1u);
                    const bool e1 = v1 && ((mask >> s1) & 1u);
                    const bool e2 = v2 && ((mask >> s2) & 1u);
                    /* the call tried first must be popped first: slot = number of later calls */
                    if (e0) stack.put (s.sp + (int)(e1 && b01) + (int)(e2 && b02),
                                       make_float4 (p0.x, p0.y, p0.z, __uint_as_float (child_id (link, mask, s0))));
                    if (e1) stack.put (s.sp + (int)(e0 && !b01) + (int)(e2 && b12),
                                       make_float4 (p1.x, p1.y, p1.z, __uint_as_float (child_id (link, mask, s1))));
                    if (e2) stack.put (s.sp + (int)(e0 && !b02) + (int)(e1 && !b12),
                                       make_float4 (p2.x, p2.y, p2.z, __uint_as_float (child_id (link, mask, s2))));
                    s.sp += (int) e0 + (int) e1 + (int) e2;
                }
                /* search.c:83-97: the child the entry point lies in goes first */
                if ((mask >> sub) & 1u) { s.cur = child_id (link, mask, sub); s.org = entry; st = GO; }
            }
            else if (kind == VOXGPU_KIND_DENSE)
            {
                /* search.c:31-37 */
                s.leaf = (int) s.cur;
                stack.put (0, make_float4 (entry.x, entry.y, entry.z, __int_as_float (-1)));
                s.cur = VOXGPU_DEAD; st = IDLE;
            }
            else
            {
                s.lf_e = entry; s.lf_link = link; s.lf_info = (meta >> 2) | (k.dok ? 0x80000000u : 0u);
                st = LEAF;
            }
        }
    }

    /* ---- leaves, search.c:39-81 ---- */
    const unsigned leafmask = __ballot_sync (VOXGPU_FULL, st == LEAF);
    if (leafmask)
    {
        float best = inf;           /* owner's result */
        int best_i = 0;
        V3 best_p = s.lf_e;
        const bool coop = __popc (leafmask) <= VOXGPU_COOP_MAX_LANES &&
                          !__any_sync (VOXGPU_FULL, st == LEAF && (s.lf_info & 0x7fffffffu) > 8u);
        if (coop)
        {
            const int group = lane >> 3, slot = lane & 7;
            const int myrank = __popc (leafmask & ((1u << lane) - 1u));
            unsigned m = leafmask;
            for (int base = 0; m; base += 4)
            {
                /* the four lowest remaining leaf lanes are served by the four groups of eight */
                const int l0 = __ffs (m) - 1; m &= m - 1;
                const int l1 = __ffs (m) - 1; m &= m - 1;
                const int l2 = __ffs (m) - 1; m &= m - 1;
                const int l3 = __ffs (m) - 1; m &= m - 1;
                const int src = group == 0 ? l0 : (group == 1 ? l1 : (group == 2 ? l2 : l3));
                const int from = src >= 0 ? src : lane;
                V3 e, dd, rr;
                e.x = __shfl_sync (VOXGPU_FULL, s.lf_e.x, from); e.y = __shfl_sync (VOXGPU_FULL, s.lf_e.y, from); e.z = __shfl_sync (VOXGPU_FULL, s.lf_e.z, from);
                dd.x = __shfl_sync (VOXGPU_FULL, d.x, from); dd.y = __shfl_sync (VOXGPU_FULL, d.y, from); dd.z = __shfl_sync (VOXGPU_FULL, d.z, from);
                rr.x = __shfl_sync (VOXGPU_FULL, k.r.x, from); rr.y = __shfl_sync (VOXGPU_FULL, k.r.y, from); rr.z = __shfl_sync (VOXGPU_FULL, k.r.z, from);
                const unsigned lnk = __shfl_sync (VOXGPU_FULL, s.lf_link, from);
                const unsigned info = __shfl_sync (VOXGPU_FULL, s.lf_info, from);
                float dist = inf;
                V3 far = e;
                if (src >= 0 && (unsigned) slot < (info & 0x7fffffffu))
                {
                    const float4 dq = __ldg (tr.dots + (size_t) lnk + slot);
                    V3 vmin = v3 (dq), vmax;
                    vmax.x = __fadd_rn (vmin.x, tr.vx);
                    vmax.y = __fadd_rn (vmin.y, tr.vy);
                    vmax.z = __fadd_rn (vmin.z, tr.vz);
                    if (ray_box_fast (vmin, vmax, e, dd, rr, (info >> 31) != 0, far))
                    {
                        const float t = l1_metric (e, far);
                        if (t < inf) dist = t;              /* search.c:72: NaN and inf never win */
                    }
                }
                /* smallest distance, earliest voxel on ties, within the group of eight */
                int idx = slot;
                #pragma unroll
                for (int off = 4; off; off >>= 1)
                {
                    const float od = __shfl_xor_sync (VOXGPU_FULL, dist, off);
                    const int oi = __shfl_xor_sync (VOXGPU_FULL, idx, off);
                    if (od < dist || (od == dist && oi < idx)) { dist = od; idx = oi; }
                }
                /* back to the owner: the group that served lane rank myrank holds its answer */
                const int g = myrank - base;
                const bool mine = st == LEAF && g >= 0 && g < 4;
                const int head = mine ? 8 * g : lane;
                const float bd = __shfl_sync (VOXGPU_FULL, dist, head);
                const int bi = __shfl_sync (VOXGPU_FULL, idx, head);
                const int win = mine ? head + bi : lane;
                const float fx = __shfl_sync (VOXGPU_FULL, far.x, win);
                const float fy = __shfl_sync (VOXGPU_FULL, far.y, win);
                const float fz = __shfl_sync (VOXGPU_FULL, far.z, win);
                if (mine) { best = bd; best_i = bi; best_p.x = fx; best_p.y = fy; best_p.z = fz; }
            }
        }
        else if (st == LEAF)
        {
            const unsigned n = s.lf_info & 0x7fffffffu;
            for (unsigned i=0; i<n; i++)
            {
                const float4 dq = __ldg (tr.dots + (size_t) s.lf_link + i);
                V3 vmin = v3 (dq), vmax, far;
                vmax.x = __fadd_rn (vmin.x, tr.vx);
                vmax.y = __fadd_rn (vmin.y, tr.vy);
                vmax.z = __fadd_rn (vmin.z, tr.vz);
                if (ray_box_fast (vmin, vmax, s.lf_e, d, k.r, k.dok, far))
                {
                    const float dist = l1_metric (s.lf_e, far);
                    if (dist < best) { best = dist; best_i = (int) i; best_p = far; }
                    if (dist == 0.0f) break;                /* search.c:65-71 */
                }
            }
        }
        if (st == LEAF)
        {
            if (best < inf)
            {
                s.leaf = (int) s.cur;
                stack.put (0, make_float4 (best_p.x, best_p.y, best_p.z, __uint_as_float (s.lf_link + (unsigned) best_i)));
                s.cur = VOXGPU_DEAD;
            }
            else st = POP;
        }
    }

    if (st == POP)
    {
        /* this call missed: go on with the nearest pending one */
        if (s.sp == 0) s.cur = VOXGPU_DEAD;
        else
        {
            const float4 e = stack.get (--s.sp);
            s.org.x = e.x; s.org.y = e.y; s.org.z = e.z;
            s.cur = __float_as_uint (e.w);
        }
    }
}

/* the result of a finished lane (s.cur == VOXGPU_DEAD) */
__device__ __forceinline__ void lane_search_result (const LaneSearch &s, const RayStack &stack, RayResult &out)
{
    out.leaf = s.leaf;
    out.voxel = -1;
    if (s.leaf >= 0)
    {
        const float4 h = stack.get (0);
        out.point.x = h.x; out.point.y = h.y; out.point.z = h.z;
        out.voxel = __float_as_int (h.w);
    }
}

/*
 * vox_ray_tree_intersection (search.c:8-153) for the 32 rays of a warp.
 * WARP-SYNCHRONOUS: all 32 lanes must call it together; `has_ray` says whether
 * this lane has a ray at all.  `stack`: VOXGPU_DECLARE_STACK.  Returns the number of iterations the warp needed.
 */
__device__ __forceinline__ int ray_tree_warp (const TreeRef &tr, unsigned start, const V3 origin, const V3 dir,
                                              bool has_ray, const RayStack &stack, RayResult &out)
{
    const RayK k = ray_constants (dir);
    const int lane = threadIdx.x & 31;
    LaneSearch s;
    lane_search_begin (s, start, origin, has_ray);
    int iterations = 0;             /* node visits of the slowest lane: the cost of this warp's rays */
    while (__any_sync (VOXGPU_FULL, s.cur != VOXGPU_DEAD))
    {
        iterations++;
        warp_search_step (tr, k, s, stack, lane);
    }
    lane_search_result (s, stack, out);
    return iterations;
}

#endif /* VOXGPU_TRAVERSE_CUH */
