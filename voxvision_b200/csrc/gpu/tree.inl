/* tree.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   device copy of a tree: layout conversion, upload, the cut for the screen mask. */
/* -------------------------------------------------------------------- tree */

struct voxgpu_tree
{
    int device;
    uint64_t n_nodes, n_dots;
    uint32_t depth;
    NodeRec *rec;    /* 32 B per node: box + meta + link */
    float4 *ctr;     /* 16 B per node: split centre (inner nodes) */
    float4 *dots;
    float voxel[3];
    float root_min[3], root_max[3];   /* host copy of node 0's box, for screen-space culling */
    /* a cut through the tree for the screen mask (k_mark): every node of one level plus the leaves
       above it, so that every voxel lies in the box of exactly one node of the cut */
    unsigned *cut;
    unsigned n_cut, cut_level;
    /* tile cost feedback (render_launch.inl): node visits of the slowest lane of every 8x4 tile in the last BEST
       frame of this geometry, a hint for the ticket order of the next one; made on first use, under the
       device's render lock */
    unsigned char *cost;
    int cost_tiles_x, cost_tiles_y;
    /* a coarse cut kept on the HOST (boxes as six arrays: min x, y, z, max x, y, z): the presenter projects it
       every frame to bound the columns each group of rows can show (present.inl: visible_spans) */
    float *hbox;
    unsigned n_hbox;
};
#define VOXGPU_CUT_MAX 131072u
#define VOXGPU_HOST_CUT_MAX 512u

extern "C" voxgpu_tree* voxgpu_tree_upload (const voxgpu_flat *f)
{
    if (!f || f->n_nodes == 0) { fail ("voxgpu_tree_upload: empty tree"); return NULL; }
    if (f->n_nodes >= 0x7fffffffull || f->n_dots >= 0x7fffffffull)
    { fail ("voxgpu_tree_upload: tree too large for 31-bit ids (%llu nodes, %llu dots)",
            (unsigned long long) f->n_nodes, (unsigned long long) f->n_dots); return NULL; }
    if (f->depth > VOXGPU_MAX_DEPTH)
    { fail ("voxgpu_tree_upload: tree depth %u exceeds the device stack (%d levels)", f->depth, VOXGPU_MAX_DEPTH); return NULL; }

    voxgpu_tree *t = (voxgpu_tree*) calloc (1, sizeof *t);
    t->n_nodes = f->n_nodes; t->n_dots = f->n_dots; t->depth = f->depth;
    memcpy (t->root_min, f->bb_min, sizeof t->root_min);
    memcpy (t->root_max, f->bb_max, sizeof t->root_max);
    memcpy (t->voxel, f->voxel, sizeof t->voxel);
    if (cudaGetDevice (&t->device) != cudaSuccess) { fail ("no CUDA device: %s", cudaGetErrorString (cudaGetLastError ())); free (t); return NULL; }

    const size_t rec_bytes = (size_t) f->n_nodes * sizeof (NodeRec);
    const size_t ctr_bytes = (size_t) f->n_nodes * sizeof (float4);
    const size_t dot_bytes = (size_t) (f->n_dots ? f->n_dots : 1) * sizeof (float4);
    size_t stage_bytes = rec_bytes > dot_bytes ? rec_bytes : dot_bytes;
    void *stage_v = NULL;
    cudaStream_t s = NULL;
    cudaError_t e;
    if ((e = cudaMalloc (&t->rec, rec_bytes)) != cudaSuccess ||
        (e = cudaMalloc (&t->ctr, ctr_bytes)) != cudaSuccess ||
        (e = cudaMalloc (&t->dots, dot_bytes)) != cudaSuccess ||
        (e = cudaHostAlloc (&stage_v, stage_bytes, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking)) != cudaSuccess)
    {
        fail ("voxgpu_tree_upload: %s", cudaGetErrorString (e));
        goto bad;
    }

    /* device records, written straight into the pinned staging buffer */
    {
        NodeRec *stage = (NodeRec*) stage_v;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            const uint32_t fl = f->flags[id];
            uint32_t meta, link;
            if (fl & VOXGPU_DENSE_LEAF) { meta = VOXGPU_KIND_DENSE; link = 0; }
            else if (fl & VOXGPU_LEAF)
            {
                if (f->dots_num[id] >= (1u << 30)) { fail ("leaf with %u dots", f->dots_num[id]); goto bad; }
                meta = VOXGPU_KIND_LEAF | (f->dots_num[id] << 2);
                link = (uint32_t) f->dots_off[id];
            }
            else { meta = VOXGPU_KIND_INNER | ((f->child_mask[id] & 0xffu) << 8); link = f->first_child[id]; }
            NodeRec r;
            r.lo.x = f->bb_min[3*id]; r.lo.y = f->bb_min[3*id+1]; r.lo.z = f->bb_min[3*id+2];
            memcpy (&r.lo.w, &meta, 4);
            r.hi.x = f->bb_max[3*id]; r.hi.y = f->bb_max[3*id+1]; r.hi.z = f->bb_max[3*id+2];
            memcpy (&r.hi.w, &link, 4);
            stage[id] = r;
        }
    }
    if ((e = cudaMemcpyAsync (t->rec, stage_v, rec_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize (s)) != cudaSuccess)
    { fail ("voxgpu_tree_upload: node copy: %s", cudaGetErrorString (e)); goto bad; }
    {
        float4 *stage = (float4*) stage_v;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            float4 q; q.x = f->center[3*id]; q.y = f->center[3*id+1]; q.z = f->center[3*id+2]; q.w = 0;
            stage[id] = q;
        }
        if ((e = cudaMemcpyAsync (t->ctr, stage_v, ctr_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
            (e = cudaStreamSynchronize (s)) != cudaSuccess)
        { fail ("voxgpu_tree_upload: centre copy: %s", cudaGetErrorString (e)); goto bad; }
        for (uint64_t i=0; i<f->n_dots; i++)
        {
            float4 q; q.x = f->dots[3*i]; q.y = f->dots[3*i+1]; q.z = f->dots[3*i+2]; q.w = 0;
            stage[i] = q;
        }
        if (f->n_dots &&
            ((e = cudaMemcpyAsync (t->dots, stage_v, f->n_dots * sizeof (float4), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
             (e = cudaStreamSynchronize (s)) != cudaSuccess))
        { fail ("voxgpu_tree_upload: dots copy: %s", cudaGetErrorString (e)); goto bad; }
    }
    {
        /* levels from the breadth-first links; the deepest level whose cut stays below VOXGPU_CUT_MAX nodes */
        unsigned char *level = (unsigned char*) malloc (f->n_nodes);
        uint64_t at[VOXGPU_MAX_DEPTH + 2] = {0}, leaves[VOXGPU_MAX_DEPTH + 2] = {0};
        level[0] = 0;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            const unsigned l = level[id];
            at[l]++;
            if (f->flags[id] & 3u) { leaves[l]++; continue; }
            const int nc = __builtin_popcount (f->child_mask[id] & 0xffu);
            for (int j=0; j<nc; j++) level[f->first_child[id] + j] = (unsigned char) (l + 1);
        }
        unsigned best = 0, hbest = 0;
        uint64_t above = 0;
        for (unsigned l=0; l<f->depth && l <= VOXGPU_MAX_DEPTH; l++)
        {
            if (at[l] + above <= VOXGPU_CUT_MAX) best = l;
            if (at[l] + above <= VOXGPU_HOST_CUT_MAX) hbest = l;
            above += leaves[l];
        }
        {
            /* the host's cut: the same construction, a shallower level */
            unsigned hn = 0;
            for (uint64_t id=0; id<f->n_nodes; id++)
                if (level[id] == hbest || (level[id] < hbest && (f->flags[id] & 3u))) hn++;
            t->hbox = (float*) malloc ((size_t) (hn ? hn : 1) * 6 * sizeof (float));
            t->n_hbox = hn;
            unsigned k = 0;
            for (uint64_t id=0; id<f->n_nodes; id++)
                if (level[id] == hbest || (level[id] < hbest && (f->flags[id] & 3u)))
                {
                    for (int a=0; a<3; a++) { t->hbox[(size_t) a * hn + k] = f->bb_min[3*id+a]; t->hbox[(size_t) (3 + a) * hn + k] = f->bb_max[3*id+a]; }
                    k++;
                }
        }
        unsigned n = 0;
        unsigned *tmp = (unsigned*) malloc ((size_t) VOXGPU_CUT_MAX * sizeof (unsigned));
        for (uint64_t id=0; id<f->n_nodes && n < VOXGPU_CUT_MAX; id++)
            if (level[id] == best || (level[id] < best && (f->flags[id] & 3u))) tmp[n++] = (unsigned) id;
        free (level);
        t->n_cut = n; t->cut_level = best;
        if ((e = cudaMalloc (&t->cut, (size_t) (n ? n : 1) * sizeof (unsigned))) != cudaSuccess ||
            (e = cudaMemcpy (t->cut, tmp, (size_t) n * sizeof (unsigned), cudaMemcpyHostToDevice)) != cudaSuccess)
        { free (tmp); fail ("voxgpu_tree_upload: cut copy: %s", cudaGetErrorString (e)); goto bad; }
        free (tmp);
    }
    cudaFreeHost (stage_v);
    cudaStreamDestroy (s);
    return t;
bad:
    if (stage_v) cudaFreeHost (stage_v);
    if (s) cudaStreamDestroy (s);
    cudaFree (t->rec); cudaFree (t->ctr); cudaFree (t->dots); cudaFree (t->cut);
    free (t->hbox);
    free (t);
    return NULL;
}

extern "C" void voxgpu_tree_free (voxgpu_tree *t)
{
    if (!t) return;
    int cur = -1;
    cudaGetDevice (&cur);
    if (cur != t->device) cudaSetDevice (t->device);
    cudaFree (t->rec); cudaFree (t->ctr); cudaFree (t->dots); cudaFree (t->cut); cudaFree (t->cost);
    if (cur != t->device && cur >= 0) cudaSetDevice (cur);
    free (t->hbox);
    free (t);
}
extern "C" size_t voxgpu_tree_device_bytes (const voxgpu_tree *t)
{ return t ? (size_t) t->n_nodes * 48 + (size_t) t->n_dots * 16 : 0; }
extern "C" int voxgpu_tree_device (const voxgpu_tree *t) { return t ? t->device : -1; }

static TreeRef tree_ref (const voxgpu_tree *t)
{
    TreeRef r;
    r.rec = t->rec; r.ctr = t->ctr; r.dots = t->dots;
    r.vx = t->voxel[0]; r.vy = t->voxel[1]; r.vz = t->voxel[2];
    r.top_rec = r.top_ctr = NULL;
    r.ntop = 0;
    r.ntop_avail = 0;
#if VOXGPU_SMEM_TOP
    r.ntop_avail = t->n_nodes < VOXGPU_SMEM_TOP ? (unsigned) t->n_nodes : VOXGPU_SMEM_TOP;
#endif
    return r;
}

static int check_tree (const voxgpu_tree *t, uint32_t start_node, const char *who)
{
    if (!t) return fail ("%s: NULL device tree", who);
    if (start_node >= t->n_nodes) return fail ("%s: start node %u out of range", who, start_node);
    int cur = -1;
    CU (cudaGetDevice (&cur));
    if (cur != t->device) return fail ("%s: tree lives on device %d, current device is %d", who, t->device, cur);
    return 0;
}
