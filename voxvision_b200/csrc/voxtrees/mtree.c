/*
 * mtree.c -- M-tree of spheres on the host (reference: src/voxtrees/mtree.c).
 *
 * Used by the light manager (voxrnd/lights.c).  Which lights reach a point is
 * decided by the nodes' bounding spheres, which are approximations grown by a
 * fixed sequence of fp32 operations and refreshed only along certain paths
 * (mtree.c:22-87); the tree here goes through the same sequence so that a frame
 * lit on the GPU from the flattened tree equals the reference's.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "mtree-node.h"
#include "voxtrees/geom.h"

static struct vox_mtree_node* new_node (int leaf)
{
    struct vox_mtree_node *n = calloc (1, sizeof *n);
    n->leaf = leaf;
    return n;
}

static inline const struct vox_sphere* item (const struct vox_mtree_node *n, unsigned i)
{
    return n->leaf ? &n->spheres[i] : &n->children[i]->bounding_sphere;
}

static inline int same_sphere (const struct vox_sphere *a, const struct vox_sphere *b)
{
    return vox_dot_equalp (a->center, b->center) && a->radius == b->radius;
}

/* mtree.c:22-56: grow `bs` so that it also covers s; 1 if bs changed */
static int cover (const struct vox_sphere *s, struct vox_sphere *bs)
{
    float dist = sqrtf (vox_sqr_metric (s->center, bs->center));
    float crit = (1 + (bs->radius - s->radius) / dist) / 2;
    if (crit < 0) { *bs = *s; return 1; }          /* s swallows bs */
    if (crit < 1)
    {
        bs->radius = (s->radius + bs->radius + dist) / 2;
        bs->sqr_radius = bs->radius * bs->radius;
        for (int i=0; i<3; i++)
        {
            float d = bs->center[i] - s->center[i];
            d = d * crit;
            bs->center[i] = d + s->center[i];
        }
        return 1;
    }
    return 0;                                      /* bs swallows s (or the test was NaN) */
}

/* mtree.c:58-74 */
static int refresh_bounds (struct vox_mtree_node *n)
{
    int changed = 0;
    n->bounding_sphere = *item (n, 0);
    for (unsigned i=1; i<n->num; i++) changed += cover (item (n, i), &n->bounding_sphere);
    return changed != 0;
}

/* mtree.c:76-87: upwards, until a node reports no change */
static void refresh_upwards (struct vox_mtree_node *n)
{
    while (n && refresh_bounds (n)) n = n->parent;
}

/* mtree.c:113-188.  Returns the node above the split, or the new root. */
static struct vox_mtree_node* split (struct vox_mtree_node *n)
{
    struct vox_mtree_node *parent = n->parent;
    struct vox_mtree_node *half[2] = { new_node (n->leaf), new_node (n->leaf) };
    half[0]->parent = half[1]->parent = parent;

    /* the two items farthest apart seed the halves (mtree.c:89-111) */
    vox_dot seed[2];
    float maxdist = 0;
    memset (seed, 0, sizeof seed);
    for (unsigned i=0; i<n->num; i++)
        for (unsigned j=i+1; j<n->num; j++)
        {
            float d = vox_sqr_metric (item (n, i)->center, item (n, j)->center);
            if (d > maxdist)
            {
                vox_dot_copy (seed[0], item (n, i)->center);
                vox_dot_copy (seed[1], item (n, j)->center);
                maxdist = d;
            }
        }
    for (unsigned i=0; i<n->num; i++)
    {
        float d0 = vox_sqr_metric (item (n, i)->center, seed[0]);
        float d1 = vox_sqr_metric (item (n, i)->center, seed[1]);
        struct vox_mtree_node *to = half[d0 < d1 ? 0 : 1];
        if (n->leaf) to->spheres[to->num] = n->spheres[i];
        else
        {
            to->children[to->num] = n->children[i];
            n->children[i]->parent = to;
        }
        to->num++;
    }
    refresh_bounds (half[0]);
    refresh_bounds (half[1]);

    if (!parent)
    {
        free (n);
        struct vox_mtree_node *root = new_node (0);
        root->num = 2;
        root->children[0] = half[0];
        root->children[1] = half[1];
        refresh_bounds (root);
        half[0]->parent = half[1]->parent = root;
        return root;
    }
    for (unsigned i=0; i<parent->num; i++) if (parent->children[i] == n) parent->children[i] = half[0];
    free (n);
    parent->children[parent->num++] = half[1];
    refresh_upwards (parent);
    return parent->num == MTREE_MAX_CHILDREN + 1 ? split (parent) : parent;
}

const struct vox_mtree_node*
vox_mtree_contains_sphere (const struct vox_mtree_node *n, const struct vox_sphere *s)
{
    if (!n) return NULL;
    if (!(vox_sqr_metric (s->center, n->bounding_sphere.center) < n->bounding_sphere.sqr_radius)) return NULL;
    if (n->leaf)
    {
        for (unsigned i=0; i<n->num; i++) if (same_sphere (&n->spheres[i], s)) return n;
        return NULL;
    }
    for (unsigned i=0; i<n->num; i++)
    {
        const struct vox_mtree_node *found = vox_mtree_contains_sphere (n->children[i], s);
        if (found) return found;
    }
    return NULL;
}

/* mtree.c:225-257.  Returns the lowest node the insertion left in place, or a new root. */
static struct vox_mtree_node* add (struct vox_mtree_node *n, const struct vox_sphere *s)
{
    struct vox_sphere in = *s;
    in.sqr_radius = in.radius * in.radius;
    if (!n)
    {
        n = new_node (1);
        n->bounding_sphere = in;
        n->spheres[0] = in;
        n->num = 1;
        return n;
    }
    while (!n->leaf)
    {
        /* towards the child whose bounding centre is nearest (first wins ties) */
        struct vox_mtree_node *next = n->children[0];
        float mindist = INFINITY;
        for (unsigned i=0; i<n->num; i++)
        {
            float d = vox_sqr_metric (s->center, n->children[i]->bounding_sphere.center);
            if (d < mindist) { mindist = d; next = n->children[i]; }
        }
        n = next;
    }
    n->spheres[n->num++] = in;
    refresh_upwards (n);
    return n->num == MTREE_MAX_CHILDREN + 1 ? split (n) : n;
}

int vox_mtree_add_sphere (struct vox_mtree_node **nodeptr, const struct vox_sphere *s)
{
    if (!(s->radius > 0) || vox_mtree_contains_sphere (*nodeptr, s)) return 0;
    struct vox_mtree_node *n = add (*nodeptr, s);
    if (!n->parent) *nodeptr = n;
    return 1;
}

/* mtree.c:274-330: drop an emptied leaf; a parent left with one child is replaced by that child.
   Returns the lowest surviving node above the change (NULL if the tree is gone). */
static struct vox_mtree_node* drop_leaf (struct vox_mtree_node *leaf)
{
    struct vox_mtree_node *parent = leaf->parent;
    free (leaf);
    if (!parent) return NULL;
    unsigned i = 0;
    while (i < parent->num && parent->children[i] != leaf) i++;
    if (parent->num == 2)
    {
        struct vox_mtree_node *other = parent->children[1 - i], *top = parent->parent;
        free (parent);
        other->parent = top;
        if (!top) return other;      /* `other` is the root now; its bounds are already right */
        for (i=0; i<top->num; i++) if (top->children[i] == parent) break;
        top->children[i] = other;
        parent = top;
    }
    else
    {
        parent->num--;
        memmove (&parent->children[i], &parent->children[i+1], sizeof (parent->children[0]) * (parent->num - i));
        parent->children[parent->num] = NULL;
    }
    refresh_upwards (parent);
    return parent;
}

int vox_mtree_remove_sphere (struct vox_mtree_node **nodeptr, const struct vox_sphere *s)
{
    /* mtree.c:332-382 */
    struct vox_mtree_node *leaf = (struct vox_mtree_node*) vox_mtree_contains_sphere (*nodeptr, s);
    if (!(s->radius > 0) || !leaf) return 0;
    struct vox_mtree_node *last = leaf;
    if (leaf->num == 1) last = drop_leaf (leaf);
    else
    {
        unsigned i = 0;
        while (i < leaf->num && !same_sphere (&leaf->spheres[i], s)) i++;
        memmove (&leaf->spheres[i], &leaf->spheres[i+1], sizeof (struct vox_sphere) * (leaf->num - i - 1));
        leaf->num--;
        refresh_upwards (leaf);
    }
    if (!last || !last->parent) *nodeptr = last;
    return 1;
}

void vox_mtree_destroy (struct vox_mtree_node *n)
{
    if (!n) return;
    if (!n->leaf) for (unsigned i=0; i<n->num; i++) vox_mtree_destroy (n->children[i]);
    free (n);
}

unsigned int vox_mtree_items (const struct vox_mtree_node *n)
{
    if (!n) return 0;
    if (n->leaf) return n->num;
    unsigned count = 0;
    for (unsigned i=0; i<n->num; i++) count += vox_mtree_items (n->children[i]);
    return count;
}

void vox_mtree_spheres_containing_f (const struct vox_mtree_node *n, const vox_dot dot,
                                     void (*callback)(const struct vox_sphere *s, void *arg), void *thunk)
{
    /* mtree.c:463-493 */
    if (!n) return;
    if (!(vox_sqr_metric (n->bounding_sphere.center, dot) < n->bounding_sphere.sqr_radius)) return;
    if (!n->leaf)
    {
        for (unsigned i=0; i<n->num; i++) vox_mtree_spheres_containing_f (n->children[i], dot, callback, thunk);
        return;
    }
    if (n->num == 1) { callback (&n->spheres[0], thunk); return; }   /* its bounds ARE the sphere */
    for (unsigned i=0; i<n->num; i++)
        if (vox_sqr_metric (n->spheres[i].center, dot) < n->spheres[i].sqr_radius) callback (&n->spheres[i], thunk);
}

void vox_mtree_dump (const struct vox_mtree_node *n)
{
    printf ("Node %p ", (const void*) n);
    if (!n) { printf ("empty node\n"); return; }
    printf ("%s, parent=%p\n", n->leaf ? "leaf" : "inner node", (void*) n->parent);
    printf ("Bounding sphere: <%7.4f, %7.4f, %7.4f>, %7.4f\n", n->bounding_sphere.center[0],
            n->bounding_sphere.center[1], n->bounding_sphere.center[2], n->bounding_sphere.radius);
    if (n->leaf)
    {
        printf ("Leaf content (%u spheres):\n", n->num);
        for (unsigned i=0; i<n->num; i++)
            printf ("<%7.4f, %7.4f, %7.4f>, %7.4f\n", n->spheres[i].center[0], n->spheres[i].center[1],
                    n->spheres[i].center[2], n->spheres[i].radius);
        printf ("==========\n");
        return;
    }
    printf ("Node children\n");
    for (unsigned i=0; i<n->num; i++) printf ("%p ", (void*) n->children[i]);
    printf ("\n==========\n");
    for (unsigned i=0; i<n->num; i++) vox_mtree_dump (n->children[i]);
}
