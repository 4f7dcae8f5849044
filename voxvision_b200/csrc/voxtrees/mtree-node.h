/* mtree-node.h -- M-tree node layout (reference: src/voxtrees/mtree.h:44-58), private to
   libvoxtrees and the light manager's flattening code in libvoxrnd. */
#ifndef VOXTREES_MTREE_NODE_H
#define VOXTREES_MTREE_NODE_H

#include "voxtrees/mtree.h"

#define MTREE_MAX_CHILDREN 9

struct vox_mtree_node
{
    struct vox_sphere bounding_sphere;
    struct vox_mtree_node *parent;
    unsigned int leaf;
    unsigned int num;
    /* one slot more than the maximum: the overfull state right before a split */
    struct vox_mtree_node *children[MTREE_MAX_CHILDREN + 1];   /* inner */
    struct vox_sphere spheres[MTREE_MAX_CHILDREN + 1];          /* leaf */
};

#endif
