/**
   @file mtree.h
   @brief M-tree of spheres (reference: src/voxtrees/mtree.h).

   The renderer's light manager keeps its shadowless lights in one of these
   (src/voxrnd/lights.c:12-15).  The tree lives on the host; vox_render hands
   the GPU a flattened copy ("lights").  Same insertion sequence,
   same tree as the reference's: which spheres a point "sees" depends on the
   (approximate) bounding spheres of the nodes above them, so the shape is part
   of the result.

   Not provided: vox_mtree_spheres_containing() (mtree.h:128-130), whose
   argument is a clang Block; the callback form below is the one the light
   manager uses.
**/
#ifndef VOXTREES_MTREE_H
#define VOXTREES_MTREE_H

#include "../voxvision.h"

/** \brief A sphere held in an M-tree (mtree.h:18-41). **/
struct vox_sphere
{
    vox_dot center;
    vox_dot color;      /**< payload: the light's colour */
    float radius;
    float sqr_radius;   /**< set by the library */
};

struct vox_mtree_node;

/** \brief Add a sphere; 1 if it was not in the tree and radius > 0 (mtree.h:62-74, mtree.c:259-272). **/
VOX_EXPORT int vox_mtree_add_sphere (struct vox_mtree_node **nodeptr, const struct vox_sphere *s);
/** \brief Remove the sphere with this centre and radius; 1 if it was there (mtree.h:76-88, mtree.c:362-382). **/
VOX_EXPORT int vox_mtree_remove_sphere (struct vox_mtree_node **nodeptr, const struct vox_sphere *s);
/** \brief Free a tree (mtree.h:90-93, mtree.c:384-395). **/
VOX_EXPORT void vox_mtree_destroy (struct vox_mtree_node *node);
/** \brief Print a tree to stdout (mtree.h:95-100, mtree.c:397-430). **/
VOX_EXPORT void vox_mtree_dump (const struct vox_mtree_node *node);
/** \brief Number of spheres in the tree (mtree.h:102-105, mtree.c:432-444). **/
VOX_EXPORT unsigned int vox_mtree_items (const struct vox_mtree_node *node);
/** \brief The leaf holding a sphere with this centre and radius, or NULL (mtree.h:107-116, mtree.c:191-223). **/
VOX_EXPORT const struct vox_mtree_node*
vox_mtree_contains_sphere (const struct vox_mtree_node *node, const struct vox_sphere *s);
/** \brief Call back for every sphere that contains the dot, in tree order (mtree.h:132-150, mtree.c:463-493). **/
VOX_EXPORT void
vox_mtree_spheres_containing_f (const struct vox_mtree_node *node, const vox_dot dot,
                                void (*callback)(const struct vox_sphere *s, void *arg),
                                void *thunk);

#endif
