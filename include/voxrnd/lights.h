/**
   @file lights.h
   @brief Light manager (reference: src/voxrnd/lights.h).

   Colours are `vox_dot` triplets (red, green, blue) in [0, 1].  With a light
   manager attached to a context (vox_context_set_light_manager), a visible
   voxel's grey texture value is multiplied, channel by channel, by
   min (1, ambient + sum of the lights that reach it) (renderer.c:41-49).  The
   sum runs on the GPU inside the render kernel, over a flattened copy of the
   manager's M-tree that is refreshed whenever the lights changed.
**/
#ifndef VOXRND_LIGHTS_H
#define VOXRND_LIGHTS_H

#include "../voxtrees/mtree.h"

struct vox_light_manager;

/** \brief A soft light bound by a sphere: full at the centre, fading linearly to nothing at the
    surface (lights.h:22-39, lights.c:45-58).  1 on success, 0 for a bad colour / radius or if a light
    with this centre and radius exists. **/
VOX_EXPORT int vox_insert_shadowless_light (struct vox_light_manager *light_manager,
                                            const vox_dot center, float radius, const vox_dot color);
/** \brief Delete the light with exactly this centre and radius (lights.h:41-48, lights.c:60-68). **/
VOX_EXPORT int vox_delete_shadowless_light (struct vox_light_manager *light_manager,
                                            const vox_dot center, float radius);
/** \brief Delete all shadowless lights (lights.h:50-53, lights.c:70-74). **/
VOX_EXPORT void vox_delete_shadowless_lights (struct vox_light_manager *light_manager);
/** \brief Number of shadowless lights (lights.h:55-58, lights.c:76-79). **/
VOX_EXPORT int vox_shadowless_lights_number (const struct vox_light_manager *light_manager);
/** \brief Light added to every visible voxel; 1 on success, 0 for a colour outside [0,1]
    (lights.h:60-67, lights.c:32-43).  (1,1,1) at creation. **/
VOX_EXPORT int vox_set_ambient_light (struct vox_light_manager *light_manager, const vox_dot color);
/** \brief A manager with no lights (lights.h:69-72, lights.c:19-26). **/
VOX_EXPORT struct vox_light_manager* vox_make_light_manager (void);
/** \brief Destroy a manager; detach it from its contexts first (lights.h:74-77, lights.c:28-30). **/
VOX_EXPORT void vox_destroy_light_manager (struct vox_light_manager *light_manager);

/** \brief Light reaching a point, evaluated on the HOST (lights.c:104-116; library-internal in the
    reference, lights.h:79-83).  For user code and tests; vox_render evaluates it on the GPU. **/
VOX_EXPORT void vox_get_light (const struct vox_light_manager *light_manager,
                               const vox_dot intersection, vox_dot light);

/* ---- additions ---- */
/** Sizes and contents of the flattened light set vox_render uploads (voxgpu_lights_flat in
    voxgpu.h): nodes in pre-order.  Call with NULL arrays to get the counts. */
VOX_EXPORT void vox_light_manager_flatten (const struct vox_light_manager *light_manager,
                                           uint32_t *n_nodes, uint32_t *n_spheres,
                                           float *node_bounds, uint32_t *node_link,
                                           float *sphere_geom, float *sphere_light);

#endif
