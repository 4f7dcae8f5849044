/**
   @file datareader.h
   @brief Raw volume reader (reference: src/voxtrees/datareader.h).

   The reference's vox_read_raw_data() takes a clang Block as the sample
   predicate (datareader.c:11-13), which plain C cannot express.  This is the
   ABI-adjacent variant: a function pointer + user pointer, or NULL to use the
   closed interval [min, max] test of the Lua binding
   (src/voxengine/modules/voxtrees.c:306-342).
**/
#ifndef VOXTREES_DATAREADER_H
#define VOXTREES_DATAREADER_H

#include "params.h"

struct vox_node;
typedef int (*vox_sample_test) (unsigned int sample, void *user);

/**
   \brief Build a tree from a raw density file of dim[0]*dim[1]*dim[2] samples,
   `samplesize` little-endian bytes each, x-major (datareader.c:51-76).
   \return tree or NULL; on error `*error` points at a static string.
**/
VOX_EXPORT struct vox_node* vox_read_raw_data_fn (const char *filename, unsigned int dim[],
                                                  unsigned int samplesize,
                                                  vox_sample_test test, void *user,
                                                  const char **error);

/** Same, keeping samples with `sample > threshold` (the demo's predicate, demo/main.c:169). **/
VOX_EXPORT struct vox_node* vox_read_raw_data_threshold (const char *filename, unsigned int dim[],
                                                         unsigned int samplesize,
                                                         unsigned int threshold,
                                                         const char **error);

/**
   \brief Same, keeping samples with min <= sample < max, the predicate of the reference's Lua binding
   (`voxtrees.read_raw_data_ranged`, src/voxengine/modules/voxtrees.c:306-342).

   This and vox_read_raw_data_threshold() extract the voxel list ON THE DEVICE when a CUDA device is
   present (libvoxgpu: volume upload, threshold, stable stream compaction in the reference's x-outer /
   z-inner order, voxgpu_volume_extract) and hand it to vox_make_tree(); without a device the host loop of
   vox_read_raw_data_fn() does it.  $VOXTREES_LOADER=host|gpu forces one of the two (tests compare them).
**/
VOX_EXPORT struct vox_node* vox_read_raw_data_range (const char *filename, unsigned int dim[],
                                                     unsigned int samplesize,
                                                     unsigned int min, unsigned int max,
                                                     const char **error);

/** \brief Find a data file in the usual places (datareader.c:105-125). **/
VOX_EXPORT int vox_find_data_file (const char *filename, char *fullpath);

#endif
