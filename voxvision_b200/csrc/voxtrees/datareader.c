/*
 * datareader.c -- raw volume loader, plain-C variant of the reference's
 * Block-typed vox_read_raw_data (src/voxtrees/datareader.c:11-85) and its
 * data-file lookup (datareader.c:87-125).
 */
#define _GNU_SOURCE
#include <sys/param.h>
#include <sys/stat.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include "store.h"

struct vox_node* vox_read_raw_data_fn (const char *filename, unsigned int dim[],
                                       unsigned int samplesize, vox_sample_test test, void *user,
                                       const char **error)
{
#define FAIL(msg) do { if (error) *error = (msg); goto done; } while (0)
    struct vox_node *tree = NULL;
    unsigned char *raw = NULL;
    vox_dot *array = NULL;
    FILE *f = fopen (filename, "rb");
    if (!f) { if (error) *error = strerror (errno); return NULL; }
    if (!test) FAIL ("No sample test given");
    if (samplesize < 1 || samplesize > 4) FAIL ("Unsupported sample size");

    size_t n = (size_t) dim[0] * dim[1] * dim[2];
    struct stat sb;
    if (fstat (fileno (f), &sb) == -1) FAIL (strerror (errno));
    /* datareader.c:39-42 */
    if ((off_t) (n * samplesize) != sb.st_size) FAIL ("Wrong size of dataset");

    raw = malloc (n * samplesize);
    if (!raw) FAIL (strerror (errno));
    if (fread (raw, samplesize, n, f) != n) FAIL ("Short read");

    /* first pass counts, so only the kept voxels are allocated */
    size_t kept = 0, pos = 0;
    for (size_t s=0; s<n; s++)
    {
        unsigned int value = 0;
        for (unsigned int l=0; l<samplesize; l++) value |= (unsigned int) raw[s*samplesize + l] << (8*l);
        kept += test (value, user) != 0;
    }
    array = vox_alloc ((kept ? kept : 1) * sizeof (vox_dot));
    if (!array) FAIL (strerror (errno));

    /* datareader.c:51-75: i outer, k inner; voxel = index * vox_voxel */
    size_t s = 0;
    for (unsigned int i=0; i<dim[0]; i++)
        for (unsigned int j=0; j<dim[1]; j++)
            for (unsigned int k=0; k<dim[2]; k++, s++)
            {
                unsigned int value = 0;
                for (unsigned int l=0; l<samplesize; l++) value |= (unsigned int) raw[s*samplesize + l] << (8*l);
                if (test (value, user))
                {
                    vox_dot_set (array[pos], i*vox_voxel[0], j*vox_voxel[1], k*vox_voxel[2]);
                    pos++;
                }
            }
    tree = vox_make_tree (array, pos);
    if (!tree && pos && error) *error = vox_last_error ();
done:
    free (array);
    free (raw);
    fclose (f);
    return tree;
#undef FAIL
}

struct range { unsigned int min; unsigned long long max; };

static int sample_in_range (unsigned int sample, void *user)
{
    const struct range *r = user;
    return sample >= r->min && (unsigned long long) sample < r->max;     /* modules/voxtrees.c:321-323 */
}

/* The range predicate evaluated on the device: upload, threshold, compact (libvoxgpu, volume.inl), then
   vox_make_tree on the resulting list -- the list the host loop above would have made, in the same order. */
static struct vox_node* read_raw_gpu (const char *filename, unsigned int dim[], unsigned int samplesize,
                                      const struct range *r, const char **error)
{
#define FAIL(msg) do { if (error) *error = (msg); goto done; } while (0)
    struct vox_node *tree = NULL;
    unsigned char *raw = NULL;
    float *dots = NULL;
    size_t kept = 0;
    FILE *f = fopen (filename, "rb");
    if (!f) { if (error) *error = strerror (errno); return NULL; }
    if (samplesize < 1 || samplesize > 4) FAIL ("Unsupported sample size");
    size_t n = (size_t) dim[0] * dim[1] * dim[2];
    struct stat sb;
    if (fstat (fileno (f), &sb) == -1) FAIL (strerror (errno));
    if ((off_t) (n * samplesize) != sb.st_size) FAIL ("Wrong size of dataset");      /* datareader.c:39-42 */
    raw = malloc (n * samplesize + 1);
    if (!raw) FAIL (strerror (errno));
    if (fread (raw, samplesize, n, f) != n) FAIL ("Short read");
    if (voxgpu_volume_extract (raw, dim, samplesize, r->min, r->max, vox_voxel, &dots, &kept))
    {
        vox_set_error ("vox_read_raw_data: %s", voxgpu_last_error ());
        FAIL (vox_last_error ());
    }
    tree = vox_make_tree ((vox_dot*) dots, kept);
    if (!tree && kept && error) *error = vox_last_error ();
done:
    voxgpu_host_free (dots);
    free (raw);
    fclose (f);
    return tree;
#undef FAIL
}

static struct vox_node* read_raw_range (const char *filename, unsigned int dim[], unsigned int samplesize,
                                        struct range r, const char **error)
{
    const char *force = getenv ("VOXTREES_LOADER");
    int gpu = force ? !strcmp (force, "gpu") : voxgpu_device_count () > 0;
    if (force && strcmp (force, "gpu") && strcmp (force, "host"))
    { if (error) *error = "VOXTREES_LOADER must be host or gpu"; return NULL; }
    if (gpu) return read_raw_gpu (filename, dim, samplesize, &r, error);
    return vox_read_raw_data_fn (filename, dim, samplesize, sample_in_range, &r, error);
}

struct vox_node* vox_read_raw_data_threshold (const char *filename, unsigned int dim[],
                                              unsigned int samplesize, unsigned int threshold,
                                              const char **error)
{
    /* sample > threshold */
    struct range r = { threshold == 0xffffffffu ? threshold : threshold + 1, threshold == 0xffffffffu ? 0 : 1ull << 32 };
    return read_raw_range (filename, dim, samplesize, r, error);
}

struct vox_node* vox_read_raw_data_range (const char *filename, unsigned int dim[],
                                          unsigned int samplesize, unsigned int min, unsigned int max,
                                          const char **error)
{
    struct range r = { min, max };
    return read_raw_range (filename, dim, samplesize, r, error);
}

static int is_regular_file (const char *path)
{
    struct stat sb;
    return !stat (path, &sb) && S_ISREG (sb.st_mode);
}

int vox_find_data_file (const char *filename, char *fullpath)
{
    /* datareader.c:105-125: as given, then $VOXVISION_DATA (default ~/.voxvision) */
    if (strlen (filename) < MAXPATHLEN && is_regular_file (strcpy (fullpath, filename))) return 1;
    const char *dir = getenv ("VOXVISION_DATA");
    char home_dir[MAXPATHLEN];
    if (!dir && getenv ("HOME") &&
        snprintf (home_dir, MAXPATHLEN, "%s/.voxvision", getenv ("HOME")) < MAXPATHLEN) dir = home_dir;
    if (dir && snprintf (fullpath, MAXPATHLEN, "%s/%s", dir, filename) < MAXPATHLEN &&
        is_regular_file (fullpath)) return 1;
    return 0;
}
