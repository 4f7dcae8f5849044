/**
   @file screenshot.h
   @brief BMP screenshots (reference: src/voxrnd/screenshot.h).
**/
#ifndef VOXRND_SCREENSHOT_H
#define VOXRND_SCREENSHOT_H

#include "renderer.h"

/**
   \brief Save the context's surface as `<dirname>/screenshot<N>.bmp`, N being the first
   number not taken yet (screenshot.h:22, screenshot.c:7-36).  `dirname` NULL means
   `$VOXVISION_DATA` (or the current directory).  Call after vox_redraw().  The directory is
   not created.  The file is an uncompressed 24-bit BMP; without SDL there is no
   SDL_GetError(), vox_last_error() carries the reason of a failure.
   \return 1 on success, 0 on error.
**/
VOX_EXPORT int vox_screenshot (const struct vox_rnd_ctx *context, const char *dirname);

#endif
