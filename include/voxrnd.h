/**
   @file voxrnd.h
   @brief Umbrella header of libvoxrnd (reference: src/voxrnd.h:7-14; the
   fps-control part (a clang Block API) is out of scope).
**/
#ifndef VOXRND_H
#define VOXRND_H
#ifdef __cplusplus
extern "C" {
#endif
#include "voxrnd/object.h"
#include "voxrnd/modules.h"
#include "voxrnd/camera.h"
#include "voxrnd/surface.h"
#include "voxrnd/renderer.h"
#include "voxrnd/cd.h"
#include "voxrnd/lights.h"
#include "voxrnd/screenshot.h"
#ifdef __cplusplus
}
#endif
#endif
