/**
   @file voxtrees.h
   @brief Umbrella header of libvoxtrees (reference: src/voxtrees.h:8-12).
**/
#ifndef VOXTREES_H
#define VOXTREES_H
#ifdef __cplusplus
extern "C" {
#endif
#include "voxtrees/params.h"
#include "voxtrees/tree.h"
#include "voxtrees/search.h"
#include "voxtrees/geom.h"
#include "voxtrees/datareader.h"
#include "voxtrees/batch.h"
#include "voxtrees/mtree.h"
#ifdef __cplusplus
}
#endif
#endif
