/* libxsmm_b200 -- single-include entry point used by the reference's smallest samples (samples/hello/hello.c:11
 * includes <libxsmm_source.h>). In the reference that header pulls the whole library in as source
 * (include/libxsmm_source.h: "header-only" mode); here the implementation is the prebuilt CUDA library, so this
 * header only provides the same declarations (API + utilities + the C standard headers the samples rely on
 * having been included) and the program links with -lxsmm (libxsmm_b200/lib).
 */
#ifndef LIBXSMM_SOURCE_H
#define LIBXSMM_SOURCE_H

#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libxsmm.h"
#include "libxsmm_utils.h"

#endif /* LIBXSMM_SOURCE_H */
