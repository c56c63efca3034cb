/* compat/binary.h -- stand-in: the hot path uses nothing from the fork's tools/binary.h */
#ifndef __BINARY__
#define __BINARY__
#include "common.h"
#endif
