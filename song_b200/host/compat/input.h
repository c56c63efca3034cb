/* compat/input.h -- stand-in: the hot path uses nothing from CLASS input.h */
#ifndef __INPUT__
#define __INPUT__
#include "common.h"
#endif
