#ifndef __BISPECTRA__
#define __BISPECTRA__
#include "common.h"
struct bispectra { ErrorMsg error_message; };
#endif
