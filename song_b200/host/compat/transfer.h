/* compat/transfer.h -- stand-in for CLASS include/transfer.h: struct transfers is only passed through. */
#ifndef __TRANSFER__
#define __TRANSFER__
#include "common.h"
struct transfers { int dummy; ErrorMsg error_message; };
#endif
