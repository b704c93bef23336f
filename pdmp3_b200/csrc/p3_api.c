/* p3_api.c -- placeholder translation unit; the streaming API (pdmp3_new ... pdmp3_getformat)
 * is implemented in the next milestone. */
#include "../../include/pdmp3.h"
