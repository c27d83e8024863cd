/*
 * Entry points the reference's binding links against but which are outside this build's scope
 * (SURVEY section 8(f) "next": weighted sampling).  They exist so that the cython binding / ctypes loader resolves every symbol, and
 * they fail loudly with WHOLEMEMORY_NOT_IMPLEMENTED instead of silently doing nothing.
 */
#include "wm_internal.hpp"

/* (every entry point of the reference ABI is now implemented; this file intentionally defines nothing) */
