// ref_alloc.h - see ref_alloc.cc.  The malloc build defines REF_NO_ARENA and gets no-ops.
#pragma once
#include <cstddef>
#ifdef REF_NO_ARENA
static inline void ref_arena_begin(size_t) {}
static inline void ref_arena_end() {}
static inline size_t ref_arena_used() { return 0; }
#else
void ref_arena_begin(size_t node_bytes);
void ref_arena_end();
size_t ref_arena_used();
#endif
