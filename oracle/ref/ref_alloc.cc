// ref_alloc.cc - monotonic arena for the nodes of std::list<ExtractorNode> in the reference build (TEST INFRASTRUCTURE).
//
// R/src/ORBextractor.cc:682 sorts vector<pair<int, ExtractorNode*>>: nodes holding the same number of points are ordered by
// their heap ADDRESS, so which of them is split last before the loop breaks at `size >= N` (:728-729) depends on the
// allocator.  The only well-defined reading is a monotonic allocator, where address order = creation order (SURVEY H1:
// "equal size -> later-created node first"; the oracle and the CUDA kernel implement exactly that).  This file gives the
// UNMODIFIED reference source that allocator: while an extraction runs on a thread, allocations of exactly the list-node
// size are bump-allocated from a per-thread arena that is rewound before every frame; everything else goes to malloc.
// Built into libref_orb.so only (linked -Bsymbolic so that the override stays private to that library); libref_orb_malloc.so
// is the same code on glibc malloc, used to measure how often the heap order changes the result.
#include <cstdio>
#include <cstdlib>
#include <new>
#include "ref_alloc.h"

namespace {
constexpr size_t ARENA_BYTES = 64u << 20;
struct Arena { char* base; size_t used; size_t node_bytes; bool on; };
thread_local Arena tl_arena = {nullptr, 0, 0, false};
}

void ref_arena_begin(size_t node_bytes)
{
    Arena& a = tl_arena;
    if (!a.base) { a.base = (char*)malloc(ARENA_BYTES); if (!a.base) { fprintf(stderr, "ref_alloc: out of memory\n"); abort(); } }
    a.used = 0; a.node_bytes = node_bytes; a.on = true;
}
void ref_arena_end() { tl_arena.on = false; }
size_t ref_arena_used() { return tl_arena.used; }

void* operator new(size_t n)
{
    Arena& a = tl_arena;
    if (a.on && n == a.node_bytes) {
        const size_t sz = (n + 15) & ~(size_t)15;
        if (a.used + sz > ARENA_BYTES) { fprintf(stderr, "ref_alloc: arena exhausted\n"); abort(); }
        void* p = a.base + a.used; a.used += sz; return p;
    }
    void* p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept
{
    const Arena& a = tl_arena;
    if (a.base && (char*)p >= a.base && (char*)p < a.base + ARENA_BYTES) return;      // arena memory is rewound, never freed
    free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void* operator new[](size_t n) { return operator new(n); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
