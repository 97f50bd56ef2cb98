// orbx_bfknn_tc.cuh - brute-force Hamming kNN-2 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) (R/src/Frame.cc:1127-1137 and the server's cross-agent matching) is a
// distance MATRIX between two sets of 256-bit vectors: hamming(a, b) = |a| + |b| - 2 <a, b> with <a, b> the dot product of the bit
// vectors.  That contraction is the one GEMM-shaped piece of the hot path, so it runs as an integer GEMM:
//   * the descriptor bits are expanded to signed bytes in shared memory in the canonical K-major no-swizzle core-matrix layout
//     (8 rows x 16 bytes per core matrix; row-group stride 2048 B, K stride 128 B) by shifts + sign-replicating PRMT, 2.9
//     instructions per 4 bits (expand_word); a CTA holds 256 queries, so every expanded train tile serves two MMA row tiles;
//   * one elected thread issues tcgen05.mma.kind::i8 (M = 128 queries, N = 128 train descriptors, K = 32 per instruction, 2 x 8
//     per train tile) with the s32 accumulators in TMEM: two stages of 2 x 128 columns, so the tensor core works on tile t+1
//     while all warps run the epilogue of tile t;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp and load), one IMAD per column builds sortable 16-bit keys
//     (distance << 7 | column), two columns per register, and packed 16x2 min / max keep the two smallest keys of the row
//     (see bf_tile_body); ties go to the lowest train index, as BFMatcher's stable order does (the rule k_bf_knn2 and the
//     oracle use).
// The popc formulation (k_bf_knn2) is bound by the 16-lane popc pipe: 5 POPC per pair.  Here a pair costs ~2.6 issue slots (1 on
// the FMA pipe, 1.5 on the ALU pipe) plus its share of the tensor pipe, which is what lifts the kernel off the popc roofline.
#pragma once
#include <stdint.h>

namespace bftc {

constexpr int M = 128;              // rows of one MMA = TMEM lanes
constexpr int MT = 2;               // query tiles per CTA: every expanded train tile is used by 256 queries
constexpr int MQ = M * MT;          // queries per CTA
constexpr int N = 128;              // train descriptors per tile = TMEM columns of one accumulator
constexpr int NT = 512;             // threads: 16 warps; warp w reads TMEM lanes 32 (w % 4) .., columns QW (w / 4) ..
constexpr int NQ = NT / 128;        // column quarters
constexpr int QW = N / NQ;          // columns per thread, tile and query tile (32: one tcgen05.ld)
constexpr int ROWB = 256;           // expanded bytes per descriptor (one 8-bit element per bit)
constexpr int A_BYTES = M * ROWB;   // 32 KB per query tile
constexpr int B_BYTES = N * ROWB;   // 32 KB per stage
constexpr int SUB = 65536;          // train rows per key space (16-bit local index)
constexpr size_t SMEM_BYTES = MT * A_BYTES + 2 * B_BYTES + 3 * N * 2 + NQ * MQ * 2 * 4 + 64;
static_assert(QW == 32 && MT * N * 2 == 512, "TMEM: 2 stages x MT accumulators of N columns = all 512 columns");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle; LBO = 128 B between the two 16-byte K chunks of one MMA, SBO = 2048 B
// between 8-row groups; bits [46, 48) = 1 (descriptor version of sm_100)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr)
{
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | ((unsigned long long)(128 >> 4) << 16) | ((unsigned long long)(2048 >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor of kind::i8: D = s32 (bits [4,6) = 2), A and B signed 8 bit (formats at [7,10) and [10,13) = 1), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr unsigned IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);

__device__ __forceinline__ void mma_i8(unsigned tmem_d, unsigned long long a_desc, unsigned long long b_desc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(unsigned taddr, int (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One 32-bit word of a descriptor -> 32 signed bytes = two 16-byte K chunks of its row, 2.9 instructions per output word:
// w << (7 - i) brings bits i, 8 + i, 16 + i, 24 + i to the sign positions of the four bytes, and PRMT in sign-replicate mode
// (selector nibbles 8 | k) turns them into 0x00 / 0xFF; (x | orv) & andv then gives {0, 1} for a query row (orv = 0, andv =
// 0x01010101) and {+1, -1} for a train row (orv = 0x01010101, andv = ~0): the accumulator becomes sum a (1 - 2 b) = |a| - 2 <a, b>
// and only |b| is left for the epilogue.  A row past the end is all-zero bytes (orv = 0, w = 0).  Output word i holds bits
// i + 8 k: a permutation of the K axis, the same for both operands, which a dot product does not see.
// The row lives at (r >> 3) * 2048 + (r & 7) * 16 of its tile (core matrices of 8 rows x 16 bytes), its 16 K chunks 128 bytes apart.
__device__ __forceinline__ unsigned sign_bytes(unsigned x)       // byte k = 0xFF if bit 7 of byte k of x is set, else 0x00
{
    unsigned r;                                                  // (the __byte_perm intrinsic masks the selector to 3 bits per nibble)
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ void expand_word(uint8_t* row_base, int word, unsigned w, unsigned orv, unsigned andv)
{
    unsigned o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = (sign_bytes(w << (7 - i)) | orv) & andv;
    uint8_t* dst = row_base + word * 256;
    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(dst + 128) = make_uint4(o[4], o[5], o[6], o[7]);
}

__device__ __forceinline__ int popc256(const uint4& lo, const uint4& hi)
{
    return __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) + __popc(hi.z) + __popc(hi.w);
}

// lexicographic insert of (d, i) into the running top-2 (d0, i0), (d1, i1); i < 0 = nothing
__device__ __forceinline__ void top2_insert(int d, int i, int& d0, int& i0, int& d1, int& i1)
{
    if (i < 0) return;
    if (i0 < 0 || d < d0 || (d == d0 && i < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
    else if (i1 < 0 || d < d1 || (d == d1 && i < i1)) { d1 = d; i1 = i; }
}

// One CTA: queries [q_first, q_first + 256) of `q` (nq rows) against train rows [t_begin, t_end) of `t`.
// Writes (idx, dist) x 2 per live query to oi / od (row stride 2 ints, indexed by the query's row in the whole set).
//
// Keys.  Inside a tile a thread owns 32 columns of its query row, so (distance << 7 | column) fits 16 bits (distance <= 256) and
// TWO columns share a register: one IMAD per column adds (|a| - 2 <a, b>) << 7 to the packed bases (|b| << 7 | column) of an even /
// odd column pair, three VIMNMX.U16x2 keep the two smallest keys of both lanes, and two independent accumulators halve the
// dependency chain.  After the tile the (at most) two survivors become 32-bit keys (distance << 16 | local train index), skipped
// outright when the tile's best distance cannot enter the row's top-2.
__device__ __forceinline__ void bf_tile_body(const uint8_t* __restrict__ q, int nq, int q_first, const uint8_t* __restrict__ t,
                                             long long t_begin, long long t_end, int idx_base, int32_t* oi, int32_t* od, uint8_t* smem)
{
    uint8_t* As = smem;                                                                    // MT query tiles
    uint8_t* Bs = smem + MT * A_BYTES;                                                     // 2 stages
    // [3][N] 16-bit key bases of a tile's columns.  THREE slots (tile % 3): the bases of tile t+2 are written while slower warps
    // may still read those of tile t in its epilogue (no barrier separates an epilogue from the next expansion)
    unsigned short* base_s = reinterpret_cast<unsigned short*>(Bs + 2 * B_BYTES);
    unsigned* keys_s = reinterpret_cast<unsigned*>(base_s + 3 * N);                        // [NQ][MQ][2] best keys of the column quarters
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(keys_s + NQ * MQ * 2); // [2] MMA-complete barriers of the two stages
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, quarter = warp >> 2;              // this thread's TMEM lane (query row of both tiles) and column range

    // ---- prologue: TMEM (all 512 columns: 2 stages x MT accumulators), barriers, the query tiles ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    {
        // 256 query rows x 8 words over 512 threads: thread -> row tid & 255, words 4 (tid >> 8) .. + 3
        const int r = tid & (MQ - 1), qi = q_first + r;
        uint4 src = make_uint4(0, 0, 0, 0);
        if (qi < nq) src = __ldg(reinterpret_cast<const uint4*>(q) + 2 * (long long)qi + (tid >> 8));
        uint8_t* rb = As + (r >> 7) * A_BYTES + ((r & 127) >> 3) * 2048 + (r & 7) * 16;
        const int w0 = (tid >> 8) * 4;
        expand_word(rb, w0, src.x, 0u, 0x01010101u); expand_word(rb, w0 + 1, src.y, 0u, 0x01010101u);
        expand_word(rb, w0 + 2, src.z, 0u, 0x01010101u); expand_word(rb, w0 + 3, src.w, 0u, 0x01010101u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    // A train tile goes through registers: its words are requested one iteration ahead (fetch_b), so that the global-memory
    // latency hides under the epilogue of the tile before, and are expanded when their stage is free (expand_b).
    // 128 train rows x 8 words over 512 threads: thread -> row tid & 127, words 2 (tid >> 7), + 1; the first thread of a row
    // also reads the whole row for |b|.
    const int br = tid & (N - 1), bpart = tid >> 7;
    uint2 nsrc = make_uint2(0, 0); uint4 nlo = make_uint4(0, 0, 0, 0), nhi = nlo; bool nvalid = false;
    auto fetch_b = [&](long long tile_first) {
        const long long g = tile_first + br;
        nvalid = g < t_end;
        nsrc = make_uint2(0, 0);
        if (nvalid) {
            nsrc = __ldg(reinterpret_cast<const uint2*>(t) + 4 * g + bpart);
            if (bpart == 0) { nlo = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g); nhi = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g + 1); }
        }
    };
    auto expand_b = [&](int stage, unsigned tile) {
        uint8_t* rb = Bs + stage * B_BYTES + (br >> 3) * 2048 + (br & 7) * 16;
        const unsigned orv = nvalid ? 0x01010101u : 0u;
        expand_word(rb, 2 * bpart, nsrc.x, orv, 0xFFFFFFFFu); expand_word(rb, 2 * bpart + 1, nsrc.y, orv, 0xFFFFFFFFu);
        if (bpart == 0)                                                        // key base of the column: |b| << 7 | column inside the quarter
            base_s[(tile % 3u) * N + br] = nvalid ? (unsigned short)((popc256(nlo, nhi) << 7) | (br & (QW - 1))) : (unsigned short)0xFFFFu;
    };
    auto issue = [&](int stage) {                                              // one thread: MT x 8 x (128 x 128 x 32) into the accumulators of `stage`
        const unsigned a0 = smem_u32(As), b0 = smem_u32(Bs + stage * B_BYTES);
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int k = 0; k < 8; k++)
                mma_i8(tmem + (stage * MT + mt) * N, smem_desc(a0 + mt * A_BYTES + k * 256), smem_desc(b0 + k * 256), k > 0);
        mma_commit(&bar[stage]);
    };
    int D0[MT], I0[MT], D1[MT], I1[MT];                                        // running top-2 over the sub-ranges (decoded)
#pragma unroll
    for (int mt = 0; mt < MT; mt++) { D0[mt] = 0; I0[mt] = -1; D1[mt] = 0; I1[mt] = -1; }
    unsigned uses = 0;                                                         // tiles issued so far (stage = uses & 1, parity = (uses >> 1) & 1)
    for (long long sub = t_begin; sub < t_end; sub += SUB) {
        const long long sub_end = sub + SUB < t_end ? sub + SUB : t_end;
        const int ntiles = (int)((sub_end - sub + N - 1) / N);
        unsigned G1[MT], G2[MT];                                               // two smallest (distance << 16 | local index) of this sub-range
#pragma unroll
        for (int mt = 0; mt < MT; mt++) { G1[mt] = 0xFFFFFFFFu; G2[mt] = 0xFFFFFFFFu; }
        fetch_b(sub);
        expand_b(uses & 1, uses);
        if (ntiles > 1) fetch_b(sub + N);
        proxy_fence(); tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); issue(uses & 1); }
        for (int tl = 0; tl < ntiles; tl++) {
            const unsigned cur = uses + tl;
            if (tl + 1 < ntiles) expand_b((cur + 1) & 1, cur + 1);
            proxy_fence(); tc_fence_before();
            __syncthreads();                       // stage (cur+1)&1: its smem is written, its accumulators were drained by the epilogue of tile cur-1
            if (tid == 0 && tl + 1 < ntiles) { tc_fence_after(); issue((cur + 1) & 1); }
            if (tl + 2 < ntiles) fetch_b(sub + (long long)(tl + 2) * N);       // in flight during the epilogue below
            mbar_wait(&bar[cur & 1], (cur >> 1) & 1);
            tc_fence_after();
            // ---- epilogue of tile cur: 32 columns of this thread's row in each query tile ----
            const uint4* kb4 = reinterpret_cast<const uint4*>(base_s + (cur % 3u) * N + quarter * QW);
            const unsigned colbase = (unsigned)(tl * N + quarter * QW);
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + ((cur & 1) * MT + mt) * N + quarter * QW;
                unsigned m1a = 0xFFFFFFFFu, m2a = 0xFFFFFFFFu, m1b = 0xFFFFFFFFu, m2b = 0xFFFFFFFFu;
                int acc[32];
                tmem_ld32(taddr, acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 kb = kb4[j >> 3];                              // packed bases of 8 columns
                    // (|a| - 2 <a, b>) << 7 onto both 16-bit lanes; mod 2^32 arithmetic, every lane ends in [0, 2^16)
                    const unsigned p0 = (unsigned)acc[j] * 128u + (unsigned)acc[j + 1] * (1u << 23) + kb.x;
                    const unsigned p1 = (unsigned)acc[j + 2] * 128u + (unsigned)acc[j + 3] * (1u << 23) + kb.y;
                    const unsigned p2 = (unsigned)acc[j + 4] * 128u + (unsigned)acc[j + 5] * (1u << 23) + kb.z;
                    const unsigned p3 = (unsigned)acc[j + 6] * 128u + (unsigned)acc[j + 7] * (1u << 23) + kb.w;
                    m2a = __vminu2(m2a, __vmaxu2(m1a, p0)); m1a = __vminu2(m1a, p0);
                    m2b = __vminu2(m2b, __vmaxu2(m1b, p1)); m1b = __vminu2(m1b, p1);
                    m2a = __vminu2(m2a, __vmaxu2(m1a, p2)); m1a = __vminu2(m1a, p2);
                    m2b = __vminu2(m2b, __vmaxu2(m1b, p3)); m1b = __vminu2(m1b, p3);
                }
                // ---- the tile's survivors -> 32-bit keys of the sub-range ----
                const unsigned n1 = __vminu2(m1a, m1b);
                const unsigned t1 = min(n1 & 0xFFFFu, n1 >> 16);
                if ((t1 >> 7) <= (G2[mt] >> 16)) {                             // otherwise nothing of this tile can enter the top-2
                    const unsigned n2 = __vminu2(__vmaxu2(m1a, m1b), __vminu2(m2a, m2b));
                    const unsigned a1 = n1 & 0xFFFFu, b1 = n1 >> 16, a2 = n2 & 0xFFFFu, b2 = n2 >> 16;
                    const unsigned t2 = min(max(a1, b1), min(a2, b2));
                    const unsigned g1 = ((t1 >> 7) << 16) | (colbase + (t1 & 127u)), g2 = ((t2 >> 7) << 16) | (colbase + (t2 & 127u));
                    G2[mt] = min(G2[mt], max(G1[mt], g1)); G1[mt] = min(G1[mt], g1);
                    G2[mt] = min(G2[mt], max(G1[mt], g2)); G1[mt] = min(G1[mt], g2);
                }
            }
        }
        uses += ntiles;
        // ---- the column quarters of a row meet in shared memory; quarter 0's thread decodes and merges ----
        tc_fence_before();
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            keys_s[(quarter * MQ + mt * M + row) * 2] = G1[mt]; keys_s[(quarter * MQ + mt * M + row) * 2 + 1] = G2[mt];
        }
        __syncthreads();
        if (quarter == 0) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                unsigned b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    const unsigned o1 = keys_s[(k * MQ + mt * M + row) * 2], o2 = keys_s[(k * MQ + mt * M + row) * 2 + 1];
                    b2 = min(b2, max(b1, o1)); b1 = min(b1, o1);
                    b2 = min(b2, max(b1, o2)); b1 = min(b1, o2);
                }
                if ((b1 >> 16) <= 256u) top2_insert((int)(b1 >> 16), (int)(sub + (b1 & 0xFFFFu)) + idx_base, D0[mt], I0[mt], D1[mt], I1[mt]);
                if ((b2 >> 16) <= 256u) top2_insert((int)(b2 >> 16), (int)(sub + (b2 & 0xFFFFu)) + idx_base, D0[mt], I0[mt], D1[mt], I1[mt]);
            }
        }
    }
    if (quarter == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int qi = q_first + mt * M + row;
            if (qi < nq) {
                oi[2ll * qi] = I0[mt]; oi[2ll * qi + 1] = I1[mt];
                od[2ll * qi] = I0[mt] >= 0 ? D0[mt] : -1; od[2ll * qi + 1] = I1[mt] >= 0 ? D1[mt] : -1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- warp-specialised form of the same search: no block-wide barrier inside the tile loop --------------------------------
//   warps 0-3   producers: expand the next train tile (one row per thread) into a shared-memory stage, write its key bases
//   warps 4-11  consumers: epilogue of an accumulator stage (warp w: TMEM lanes 32 (w % 4) .., columns 64 ((w - 4) / 4) ..)
//   warp 12     one elected thread waits for a full stage and a drained accumulator, issues the 2 x 8 MMAs, commits twice
// Four mbarrier families hand the stages around: b_full (128 producer arrivals), b_empty (tcgen05.commit: the MMAs that read the
// stage completed), acc_full (tcgen05.commit), acc_empty (256 consumer arrivals after their tcgen05.ld completed).  A thread of
// the epilogue owns 64 columns of its row in both query tiles: 128 pairs per tile and thread, which halves the per-tile overhead
// of the lock-step kernel above, and the expansion of tile t+1/t+2 overlaps the epilogue of tile t instead of preceding it.
namespace ws {
constexpr int NT = 13 * 32;
constexpr int NPROD = 128, NCONS = 256;
constexpr int HW = N / 2;                                // columns per consumer thread, tile and query tile
constexpr int BASE_SLOTS = 4;                            // key bases of tile t are read until the epilogue of t ends: tile t+4 is the first
                                                         // whose expansion provably starts after that (its stage waits MMA(t+2), which waited acc_empty(t))
constexpr size_t SMEM_BYTES = MT * A_BYTES + 2 * B_BYTES + BASE_SLOTS * N * 2 + 2 * MQ * 2 * 4 + 8 * 8 + 16;

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cons_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }    // the 256 consumer threads only

__device__ __forceinline__ void bf_tile_body(const uint8_t* __restrict__ q, int nq, int q_first, const uint8_t* __restrict__ t,
                                             long long t_begin, long long t_end, int idx_base, int32_t* oi, int32_t* od, uint8_t* smem)
{
    uint8_t* As = smem;
    uint8_t* Bs = smem + MT * A_BYTES;
    unsigned short* base_s = reinterpret_cast<unsigned short*>(Bs + 2 * B_BYTES);          // [BASE_SLOTS][N]
    unsigned* keys_s = reinterpret_cast<unsigned*>(base_s + BASE_SLOTS * N);               // [MQ][2] best keys of the upper column half
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(keys_s + MQ * 2 * 2);
    unsigned long long *b_full = bars, *b_empty = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&b_full[i], NPROD); mbar_init(&b_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the query tiles: 256 rows x 8 words, every thread helps
    for (int task = tid; task < MQ * 8; task += NT) {
        const int r = task >> 3, word = task & 7, qi = q_first + r;
        const unsigned w = qi < nq ? __ldg(reinterpret_cast<const unsigned*>(q) + 8 * (long long)qi + word) : 0u;
        expand_word(As + (r >> 7) * A_BYTES + ((r & 127) >> 3) * 2048 + (r & 7) * 16, word, w, 0u, 0x01010101u);
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;
    const long long ntiles_total = (t_end - t_begin + N - 1) / N;          // sub-ranges are multiples of N rows: tiles never straddle them

    if (warp < 4) {
        // ================= producers =================
        const int r = tid;                                                 // row of the tile
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo; bool valid = false;
        auto fetch = [&](long long tile) {
            const long long g = t_begin + tile * N + r;
            valid = tile < ntiles_total && g < t_end;
            lo = make_uint4(0, 0, 0, 0); hi = lo;
            if (valid) { lo = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g); hi = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g + 1); }
        };
        fetch(0);
        for (long long tile = 0; tile < ntiles_total; tile++) {
            const int s = (int)(tile & 1);
            const uint4 clo = lo, chi = hi; const bool cvalid = valid;
            fetch(tile + 1);                                               // in flight while this tile is expanded
            mbar_wait(&b_empty[s], (unsigned)(((tile >> 1) & 1) ^ 1));     // the MMAs that read this stage two tiles ago completed
            uint8_t* rb = Bs + s * B_BYTES + (r >> 3) * 2048 + (r & 7) * 16;
            const unsigned orv = cvalid ? 0x01010101u : 0u;
            expand_word(rb, 0, clo.x, orv, 0xFFFFFFFFu); expand_word(rb, 1, clo.y, orv, 0xFFFFFFFFu);
            expand_word(rb, 2, clo.z, orv, 0xFFFFFFFFu); expand_word(rb, 3, clo.w, orv, 0xFFFFFFFFu);
            expand_word(rb, 4, chi.x, orv, 0xFFFFFFFFu); expand_word(rb, 5, chi.y, orv, 0xFFFFFFFFu);
            expand_word(rb, 6, chi.z, orv, 0xFFFFFFFFu); expand_word(rb, 7, chi.w, orv, 0xFFFFFFFFu);
            base_s[(int)(tile % BASE_SLOTS) * N + r] = cvalid ? (unsigned short)((popc256(clo, chi) << 7) | (r & (HW - 1))) : (unsigned short)0xFFFFu;
            proxy_fence();                                                 // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(&b_full[s]);
        }
    } else if (warp == 12) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const unsigned a0 = smem_u32(As);
            for (long long tile = 0; tile < ntiles_total; tile++) {
                const int s = (int)(tile & 1);
                const unsigned ph = (unsigned)((tile >> 1) & 1);
                mbar_wait(&acc_empty[s], ph ^ 1);                          // the epilogue drained this accumulator stage
                mbar_wait(&b_full[s], ph);                                 // the stage is expanded (and its base row written)
                tc_fence_after();
                const unsigned b0 = smem_u32(Bs + s * B_BYTES);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        mma_i8(tmem + (s * MT + mt) * N, smem_desc(a0 + mt * A_BYTES + k * 256), smem_desc(b0 + k * 256), k > 0);
                mma_commit(&b_empty[s]);
                mma_commit(&acc_full[s]);
            }
        }
    } else {
        // ================= consumers =================
        const int ew = warp - 4, row = (warp & 3) * 32 + lane, half = ew >> 2;
        int D0[MT], I0[MT], D1[MT], I1[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) { D0[mt] = 0; I0[mt] = -1; D1[mt] = 0; I1[mt] = -1; }
        long long tile = 0;
        for (long long sub = t_begin; sub < t_end; sub += SUB) {
            const long long sub_end = sub + SUB < t_end ? sub + SUB : t_end;
            const int ntiles = (int)((sub_end - sub + N - 1) / N);
            unsigned G1[MT], G2[MT];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) { G1[mt] = 0xFFFFFFFFu; G2[mt] = 0xFFFFFFFFu; }
            for (int tl = 0; tl < ntiles; tl++, tile++) {
                const int s = (int)(tile & 1);
                mbar_wait(&acc_full[s], (unsigned)((tile >> 1) & 1));
                tc_fence_after();
                const uint4* kb4 = reinterpret_cast<const uint4*>(base_s + (int)(tile % BASE_SLOTS) * N + half * HW);
                const unsigned colbase = (unsigned)(tl * N + half * HW);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (s * MT + mt) * N + half * HW;
                    unsigned m1a = 0xFFFFFFFFu, m2a = 0xFFFFFFFFu, m1b = 0xFFFFFFFFu, m2b = 0xFFFFFFFFu;
#pragma unroll
                    for (int c = 0; c < HW / 32; c++) {
                        int acc[32];
                        tmem_ld32(taddr + c * 32, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            const uint4 kb = kb4[c * 4 + (j >> 3)];
                            const unsigned p0 = (unsigned)acc[j] * 128u + (unsigned)acc[j + 1] * (1u << 23) + kb.x;
                            const unsigned p1 = (unsigned)acc[j + 2] * 128u + (unsigned)acc[j + 3] * (1u << 23) + kb.y;
                            const unsigned p2 = (unsigned)acc[j + 4] * 128u + (unsigned)acc[j + 5] * (1u << 23) + kb.z;
                            const unsigned p3 = (unsigned)acc[j + 6] * 128u + (unsigned)acc[j + 7] * (1u << 23) + kb.w;
                            m2a = __vminu2(m2a, __vmaxu2(m1a, p0)); m1a = __vminu2(m1a, p0);
                            m2b = __vminu2(m2b, __vmaxu2(m1b, p1)); m1b = __vminu2(m1b, p1);
                            m2a = __vminu2(m2a, __vmaxu2(m1a, p2)); m1a = __vminu2(m1a, p2);
                            m2b = __vminu2(m2b, __vmaxu2(m1b, p3)); m1b = __vminu2(m1b, p3);
                        }
                    }
                    const unsigned n1 = __vminu2(m1a, m1b);
                    const unsigned t1 = min(n1 & 0xFFFFu, n1 >> 16);
                    if ((t1 >> 7) <= (G2[mt] >> 16)) {
                        const unsigned n2 = __vminu2(__vmaxu2(m1a, m1b), __vminu2(m2a, m2b));
                        const unsigned a1 = n1 & 0xFFFFu, b1 = n1 >> 16, a2 = n2 & 0xFFFFu, b2 = n2 >> 16;
                        const unsigned t2 = min(max(a1, b1), min(a2, b2));
                        const unsigned g1 = ((t1 >> 7) << 16) | (colbase + (t1 & 127u)), g2 = ((t2 >> 7) << 16) | (colbase + (t2 & 127u));
                        G2[mt] = min(G2[mt], max(G1[mt], g1)); G1[mt] = min(G1[mt], g1);
                        G2[mt] = min(G2[mt], max(G1[mt], g2)); G1[mt] = min(G1[mt], g2);
                    }
                }
                // accumulators read, key bases of the tile no longer needed: the stage goes back to the MMA issuer
                tc_fence_before();
                mbar_arrive(&acc_empty[s]);
            }
            // the two column halves of a row meet in shared memory (consumer-only barrier); the lower half decodes and merges
            cons_barrier();
            if (half == 1) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) { keys_s[(mt * M + row) * 2] = G1[mt]; keys_s[(mt * M + row) * 2 + 1] = G2[mt]; }
            }
            cons_barrier();
            if (half == 0) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const unsigned o1 = keys_s[(mt * M + row) * 2], o2 = keys_s[(mt * M + row) * 2 + 1];
                    unsigned b1 = G1[mt], b2 = G2[mt];
                    b2 = min(b2, max(b1, o1)); b1 = min(b1, o1);
                    b2 = min(b2, max(b1, o2)); b1 = min(b1, o2);
                    if ((b1 >> 16) <= 256u) top2_insert((int)(b1 >> 16), (int)(sub + (b1 & 0xFFFFu)) + idx_base, D0[mt], I0[mt], D1[mt], I1[mt]);
                    if ((b2 >> 16) <= 256u) top2_insert((int)(b2 >> 16), (int)(sub + (b2 & 0xFFFFu)) + idx_base, D0[mt], I0[mt], D1[mt], I1[mt]);
                }
            }
        }
        if (half == 0) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int qi = q_first + mt * M + row;
                if (qi < nq) {
                    oi[2ll * qi] = I0[mt]; oi[2ll * qi + 1] = I1[mt];
                    od[2ll * qi] = I0[mt] >= 0 ? D0[mt] : -1; od[2ll * qi + 1] = I1[mt] >= 0 ? D1[mt] : -1;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
}  // namespace ws

}  // namespace bftc
