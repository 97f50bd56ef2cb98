// orbx_bfknn_tc.cuh - brute-force Hamming kNN-2 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) (R/src/Frame.cc:1127-1137 and the server's cross-agent matching) is a
// distance MATRIX between two sets of 256-bit vectors: hamming(a, b) = |a| + |b| - 2 <a, b> with <a, b> the dot product of the bit
// vectors.  That contraction is the one GEMM-shaped piece of the hot path, so it runs as an integer GEMM:
//   * the descriptor bits are expanded to u8 {0, 1} in shared memory in the canonical K-major no-swizzle core-matrix layout
//     (8 rows x 16 bytes per core matrix; row-group stride 2048 B, K stride 128 B), 3 ALU instructions per 4 bits;
//   * one elected thread issues tcgen05.mma.kind::i8 (M = 128 queries, N = 256 train descriptors, K = 32 per instruction, 8 per
//     tile) with the s32 accumulators in TMEM: two accumulator stages of 256 columns, so the tensor core works on tile t+1
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

constexpr int M = 128;              // queries per CTA = TMEM lanes
constexpr int N = 256;              // train descriptors per tile = TMEM columns of one accumulator stage
constexpr int NT = 512;             // threads: 16 warps; warp w reads TMEM lanes 32 (w % 4) .., columns QW (w / 4) ..
constexpr int NQ = NT / 128;        // column quarters
constexpr int QW = N / NQ;          // columns per thread and tile
constexpr int ROWB = 256;           // expanded bytes per descriptor (one 8-bit element per bit)
constexpr int A_BYTES = M * ROWB;   // 32 KB
constexpr int B_BYTES = N * ROWB;   // 64 KB per stage
constexpr int SUB = 65536;          // train rows per key space (16-bit local index)
constexpr size_t SMEM_BYTES = A_BYTES + 2 * B_BYTES + 2 * N * 2 + NQ * M * 2 * 4 + 64;

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

// 16 descriptor bits -> 16 signed bytes.  nibble * (1 + 2^7 + 2^14 + 2^21) puts bit i of the nibble at bit 8 i (the four partial
// products do not overlap: no carries), the mask drops the rest: bytes {0, 1}.  Queries keep that (mul = 1, add = 0); train rows
// become 1 - 2 b = {+1, -1} (mul = 0xFE, add = 0x01010101: 0x01 * 0xFE stays inside its byte), so that the accumulator is
// sum a (1 - 2 b) = |a| - 2 <a, b> and only |b| is left for the epilogue; rows past the end are all-zero bytes (mul = add = 0).
__device__ __forceinline__ uint4 expand16(unsigned h, unsigned mul, unsigned add)
{
    uint4 o;
    o.x = (((h & 0xFu) * 0x00204081u) & 0x01010101u) * mul + add;
    o.y = ((((h >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * mul + add;
    o.z = ((((h >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) * mul + add;
    o.w = ((((h >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) * mul + add;
    return o;
}

// NC consecutive chunks (16 bits each, NC / 2 words `w`) starting at chunk c0 of one descriptor row -> the core-matrix layout:
// row r lives at (r >> 3) * 2048 + (r & 7) * 16, its 16 K chunks 128 bytes apart
template <int NC>
__device__ __forceinline__ void expand_chunks(uint8_t* tile, int r, int c0, const unsigned (&w)[NC / 2], unsigned mul, unsigned add)
{
    uint8_t* dst = tile + (r >> 3) * 2048 + (r & 7) * 16 + c0 * 128;
#pragma unroll
    for (int i = 0; i < NC; i++) *reinterpret_cast<uint4*>(dst + i * 128) = expand16((w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu, mul, add);
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

// One CTA: queries [q_first, q_first + 128) of `q` (nq rows) against train rows [t_begin, t_end) of `t`.
// Writes (idx, dist) x 2 per live query to oi / od (row stride 2 ints, indexed by the query's row in the whole set).
//
// Keys.  Inside a tile a thread owns QW (<= 128) columns of its query row, so (distance << 7 | column) fits 16 bits (distance <=
// 256) and TWO columns share a register: one IMAD per column adds (|a| - 2 <a, b>) << 7 to the packed bases (|b| << 7 | column) of
// an even / odd column pair, three VIMNMX.U16x2 keep the two smallest keys of both lanes, and two independent accumulators halve
// the dependency chain.  After the tile the (at most) two survivors become 32-bit keys (distance << 16 | local train index),
// skipped outright when the tile's best distance cannot enter the row's top-2.
__device__ __forceinline__ void bf_tile_body(const uint8_t* __restrict__ q, int nq, int q_first, const uint8_t* __restrict__ t,
                                             long long t_begin, long long t_end, int idx_base, int32_t* oi, int32_t* od, uint8_t* smem)
{
    uint8_t* As = smem;
    uint8_t* Bs = smem + A_BYTES;                                                          // 2 stages
    unsigned short* base_s = reinterpret_cast<unsigned short*>(smem + A_BYTES + 2 * B_BYTES);   // [2][N] 16-bit key bases of the stage's columns
    unsigned* keys_s = reinterpret_cast<unsigned*>(base_s + 2 * N);                        // [NQ][M][2] best keys of the column quarters
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(keys_s + NQ * M * 2);  // [2] MMA-complete barriers of the accumulator stages
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, quarter = warp >> 2;              // this thread's TMEM lane (query row) and column range

    // ---- prologue: TMEM (all 512 columns: 2 accumulator stages), barriers, the query tile ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    {
        const int r = tid & (M - 1), qi = q_first + r;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (qi < nq) { lo = __ldg(reinterpret_cast<const uint4*>(q) + 2 * (long long)qi); hi = __ldg(reinterpret_cast<const uint4*>(q) + 2 * (long long)qi + 1); }
        static_assert(NQ == 4, "the query tile is expanded 4 chunks (2 words) per thread");
        const int part = tid >> 7;                                             // chunks 4 part .. 4 part + 3 = words 2 part, 2 part + 1
        const uint4 src = (part & 2) ? hi : lo;
        const unsigned w2[2] = {(part & 1) ? src.z : src.x, (part & 1) ? src.w : src.y};
        expand_chunks<4>(As, r, part * 4, w2, 1u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    auto expand_b = [&](long long tile_first, int stage) {
        const int r = tid & (N - 1);
        const long long g = tile_first + r;                                    // train row; NT / N threads share it
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        const bool valid = g < t_end;
        if (valid) { lo = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g); hi = __ldg(reinterpret_cast<const uint4*>(t) + 2 * g + 1); }
        static_assert(NT == 2 * N, "a train row is expanded by two threads, 8 chunks (4 words) each");
        const uint4 src = (tid >= N) ? hi : lo;
        const unsigned w4[4] = {src.x, src.y, src.z, src.w};
        expand_chunks<8>(Bs + stage * B_BYTES, r, (tid / N) * 8, w4, valid ? 0xFEu : 0u, valid ? 0x01010101u : 0u);
        if (tid < N) base_s[stage * N + r] = valid ? (unsigned short)((popc256(lo, hi) << 7) | (r & (QW - 1))) : (unsigned short)0xFFFFu;
    };
    auto issue = [&](int stage) {                                              // one thread: 8 x (128 x 256 x 32) into accumulator `stage`
        const unsigned a0 = smem_u32(As), b0 = smem_u32(Bs + stage * B_BYTES);
#pragma unroll
        for (int k = 0; k < 8; k++) mma_i8(tmem + stage * N, smem_desc(a0 + k * 256), smem_desc(b0 + k * 256), k > 0);
        mma_commit(&bar[stage]);
    };

    int D0 = 0, I0 = -1, D1 = 0, I1 = -1;                                      // running top-2 over the sub-ranges (decoded)
    unsigned uses = 0;                                                         // tiles issued so far (stage = uses & 1, parity = (uses >> 1) & 1)
    for (long long sub = t_begin; sub < t_end; sub += SUB) {
        const long long sub_end = sub + SUB < t_end ? sub + SUB : t_end;
        const int ntiles = (int)((sub_end - sub + N - 1) / N);
        unsigned G1 = 0xFFFFFFFFu, G2 = 0xFFFFFFFFu;                           // two smallest (distance << 16 | local index) of this sub-range
        expand_b(sub, uses & 1);
        proxy_fence(); tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); issue(uses & 1); }
        for (int tl = 0; tl < ntiles; tl++) {
            const unsigned cur = uses + tl;
            if (tl + 1 < ntiles) expand_b(sub + (long long)(tl + 1) * N, (cur + 1) & 1);
            proxy_fence(); tc_fence_before();
            __syncthreads();                       // stage (cur+1)&1: its smem is written, its accumulator was drained by the epilogue of tile cur-1
            if (tid == 0 && tl + 1 < ntiles) { tc_fence_after(); issue((cur + 1) & 1); }
            mbar_wait(&bar[cur & 1], (cur >> 1) & 1);
            tc_fence_after();
            // ---- epilogue of tile cur: QW columns of this thread's row ----
            const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (cur & 1) * N + quarter * QW;
            const uint4* kb4 = reinterpret_cast<const uint4*>(base_s + (cur & 1) * N + quarter * QW);
            unsigned m1a = 0xFFFFFFFFu, m2a = 0xFFFFFFFFu, m1b = 0xFFFFFFFFu, m2b = 0xFFFFFFFFu;
#pragma unroll
            for (int c = 0; c < QW / 32; c++) {
                int acc[32];
                tmem_ld32(taddr + c * 32, acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 kb = kb4[c * 4 + (j >> 3)];                    // packed bases of 8 columns
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
            }
            // ---- the tile's survivors -> 32-bit keys of the sub-range ----
            const unsigned n1 = __vminu2(m1a, m1b);
            const unsigned t1 = min(n1 & 0xFFFFu, n1 >> 16);
            if ((t1 >> 7) <= (G2 >> 16)) {                                     // otherwise nothing of this tile can enter the top-2
                const unsigned n2 = __vminu2(__vmaxu2(m1a, m1b), __vminu2(m2a, m2b));
                const unsigned a1 = n1 & 0xFFFFu, b1 = n1 >> 16, a2 = n2 & 0xFFFFu, b2 = n2 >> 16;
                const unsigned t2 = min(max(a1, b1), min(a2, b2));
                const unsigned colbase = (unsigned)(tl * N + quarter * QW);
                const unsigned g1 = ((t1 >> 7) << 16) | (colbase + (t1 & 127u)), g2 = ((t2 >> 7) << 16) | (colbase + (t2 & 127u));
                G2 = min(G2, max(G1, g1)); G1 = min(G1, g1);
                G2 = min(G2, max(G1, g2)); G1 = min(G1, g2);
            }
        }
        uses += ntiles;
        // ---- the column quarters of a row meet in shared memory; quarter 0's thread decodes and merges ----
        tc_fence_before();
        __syncthreads();
        keys_s[(quarter * M + row) * 2] = G1; keys_s[(quarter * M + row) * 2 + 1] = G2;
        __syncthreads();
        if (quarter == 0) {
            unsigned b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const unsigned o1 = keys_s[(k * M + row) * 2], o2 = keys_s[(k * M + row) * 2 + 1];
                b2 = min(b2, max(b1, o1)); b1 = min(b1, o1);
                b2 = min(b2, max(b1, o2)); b1 = min(b1, o2);
            }
            if ((b1 >> 16) <= 256u) top2_insert((int)(b1 >> 16), (int)(sub + (b1 & 0xFFFFu)) + idx_base, D0, I0, D1, I1);
            if ((b2 >> 16) <= 256u) top2_insert((int)(b2 >> 16), (int)(sub + (b2 & 0xFFFFu)) + idx_base, D0, I0, D1, I1);
        }
    }
    if (quarter == 0 && q_first + row < nq) {
        const long long o = 2ll * (q_first + row);
        oi[o] = I0; oi[o + 1] = I1;
        od[o] = I0 >= 0 ? D0 : -1; od[o + 1] = I1 >= 0 ? D1 : -1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace bftc
