// Host-side description of a multi-pass (four-step) transform: factorisation and per-pass tile
// arguments.  Pure arithmetic, shared by the library (capi.cu) and the CPU emulator tests.
#pragma once
#include <cstdint>
#include <vector>

#include "large_kernels.cuh"

namespace cfb
{
// adjacent transforms per tile = contiguous complex elements on the strided side of a pass: 16 (one full 128-byte
// line per element row) for the strided passes whenever the CTA stays at <= 512 threads, 8 (64 bytes) for
// 1024-point passes and for the contiguous-row pass
constexpr int kTileCMax = 16;
inline int& tile_c_override() { static int v = 0; return v; }       // tuning hooks (0 = policy below)
inline int& tile_c_jfast_override() { static int v = 0; return v; } // ... for the contiguous-row (last) pass only
// tuning hook "tile_tma": 1 = the persistent tensor-map TMA tile kernel (tile_tma_kernel) where a pass can be expressed as
// tensor maps, 0 = tile_fft_kernel everywhere
inline int& tile_tma_mode() { static int v = 0; return v; }
// tuning hook "tile_r": 1 = 32 complex points per thread in the 512- / 1024-point tile passes (two Stockham stages instead of three)
inline int& tile_radix32() { static int v = 0; return v; }
inline int& tile_stream_mode() { static int v = 0; return v; } // tuning hook "tile_stream": bit 0 / 1 = evict-first loads / stores in passes that carry no policy
inline int& tile_pf_distance() { static int v = 0; return v; } // tuning hook "tile_pf": tensor-map L2 prefetch distance of the tile passes in tiles (0 = off)
inline int tile_c (int logL, bool jfast = false)
{
    const int ov = (jfast && tile_c_jfast_override() != 0) ? tile_c_jfast_override() : tile_c_override();
    if (ov != 0)
        return (ov == 16 && logL > 9) ? 8 : ov;
    return (logL <= 9 && ! jfast) ? 16 : 8; // measured: the contiguous-row pass is faster with 8 (profiles/r01_large_tile_width.txt)
}
constexpr int kMinTileLog = 6;   // tile transforms are 64 .. 1024 points
constexpr int kMaxTileLog = 10;
constexpr int kMaxLargeLog = 28; // 2^28 complex points (2 GiB) is the largest single transform

struct LargeFactors
{
    int l1 = 0, l2 = 0, l3 = 0; // log2 of the pass lengths; l2 == 0: two-pass plan
    int passes() const { return l2 == 0 ? 2 : 3; }
};

// complex length 2^n, n > kMaxLogM (14)
inline LargeFactors choose_factors (int n)
{
    LargeFactors f;
    if (n <= 2 * kMaxTileLog)
    {
        f.l3 = (n + 1) / 2;
        f.l1 = n - f.l3;
    }
    else
    {
        f.l3 = (n + 2) / 3;
        f.l2 = (n - f.l3 + 1) / 2;
        f.l1 = n - f.l3 - f.l2;
    }
    return f;
}

struct TilePass
{
    int C;             // transforms per tile (8 or 16)
    int logL;          // transform length of this pass
    bool load_j_fast;  // contiguous-row pass (the last one)
    int uio;           // 0 natural order, 1 unordered input (first pass, inverse), 2 unordered output (last pass, forward)
    int which;         // 0 = pass A (length L1), 1 = pass B (L2), 2 = pass C (L3)
    TileArgs args;     // in / out / twiddle pointers are filled in by the caller
};

// pass list for a complex transform of 2^n points; buffers: src -> (tmp ... tmp) -> dst
// tw_scale: the big twiddle tables may belong to a longer length (real plans keep ONE table for 2M): W_(2^n)^e = W_table^(e * tw_scale)
inline int build_tile_passes (int n, const LargeFactors& f, TilePass (&p)[3], unsigned tw_scale = 1)
{
    const long long N = 1LL << n, L1 = 1LL << f.l1, L2 = 1LL << f.l2, L3 = 1LL << f.l3;
    int np = 0;
    {   // pass A: columns of the [L1][N/L1] view
        TilePass& a = p[np++];
        a = {};
        a.logL = f.l1;
        a.C = tile_c (f.l1);
        const int kTileC = a.C;
        a.load_j_fast = false;
        const long long S1 = N / L1;
        a.args.gdiv = (int) (S1 / kTileC);
        a.args.ntiles = (int) (S1 / kTileC);
        a.args.in_g_hi = a.args.out_g_hi = 0;
        a.args.in_g_lo = a.args.out_g_lo = kTileC;
        a.args.in_tstride = a.args.out_tstride = 1;
        a.args.in_estride = a.args.out_estride = S1;
        a.args.tw_mult = tw_scale;
        a.args.batch = 1;
        a.args.in_split_log = 31;
        a.args.peer_row_log = -1;
        a.args.logN = n;
    }
    if (f.l2 != 0)
    {   // pass B: for every row k1, columns of its [L2][L3] view
        TilePass& b = p[np++];
        b = {};
        b.which = 1;
        b.logL = f.l2;
        b.C = tile_c (f.l2);
        const int kTileC = b.C;
        b.load_j_fast = false;
        b.args.gdiv = (int) (L3 / kTileC);
        b.args.ntiles = (int) (L1 * L3 / kTileC);
        b.args.in_g_hi = b.args.out_g_hi = L2 * L3;
        b.args.in_g_lo = b.args.out_g_lo = kTileC;
        b.args.in_tstride = b.args.out_tstride = 1;
        b.args.in_estride = b.args.out_estride = L3;
        b.args.tw_mult = (unsigned) L1 * tw_scale;
        b.args.batch = 1;
        b.args.in_split_log = 31;
        b.args.peer_row_log = -1;
        b.args.logN = n;
    }
    {   // pass C: contiguous rows (k1, k2), written transposed to k1 + L1 (k2 + L2 k3)
        TilePass& c = p[np++];
        c = {};
        c.which = 2;
        c.logL = f.l3;
        c.C = tile_c (f.l3, true);
        const int kTileC = c.C;
        c.load_j_fast = true;
        c.args.gdiv = (int) (L1 / kTileC);
        c.args.ntiles = (int) (L1 * L2 / kTileC);
        c.args.in_g_hi = L3;
        c.args.in_g_lo = kTileC * L2 * L3;
        c.args.in_tstride = L2 * L3;
        c.args.in_estride = 1;
        c.args.out_g_hi = L1;
        c.args.out_g_lo = kTileC;
        c.args.out_tstride = 1;
        c.args.out_estride = L1 * L2;
        c.args.tw_mult = 0;
        c.args.batch = 1;
        c.args.in_split_log = 31;
        c.args.peer_row_log = -1;
        c.args.logN = n;
    }
    return np;
}

// Distributed four-step over `world` ranks (three-pass plans only).  Data contracts, N = L1 L2 L3, S1 = L2 L3:
//   phase 0 input  : rank r holds the column block  A_r[n1][c] = x[n1 S1 + r S1/world + c]   ([L1][S1/world])
//   phase 0        : pass A on the local columns, in place layout ([k1][c]); rows k1 of block h are the
//                    contiguous chunk sent to rank h by the all-to-all
//   phase 1 input  : exchange layout [world][L1/world][S1/world] (chunk g = columns of rank g)
//   phase 1        : pass B, reading the exchange layout, writing natural rows [L1/world][S1]
//   phase 2        : pass C, writing "transposed-out": out[q][k1_local] = X[(r L1/world + k1_local) + L1 q]
inline bool build_dist_phase (int n, const LargeFactors& f, int phase, int rank, int world, TilePass& p)
{
    if (f.l2 == 0 || world < 1 || (world & (world - 1)) != 0)
        return false;
    const long long N = 1LL << n, L1 = 1LL << f.l1, L2 = 1LL << f.l2, L3 = 1LL << f.l3, S1 = L2 * L3;
    const int kTileC = tile_c (phase == 0 ? f.l1 : (phase == 1 ? f.l2 : f.l3), phase == 2);
    if (L1 / world < kTileCMax || L2 / world < 1 || S1 / world < kTileCMax)
        return false;
    int wl = 0;
    while ((1 << wl) < world)
        ++wl;
    p = {};
    p.C = kTileC;
    p.args.batch = 1;
    p.args.in_split_log = 31;
    p.args.peer_row_log = -1;
    (void) N;
    if (phase == 0)
    {
        const long long cols = S1 / world;
        p.logL = f.l1;
        p.load_j_fast = false;
        p.args.gdiv = (int) (cols / kTileC);
        p.args.ntiles = (int) (cols / kTileC);
        p.args.in_g_lo = p.args.out_g_lo = kTileC;
        p.args.in_tstride = p.args.out_tstride = 1;
        p.args.in_estride = p.args.out_estride = cols;
        p.args.tw_mult = 1;
        p.args.tw_c_base = (unsigned) (rank * cols);
        // peer-store variant (the caller enables it by setting peer_row_log and peer_out): rank h owns rows
        // [h L1/world, (h+1) L1/world) and receives them as chunk `rank` of its [world][L1/world][cols] buffer
        p.args.peer_chunk_off = (long long) rank * (L1 / world) * cols;
    }
    else if (phase == 1)
    {
        const long long rows = L1 / world, cols = S1 / world;
        p.logL = f.l2;
        p.load_j_fast = false;
        p.args.gdiv = (int) (L3 / kTileC);
        p.args.ntiles = (int) (rows * L3 / kTileC);
        p.args.in_g_hi = cols;             // row k1_local inside every chunk
        p.args.in_g_lo = kTileC;
        p.args.in_tstride = 1;
        p.args.in_estride = L3;
        p.args.in_split_log = f.l2 - wl;   // n2 values per chunk
        p.args.in_chunk_stride = rows * cols;
        p.args.out_g_hi = S1;
        p.args.out_g_lo = kTileC;
        p.args.out_tstride = 1;
        p.args.out_estride = L3;
        p.args.tw_mult = (unsigned) L1;
    }
    else
    {
        const long long rows = L1 / world;
        p.logL = f.l3;
        p.load_j_fast = true;
        p.args.gdiv = (int) (rows / kTileC);
        p.args.ntiles = (int) (rows * L2 / kTileC);
        p.args.in_g_hi = L3;
        p.args.in_g_lo = kTileC * S1;
        p.args.in_tstride = S1;
        p.args.in_estride = 1;
        p.args.out_g_hi = rows;
        p.args.out_g_lo = kTileC;
        p.args.out_tstride = 1;
        p.args.out_estride = rows * L2;
        p.args.tw_mult = 0;
    }
    return true;
}


// ---------------------------------------------------------------------------------------------
// L2-chunked schedules.  B200 has 126 MB of L2: instead of running every pass over the whole array (each pass = one
// HBM read + one HBM write of the array), consecutive passes are run back to back on CHUNKS whose intermediate lives in
// a small ring buffer that stays L2-resident:
//   two-pass plans   (2^15 .. 2^20 points, batched): chunk = a few whole transforms;  A: src -> ring, C: ring -> dst
//                     => one HBM read + one HBM write per transform instead of two of each
//   three-pass plans (2^21 .. 2^28 points): pass A over the whole array (src -> s1), then chunks of k1-rows:
//                     B: s1 rows -> ring, C: ring -> dst   => two HBM reads + two writes instead of three of each
// (a chunk of k1-rows is the smallest unit on which B and C compose: pass C needs every column group of its C adjacent
// rows finished).  Chunks alternate over `lanes` helper streams, each with its own ring slot, so that the B of one chunk
// overlaps the C of the previous one and launch tails are filled; ring accesses carry an evict_last L2 policy, the
// HBM-facing streams evict_first.  The reference does all of this through one work buffer in cache-sized FFTPACK passes
// (/root/reference/simd/chowdsp_fft_impl_avx.cpp:430-490, :1848-1935); here "cache-sized" means L2-sized.
// ---------------------------------------------------------------------------------------------
struct LargeLaunch
{
    TilePass pass;
    int lane; // helper stream / ring slot index; -1 = the caller's stream, before the fork
};

struct LargeBuffers
{
    const float2* src;
    float2* dst;
    long long src_bs, dst_bs; // float2 between transforms of the batch (src / dst)
    float2* s1;               // full-size scratch, batch * 2^n float2 (classic schedule and pass A of chunked three-pass plans)
    float2* ring;             // lanes * ring_lane_elems float2 (chunked schedules), else nullptr
    long long ring_lane_elems;
};

// rows of pass B / C per chunk (three-pass plans): a multiple of the last pass's tile width, at most L1
inline long long chunk_rows (int n, const LargeFactors& f, int c_last, long long chunk_elems)
{
    const long long L1 = 1LL << f.l1, row = 1LL << (n - f.l1);
    long long rows = chunk_elems / row;
    rows -= rows % c_last;
    if (rows < c_last)
        rows = c_last;
    return rows > L1 ? L1 : rows;
}
// elements one ring slot needs for chunk size `chunk_elems` (0 = classic schedule, no ring)
inline long long ring_elems_needed (int n, const LargeFactors& f, int batch, long long chunk_elems)
{
    if (chunk_elems <= 0)
        return 0;
    const long long N = 1LL << n;
    if (f.l2 == 0)
    {
        long long k = chunk_elems / N;
        k = k < 1 ? 1 : (k > batch ? batch : k);
        return k * N;
    }
    return chunk_rows (n, f, tile_c (f.l3, true), chunk_elems) * (N >> f.l1);
}

// uio_in / uio_out: the transform reads / writes the unordered complex layout (logW lanes) -- folded into the first / last pass
inline void build_large_schedule (int n, const LargeFactors& f, int batch, const LargeBuffers& bufs, unsigned tw_scale, bool uio_in, bool uio_out, int logW,
                                  long long chunk_elems, int lanes, bool policies, std::vector<LargeLaunch>& out)
{
    TilePass p[3];
    const int np = build_tile_passes (n, f, p, tw_scale);
    const long long N = 1LL << n;
    if (uio_in)
    {
        p[0].uio = 1;
        p[0].args.unord_logW = logW;
    }
    if (uio_out)
    {
        p[np - 1].uio = 2;
        p[np - 1].args.unord_logW = logW;
    }
    out.clear();
    if (chunk_elems <= 0 || bufs.ring == nullptr)
    {
        // classic: every pass over the whole batch, src -> s1 -> ... -> dst
        for (int i = 0; i < np; ++i)
        {
            LargeLaunch l { p[i], -1 };
            l.pass.args.in = i == 0 ? bufs.src : bufs.s1;
            l.pass.args.in_bstride = i == 0 ? bufs.src_bs : N;
            l.pass.args.out = i == np - 1 ? bufs.dst : bufs.s1;
            l.pass.args.out_bstride = i == np - 1 ? bufs.dst_bs : N;
            l.pass.args.batch = batch;
            out.push_back (l);
        }
        return;
    }
    const int keep = policies ? POLICY_KEEP : POLICY_NORMAL, strm = policies ? POLICY_STREAM : POLICY_NORMAL;
    if (np == 2)
    {
        const long long k = bufs.ring_lane_elems / N; // transforms per chunk
        int lane = k >= batch ? -1 : 0;               // a single chunk stays on the caller's stream
        for (long long b0 = 0; b0 < batch; b0 += k, lane = lane < 0 ? -1 : (lane + 1) % lanes)
        {
            const int nb = (int) (batch - b0 < k ? batch - b0 : k);
            float2* slot = bufs.ring + (long long) (lane < 0 ? 0 : lane) * bufs.ring_lane_elems;
            LargeLaunch a { p[0], lane }, c { p[1], lane };
            a.pass.args.in = bufs.src + b0 * bufs.src_bs;
            a.pass.args.in_bstride = bufs.src_bs;
            a.pass.args.out = slot;
            a.pass.args.out_bstride = N;
            a.pass.args.batch = nb;
            a.pass.args.in_policy = strm;
            a.pass.args.out_policy = keep;
            c.pass.args.in = slot;
            c.pass.args.in_bstride = N;
            c.pass.args.out = bufs.dst + b0 * bufs.dst_bs;
            c.pass.args.out_bstride = bufs.dst_bs;
            c.pass.args.batch = nb;
            c.pass.args.in_policy = keep;
            c.pass.args.out_policy = strm;
            out.push_back (a);
            out.push_back (c);
        }
        return;
    }
    // three passes: A over everything on the caller's stream, then (B, C) per chunk of k1-rows
    {
        LargeLaunch a { p[0], -1 };
        a.pass.args.in = bufs.src;
        a.pass.args.in_bstride = bufs.src_bs;
        a.pass.args.out = bufs.s1;
        a.pass.args.out_bstride = N;
        a.pass.args.batch = batch;
        out.push_back (a);
    }
    const long long L1 = 1LL << f.l1, L2 = 1LL << f.l2, S1 = N >> f.l1;
    const long long rows = bufs.ring_lane_elems / S1;
    int lane = (batch == 1 && rows >= L1) ? -1 : 0;
    for (int b = 0; b < batch; ++b)
        for (long long r0 = 0; r0 < L1; r0 += rows, lane = lane < 0 ? -1 : (lane + 1) % lanes)
        {
            const long long nr = L1 - r0 < rows ? L1 - r0 : rows;
            float2* slot = bufs.ring + (long long) (lane < 0 ? 0 : lane) * bufs.ring_lane_elems;
            LargeLaunch bb { p[1], lane }, cc { p[2], lane };
            // B on rows [r0, r0 + nr): tile g -> (row g / gdiv, column group g % gdiv); the ring slot starts at row r0
            bb.pass.args.in = bufs.s1 + (long long) b * N;
            bb.pass.args.in_bin0 = r0 * S1;
            bb.pass.args.out = slot;
            bb.pass.args.out_bin0 = 0;
            bb.pass.args.ntiles = (int) (nr * bb.pass.args.gdiv);
            bb.pass.args.batch = 1;
            bb.pass.args.in_policy = strm;
            bb.pass.args.out_policy = keep;
            // C on the same rows: tile g -> (k2 = g / gdiv', k1 group g % gdiv'), gdiv' = nr / C
            cc.pass.args.in = slot;
            cc.pass.args.in_bin0 = 0;
            cc.pass.args.out = bufs.dst + (long long) b * bufs.dst_bs;
            cc.pass.args.out_bin0 = r0;
            cc.pass.args.gdiv = (int) (nr / cc.pass.C);
            cc.pass.args.ntiles = (int) (cc.pass.args.gdiv * L2);
            cc.pass.args.batch = 1;
            cc.pass.args.in_policy = keep;
            cc.pass.args.out_policy = strm;
            out.push_back (bb);
            out.push_back (cc);
        }
}

// Phases 1 + 2 of the distributed transform on this rank's rows, L2-chunked like the single-GPU three-pass plan (a chunk =
// nr of the rank's L1/world k1-rows: B: exchange layout -> ring, C: ring -> result), or whole-array through `s1`
// (chunk_elems = 0).  natural = false: transposed-out into `dst` ([S1][L1/world]).  natural = true: the second
// all-to-all is fused into pass C's stores -- bin k1 + L1 q belongs to rank q / (S1/world) = k3 >> (l3 - log2 world), whose
// natural-order block is mapped at peer_nat[h]; it lands at offset k1 + L1 (q mod S1/world) there.
inline bool build_dist_schedule (int n, const LargeFactors& f, int rank, int world, const float2* recv, float2* dst, float2* const* peer_nat, bool natural,
                                 float2* s1, float2* ring, long long ring_lane_elems, long long chunk_elems, int lanes, bool policies, std::vector<LargeLaunch>& out)
{
    TilePass p1, p2;
    if (! build_dist_phase (n, f, 1, rank, world, p1) || ! build_dist_phase (n, f, 2, rank, world, p2))
        return false;
    p1.which = 1;
    p2.which = 2;
    const long long L1 = 1LL << f.l1, L2 = 1LL << f.l2, S1 = 1LL << (f.l2 + f.l3), rows = L1 / world;
    int wl = 0;
    while ((1 << wl) < world)
        ++wl;
    if (natural)
    {
        if (world > 8 || f.l3 < wl)
            return false;
        p2.args.out_g_hi = L1;
        p2.args.out_g_lo = p2.C;
        p2.args.out_tstride = 1;
        p2.args.out_estride = L1 * L2;
        p2.args.peer_row_log = f.l3 - wl;
        p2.args.peer_chunk_off = (long long) rank * rows;
        for (int h = 0; h < world; ++h)
            p2.args.peer_out[h] = peer_nat[h];
    }
    out.clear();
    p1.args.in = recv;
    p2.args.out = dst;
    if (chunk_elems <= 0 || ring == nullptr)
    {
        p1.args.out = s1;
        p2.args.in = s1;
        out.push_back ({ p1, -1 });
        out.push_back ({ p2, -1 });
        return true;
    }
    const int keep = policies ? POLICY_KEEP : POLICY_NORMAL, strm = policies ? POLICY_STREAM : POLICY_NORMAL;
    long long nrc = ring_lane_elems / S1;
    nrc -= nrc % p2.C;
    if (nrc < p2.C)
        return false;
    int lane = nrc >= rows ? -1 : 0;
    for (long long r0 = 0; r0 < rows; r0 += nrc, lane = lane < 0 ? -1 : (lane + 1) % lanes)
    {
        const long long nr = rows - r0 < nrc ? rows - r0 : nrc;
        float2* slot = ring + (long long) (lane < 0 ? 0 : lane) * ring_lane_elems;
        LargeLaunch bb { p1, lane }, cc { p2, lane };
        bb.pass.args.in_bin0 = r0 * p1.args.in_g_hi; // row r0 inside every chunk of the exchange layout
        bb.pass.args.out = slot;
        bb.pass.args.out_bin0 = 0;
        bb.pass.args.ntiles = (int) (nr * bb.pass.args.gdiv);
        bb.pass.args.in_policy = strm;
        bb.pass.args.out_policy = keep;
        cc.pass.args.in = slot;
        cc.pass.args.in_bin0 = 0;
        cc.pass.args.out_bin0 = r0;
        cc.pass.args.gdiv = (int) (nr / cc.pass.C);
        cc.pass.args.ntiles = (int) (cc.pass.args.gdiv * L2);
        cc.pass.args.in_policy = keep;
        cc.pass.args.out_policy = strm;
        out.push_back (bb);
        out.push_back (cc);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Tensor-map view of one tile pass (tile_tma_kernel): dims / strides / box of the 5-D maps of its input and output
// side and the per-tile coordinate rule, from the pass's TileArgs.  Pure arithmetic (shared with the emulator tests).
// Returns false when the pass cannot be expressed (split input rows, peer stores, unordered layouts, unaligned strides).
// ---------------------------------------------------------------------------------------------
struct TileTmaSide
{
    const float2* base;
    unsigned long long dims[5], strides[4];
    unsigned box[5];
    TileTmaCoords coords;
};
inline bool build_tile_tma (const TilePass& p, TileTmaSide& in, TileTmaSide& out)
{
    const TileArgs& a = p.args;
    if (p.uio != 0 || a.peer_row_log >= 0 || a.in_split_log < 31)
        return false;
    const int C = p.C, L = 1 << p.logL;
    const int rows_per_box = L < 256 ? L : 256;
    const long long nhi = a.ntiles / a.gdiv;
    auto even = [] (long long v) { return (v & 1) == 0; };
    if (! even (a.in_bstride) || ! even (a.out_bstride) || ! even (a.in_bin0) || ! even (a.out_bin0) || ! even (a.in_g_hi) || ! even (a.out_g_hi)
        || (reinterpret_cast<uintptr_t> (a.in) & 15) != 0 || (reinterpret_cast<uintptr_t> (a.out) & 15) != 0)
        return false;
    auto strided_side = [&] (TileTmaSide& s, const float2* base, long long bin0, long long g_hi, long long estride, long long bstride)
    {
        s = {};
        s.base = base + bin0;
        s.dims[0] = (unsigned long long) a.gdiv * C * 2;
        s.dims[1] = (unsigned long long) L;
        s.dims[2] = (unsigned long long) nhi;
        s.dims[3] = (unsigned long long) a.batch;
        s.dims[4] = 1;
        s.strides[0] = (unsigned long long) estride * 8;
        s.strides[1] = (unsigned long long) (nhi > 1 ? g_hi : estride * L) * 8;
        s.strides[2] = (unsigned long long) (a.batch > 1 ? bstride : (nhi > 1 ? g_hi * nhi : estride * L)) * 8;
        s.strides[3] = s.strides[2] * (unsigned long long) (a.batch > 1 ? a.batch : 1);
        s.box[0] = (unsigned) (2 * C); s.box[1] = (unsigned) rows_per_box; s.box[2] = s.box[3] = s.box[4] = 1;
        s.coords.per_lo[0] = 2 * C;
        s.coords.per_hi[2] = 1;
        s.coords.batch_dim = 3;
        s.coords.iter_dim = 1;
        s.coords.iters = L / rows_per_box;
        s.coords.iter_step = rows_per_box;
    };
    if (! p.load_j_fast)
    {
        if (a.in_tstride != 1 || a.out_tstride != 1 || a.in_g_lo != C || a.out_g_lo != C || ! even (a.in_estride) || ! even (a.out_estride))
            return false;
        strided_side (in, a.in, a.in_bin0, a.in_g_hi, a.in_estride, a.in_bstride);
        strided_side (out, a.out, a.out_bin0, a.out_g_hi, a.out_estride, a.out_bstride);
        return true;
    }
    // contiguous-row pass: input [k1][k2][n3] (tile = C adjacent k1 at one k2), output transposed k1 + L1 (k2 + L2 k3)
    if (a.in_estride != 1 || a.out_tstride != 1 || a.out_g_lo != C || a.in_g_lo != (long long) C * a.in_tstride || ! even (a.in_tstride) || ! even (a.out_estride))
        return false;
    const int chunk = 2 * L < 256 ? 2 * L : 256;
    in = {};
    in.base = a.in + a.in_bin0;
    in.dims[0] = (unsigned long long) chunk;
    in.dims[1] = (unsigned long long) (2 * L / chunk);
    in.dims[2] = (unsigned long long) nhi;
    in.dims[3] = (unsigned long long) a.gdiv * C;
    in.dims[4] = (unsigned long long) a.batch;
    in.strides[0] = (unsigned long long) chunk * 4;
    in.strides[1] = (unsigned long long) (nhi > 1 ? a.in_g_hi : L) * 8;
    in.strides[2] = (unsigned long long) a.in_tstride * 8;
    in.strides[3] = (unsigned long long) (a.batch > 1 ? a.in_bstride : a.in_tstride * a.gdiv * C) * 8;
    in.box[0] = (unsigned) chunk; in.box[1] = (unsigned) (2 * L / chunk); in.box[2] = 1; in.box[3] = (unsigned) C; in.box[4] = 1;
    in.coords.per_hi[2] = 1;
    in.coords.per_lo[3] = C;
    in.coords.batch_dim = 4;
    in.coords.iter_dim = 0;
    in.coords.iters = 1;
    in.coords.iter_step = 0;
    out = {};
    out.base = a.out + a.out_bin0;
    out.dims[0] = (unsigned long long) a.gdiv * C * 2;
    out.dims[1] = (unsigned long long) nhi;
    out.dims[2] = (unsigned long long) L;
    out.dims[3] = (unsigned long long) a.batch;
    out.dims[4] = 1;
    out.strides[0] = (unsigned long long) (nhi > 1 ? a.out_g_hi : a.gdiv * C) * 8;
    out.strides[1] = (unsigned long long) a.out_estride * 8;
    out.strides[2] = (unsigned long long) (a.batch > 1 ? a.out_bstride : a.out_estride * L) * 8;
    out.strides[3] = out.strides[2] * (unsigned long long) (a.batch > 1 ? a.batch : 1);
    out.box[0] = (unsigned) (2 * C); out.box[1] = 1; out.box[2] = (unsigned) rows_per_box; out.box[3] = out.box[4] = 1;
    out.coords.per_lo[0] = 2 * C;
    out.coords.per_hi[1] = 1;
    out.coords.batch_dim = 3;
    out.coords.iter_dim = 2;
    out.coords.iters = L / rows_per_box;
    out.coords.iter_step = rows_per_box;
    return true;
}

inline int big_twiddle_lobits (int n) { return n < 28 ? (n + 1) / 2 : 14; }
} // namespace cfb
