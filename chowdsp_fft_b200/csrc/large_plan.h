// Host-side description of a multi-pass (four-step) transform: factorisation and per-pass tile
// arguments.  Pure arithmetic, shared by the library (capi.cu) and the CPU emulator tests.
#pragma once
#include "large_kernels.cuh"

namespace cfb
{
// adjacent transforms per tile = contiguous complex elements on the strided side of a pass: 16 (one full 128-byte
// line per element row) for the strided passes whenever the CTA stays at <= 512 threads, 8 (64 bytes) for
// 1024-point passes and for the contiguous-row pass
constexpr int kTileCMax = 16;
inline int& tile_c_override() { static int v = 0; return v; }       // tuning hooks (0 = policy below)
inline int& tile_c_jfast_override() { static int v = 0; return v; } // ... for the contiguous-row (last) pass only
// tuning hook "tile_pipe": 1 = the persistent TMA-staged tile kernel (tile_pipe_kernel) where its buffers fit and the rows are
// 16-byte aligned, 0 = tile_fft_kernel everywhere
inline int& tile_pipe_mode() { static int v = 0; return v; }
// tuning hook "tile_pf": L2 prefetch distance of the tile passes in tiles (0 = off)
inline int& tile_pf_ahead() { static int v = 0; return v; }
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
    }
    if (f.l2 != 0)
    {   // pass B: for every row k1, columns of its [L2][L3] view
        TilePass& b = p[np++];
        b = {};
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
    }
    {   // pass C: contiguous rows (k1, k2), written transposed to k1 + L1 (k2 + L2 k3)
        TilePass& c = p[np++];
        c = {};
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

inline int big_twiddle_lobits (int n) { return n < 28 ? (n + 1) / 2 : 14; }
} // namespace cfb
