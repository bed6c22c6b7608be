// Host-side description of a multi-pass (four-step) transform: factorisation and per-pass tile
// arguments.  Pure arithmetic, shared by the library (capi.cu) and the CPU emulator tests.
#pragma once
#include "large_kernels.cuh"

namespace cfb
{
constexpr int kTileC = 8;        // adjacent transforms per tile: 64 contiguous bytes on the strided side
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
    }
    if (f.l2 != 0)
    {   // pass B: for every row k1, columns of its [L2][L3] view
        TilePass& b = p[np++];
        b = {};
        b.logL = f.l2;
        b.load_j_fast = false;
        b.args.gdiv = (int) (L3 / kTileC);
        b.args.ntiles = (int) (L1 * L3 / kTileC);
        b.args.in_g_hi = b.args.out_g_hi = L2 * L3;
        b.args.in_g_lo = b.args.out_g_lo = kTileC;
        b.args.in_tstride = b.args.out_tstride = 1;
        b.args.in_estride = b.args.out_estride = L3;
        b.args.tw_mult = (unsigned) L1 * tw_scale;
        b.args.batch = 1;
    }
    {   // pass C: contiguous rows (k1, k2), written transposed to k1 + L1 (k2 + L2 k3)
        TilePass& c = p[np++];
        c = {};
        c.logL = f.l3;
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
    }
    return np;
}

inline int big_twiddle_lobits (int n) { return n < 28 ? (n + 1) / 2 : 14; }
} // namespace cfb
