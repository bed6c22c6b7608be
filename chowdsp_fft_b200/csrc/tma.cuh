// Tensor-map TMA helpers shared by the kernels that move whole tiles with one instruction (cp.async.bulk.tensor, SASS
// UTMALDG / UTMASTG): a 5-D descriptor (lower ranks are padded with size-1 dimensions), load global -> shared with completion on
// an mbarrier, store shared -> global as a bulk group.  On the device the descriptor is the driver's CUtensorMap (created on the
// host by cuTensorMapEncodeTiled, see tma_host.h); in the CPU emulator (tests/emu) it is a plain description and the copies are
// memcpy loops, so the kernels' tile addressing is checked without a GPU.
#pragma once
#include "pipe_kernels.cuh" // smem_addr, mbarrier helpers

namespace cfb
{
#ifdef CHOWDSP_EMU
struct TensorMap5
{
    const char* base;
    unsigned long long dim[5];    // floats per dimension, innermost first
    unsigned long long stride[4]; // bytes between consecutive indices of dimensions 1..4
    unsigned box[5];
};
#define CFB_TMAP5_PARAM const TensorMap5
#else
struct alignas (64) TensorMap5
{
    unsigned long long opaque[16]; // CUtensorMap
};
#define CFB_TMAP5_PARAM const __grid_constant__ TensorMap5
#endif

#ifdef CHOWDSP_EMU
FFT_HD const char* tmap5_row (const TensorMap5* map, int c0, long long i1, long long i2, long long i3, long long i4)
{
    return map->base + (unsigned long long) i4 * map->stride[3] + (unsigned long long) i3 * map->stride[2] + (unsigned long long) i2 * map->stride[1]
           + (unsigned long long) i1 * map->stride[0] + (unsigned long long) c0 * 4;
}
#endif

// box at coordinates (c0 .. c4) -> dst (dense, dimension 0 fastest); the bytes are counted on `bar`
FFT_HD void tma_load_5d (void* dst, const TensorMap5* map, int c0, int c1, int c2, int c3, int c4, unsigned long long* bar)
{
#ifdef CHOWDSP_EMU
    char* d = static_cast<char*> (dst);
    unsigned long long bytes = 0;
    for (unsigned i4 = 0; i4 < map->box[4]; ++i4)
        for (unsigned i3 = 0; i3 < map->box[3]; ++i3)
            for (unsigned i2 = 0; i2 < map->box[2]; ++i2)
                for (unsigned i1 = 0; i1 < map->box[1]; ++i1)
                {
                    std::memcpy (d + bytes, tmap5_row (map, c0, c1 + i1, c2 + i2, c3 + i3, c4 + i4), (size_t) map->box[0] * 4);
                    bytes += (unsigned long long) map->box[0] * 4;
                }
    __atomic_fetch_add (bar, bytes, __ATOMIC_RELEASE);
#else
    asm volatile ("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                  ::"r"(smem_addr (dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_addr (bar)) : "memory");
#endif
}

// ask L2 for the box at coordinates (c0 .. c4) ahead of time: no shared-memory destination, no completion (SASS UTMAPF)
FFT_HD void tma_prefetch_5d (const TensorMap5* map, int c0, int c1, int c2, int c3, int c4)
{
#ifdef CHOWDSP_EMU
    (void) map; (void) c0; (void) c1; (void) c2; (void) c3; (void) c4;
#else
    asm volatile ("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
#endif
}

// src (dense box image) -> box at coordinates (c0 .. c4); joins the calling thread's current bulk group
FFT_HD void tma_store_5d (const void* src, const TensorMap5* map, int c0, int c1, int c2, int c3, int c4)
{
#ifdef CHOWDSP_EMU
    const char* sp = static_cast<const char*> (src);
    unsigned long long bytes = 0;
    for (unsigned i4 = 0; i4 < map->box[4]; ++i4)
        for (unsigned i3 = 0; i3 < map->box[3]; ++i3)
            for (unsigned i2 = 0; i2 < map->box[2]; ++i2)
                for (unsigned i1 = 0; i1 < map->box[1]; ++i1)
                {
                    std::memcpy (const_cast<char*> (tmap5_row (map, c0, c1 + i1, c2 + i2, c3 + i3, c4 + i4)), sp + bytes, (size_t) map->box[0] * 4);
                    bytes += (unsigned long long) map->box[0] * 4;
                }
#else
    asm volatile ("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                  ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_addr (src)) : "memory");
#endif
}
FFT_HD void tma_store_commit()
{
#ifndef CHOWDSP_EMU
    asm volatile ("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
// the calling thread's bulk groups have finished READING their shared-memory sources (the buffers may be overwritten)
FFT_HD void tma_store_wait_read()
{
#ifndef CHOWDSP_EMU
    asm volatile ("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
// ... have completed (global writes performed)
FFT_HD void tma_store_wait_all()
{
#ifndef CHOWDSP_EMU
    asm volatile ("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
} // namespace cfb
