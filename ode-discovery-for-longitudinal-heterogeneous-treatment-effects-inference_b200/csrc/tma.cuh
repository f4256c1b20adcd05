// tma.cuh -- minimal sm_100a TMA / mbarrier wrappers (inline PTX) and host-side tensor-map encoding.
//
// The (N,T) float64 arrays of the reference I/O contract are row-major with a 480-byte pitch, so a
// thread-per-patient kernel cannot touch them directly without wasting 3/4 of every sector.  They are
// moved as 2-D boxes {TC time steps, P patients} by the TMA unit (cp.async.bulk.tensor, SASS
// UTMALDG/UTMASTG) into/out of swizzled shared-memory tiles, completion tracked by mbarriers
// (loads) and bulk async-groups (stores).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace b200i {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// same with an L2 eviction-priority hint (createpolicy value)
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *map, int c0, int c1, const void *smem_src, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                                 uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 3-D variants (column, row in group, group) for the line-aligned row-class mapping of sim_factual_ws
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                                 uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)),
          "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap *map, int c0, int c1, int c2, const void *smem_src,
                                                  uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
                 : "memory");
}
// one box of a tiled tensor map -> L2 (no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
// contiguous global range -> L2 (no shared-memory destination); addr and bytes multiples of 16
__device__ __forceinline__ void l2_prefetch_bulk(const void *gptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// contiguous global range -> shared memory (non-tensor bulk copy, SASS UBLKCP); both addresses and the size are multiples
// of 16 bytes; completion = complete_tx on the mbarrier
__device__ __forceinline__ void bulk_load_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// contiguous shared-memory range -> global (non-tensor bulk copy, SASS UBLKCP); both addresses and the size are
// multiples of 16 bytes; completion through the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_s2g_hint(void *gdst, const void *smem_src, uint32_t bytes, uint64_t policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
                 : "memory");
}
// all but the `N` most recently committed bulk stores have finished reading their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// byte offset of (row p, 16-byte unit q) inside a tile whose rows are ROW_BYTES (32/64/128) long and
// which TMA wrote/reads with the swizzle mode of the same span (address bits [4,4+b) ^= bits [7,7+b)).
// Tile bases are 1024-byte aligned, so the row index supplies bits 7.. directly.
template <int ROW_BYTES>
__device__ __forceinline__ uint32_t swz_off(uint32_t p, uint32_t q)
{
    constexpr uint32_t UNITS = ROW_BYTES / 16;
    const uint32_t x = ((p * ROW_BYTES) >> 7) & (UNITS - 1);
    return p * ROW_BYTES + ((q ^ x) << 4);
}

// ---- host side -------------------------------------------------------------------------------
// Encodes a 2-D tiled map over a (rows, cols) float64 row-major array: box = {box_cols, box_rows}.
int encode_tmap_2d_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols,
                       uint32_t box_rows, uint32_t box_cols, bool promote_256);
// 2-D map with an explicit row pitch (bytes): box = {box_cols, box_rows}, 128-byte rows, SWIZZLE_128B
int encode_tmap_2d_pitched_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                               uint32_t box_rows, uint32_t box_cols);
// The same array seen as {cols, group_rows, rows / group_rows}: box = {box_cols, 1, box_groups}, i.e. every
// group_rows-th row.  rows must be a multiple of group_rows.
int encode_tmap_3d_rowgroups_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t group_rows,
                                 uint32_t box_groups, uint32_t box_cols);

}  // namespace b200i
