// encode.cu -- kernels + launchers for the fast and the HC encoders.
#include "kernels.h"
#include "lz4_encode.cuh"
#include "lz4hc_encode.cuh"

namespace lz4b200 {

// ---- fast encoder: one warp per block, persistent, dynamic block hand-out --------------------------------------------
// All encoder warps of an SM live in ONE CTA.  Up to 14 of them keep their 16 KiB position table in shared memory (14 x
// 16 KiB = 224 KiB of dynamic shared memory; separate CTAs would each pay 1 KiB of system-reserved shared memory, which
// costs the 14th warp).  The kernel is a per-warp latency chain (throughput is linear in resident warps), so the warps
// beyond those 14 -- as many as the register file holds at the launch bound NW -- run the same parse with their table
// in a global-memory arena that stays L2-resident (16 KiB per warp): a table round trip costs them an L2 round trip,
// but every one of them is one more block in flight.  The warps never synchronise with each other.
constexpr int ENC_SMEM_WARPS = 14;

template <int DUP, int NW, int GDUP>
__global__ void __launch_bounds__(32 * NW, 1)
lz4_encode_fast_kernel(BatchArgs a, uint32_t* counter, EncTune tune, int smem_warps, uint8_t* arena)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool gt = NW > ENC_SMEM_WARPS && warp >= smem_warps;
    void* table = gt ? (void*)(arena + ((size_t)blockIdx.x * NW + warp) * sizeof(EncShared))
                     : (void*)((EncShared*)smem + warp);
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(counter, 1u);
        b = simt::shfl(0xFFFFFFFFu, b, 0);
        if (b >= (uint32_t)a.n_blocks) break;
        int r;
        if (NW > ENC_SMEM_WARPS && gt)
            r = encode_block<GDUP, 0, true>(table, a.src + a.src_off[b], a.src_len[b], a.dst + a.dst_off[b], a.dst_cap[b], lane, tune);
        else
            r = encode_block<DUP, 0, false>(table, a.src + a.src_off[b], a.src_len[b], a.dst + a.dst_off[b], a.dst_cap[b], lane, tune);
        if (lane == 0) a.out_len[b] = r;
    }
}

template <int DUP, int NW, int GDUP>
static cudaError_t launch_fast_t(const BatchArgs& a, uint32_t* counter, int dyn, long long grid, int warps, int smem_warps,
                                 uint8_t* arena, const EncTune& tune, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(lz4_encode_fast_kernel<DUP, NW, GDUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    lz4_encode_fast_kernel<DUP, NW, GDUP><<<(unsigned)grid, 32 * warps, dyn, stream>>>(a, counter, tune, smem_warps, arena);
    return cudaGetLastError();
}

size_t encode_arena_bytes(int warps_per_sm, const DeviceInfo& dev)
{
    return warps_per_sm > ENC_SMEM_WARPS ? (size_t)28 * dev.num_sms * sizeof(EncShared) : 0;     // one slot per warp of the widest kernel
}

cudaError_t launch_encode_fast(const BatchArgs& a, uint32_t* counter, int warps_per_sm, int smem_warps_max, const int* tune4, int variant,
                               void* arena, const DeviceInfo& dev, cudaStream_t stream, int64_t* launches)
{
    if (a.n_blocks <= 0) return cudaSuccess;
    int smem_warps = dev.smem_optin / (int)sizeof(EncShared);
    if (smem_warps > ENC_SMEM_WARPS) smem_warps = ENC_SMEM_WARPS;
    if (smem_warps < 1) smem_warps = 1;
    if (warps_per_sm < 1) warps_per_sm = smem_warps;
    const int smem_full = smem_warps;
    if (warps_per_sm > ENC_SMEM_WARPS && smem_warps_max > 0 && smem_warps_max < smem_warps) smem_warps = smem_warps_max;   // (tests: global-table warps in small batches)
    // the launch bounds the kernel is built for: 14 (all tables in shared memory), 18, 20, 24, 28
    const int nw = warps_per_sm <= 14 ? 14 : (warps_per_sm <= 18 ? 18 : (warps_per_sm <= 20 ? 20 : (warps_per_sm <= 24 ? 24 : 28)));
    if (warps_per_sm > nw) warps_per_sm = nw;
    if (nw == 14 && warps_per_sm > smem_warps) warps_per_sm = smem_warps;
    if (nw > 14 && (smem_full < ENC_SMEM_WARPS || !arena)) return cudaErrorInvalidValue;
    // small batches: spread the blocks over the SMs first
    long long grid = dev.num_sms;
    int warps = warps_per_sm;
    if (a.n_blocks < (long long)dev.num_sms * warps) {
        warps = (int)((a.n_blocks + dev.num_sms - 1) / dev.num_sms);
        grid = (a.n_blocks + warps - 1) / warps;
    }
    const int sw = warps < smem_warps ? warps : smem_warps;
    const int dyn = sw * (int)sizeof(EncShared);
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    EncTune tune; tune.pf_dist = tune4[0]; tune.lane_copy_max = tune4[1]; tune.probe_max = tune4[2]; tune.wide_min = tune4[3];
    uint8_t* ar = (uint8_t*)arena;
    // variant % 10: how same-hash iterations of one round are found by the shared-memory warps (lz4_encode.cuh): 1 = always
    // exact (one vote per hash bit), 2 = through the table, pairs resolved in place; variant / 10 (0 -> 2): the same for
    // the warps whose table is in global memory
#define LZ4B200_ENC_CASE(NW_) \
    case NW_: return variant / 10 == 1 ? launch_fast_t<2, NW_, 1>(a, counter, dyn, grid, warps, sw, ar, tune, stream) \
                                       : launch_fast_t<2, NW_, 2>(a, counter, dyn, grid, warps, sw, ar, tune, stream);
    switch (nw) {
    LZ4B200_ENC_CASE(18) LZ4B200_ENC_CASE(20) LZ4B200_ENC_CASE(24) LZ4B200_ENC_CASE(28)
    default: break;
    }
#undef LZ4B200_ENC_CASE
    return variant % 10 == 1 ? launch_fast_t<1, 14, 1>(a, counter, dyn, grid, warps, sw, ar, tune, stream)
                             : launch_fast_t<2, 14, 2>(a, counter, dyn, grid, warps, sw, ar, tune, stream);
}

// ---- HC encoder: one THREAD per block, state arena in global memory -------------------------------------------------
constexpr int HC_THREADS = 32;

__global__ void __launch_bounds__(HC_THREADS)
lz4_encode_hc_kernel(BatchArgs a, uint8_t* arena, uint32_t* counter)
{
    const size_t slot = (size_t)blockIdx.x * HC_THREADS + threadIdx.x;
    void* state = arena + slot * HC_STATE_BYTES;
    for (;;) {
        const uint32_t b = atomicAdd(counter, 1u);
        if (b >= (uint32_t)a.n_blocks) break;
        a.out_len[b] = hc_encode_block(state, a.src + a.src_off[b], a.src_len[b], a.dst + a.dst_off[b], a.dst_cap[b]);
    }
}

size_t hc_scratch_bytes(int concurrency)
{
    size_t slots = ((size_t)concurrency + HC_THREADS - 1) / HC_THREADS * HC_THREADS;
    return slots * HC_STATE_BYTES;
}

cudaError_t launch_encode_hc(const BatchArgs& a, void* scratch, int concurrency, uint32_t* counter,
                             const DeviceInfo&, cudaStream_t stream, int64_t* launches)
{
    if (a.n_blocks <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    long long ctas = ((long long)concurrency + HC_THREADS - 1) / HC_THREADS;
    long long want = ((long long)a.n_blocks + HC_THREADS - 1) / HC_THREADS;
    if (ctas > want) ctas = want;
    if (ctas < 1) ctas = 1;
    lz4_encode_hc_kernel<<<(unsigned)ctas, HC_THREADS, 0, stream>>>(a, (uint8_t*)scratch, counter);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace lz4b200
