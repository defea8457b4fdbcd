// encode.cu -- kernels + launchers for the fast and the HC encoders.
#include "kernels.h"
#include "lz4_encode.cuh"
#include "lz4_encode_lpb.cuh"
#include "lz4hc_encode.cuh"

namespace lz4b200 {

// ---- fast encoder: persistent, dynamic block hand-out ----------------------------------------------------------------
// All encoder warps of an SM live in ONE CTA.  14 of them are warp-per-block encoders (lz4_encode.cuh) whose 16 KiB
// position tables fill the shared memory (14 x 16 KiB = 224 KiB of dynamic shared memory; separate CTAs would each pay
// 1 KiB of system-reserved shared memory, which costs the 14th warp).  That kernel is a per-warp latency chain -- issue
// slots half idle, and no room for a 15th table -- so an optional 15th warp runs the lane-per-block encoder
// (lz4_encode_lpb.cuh): 32 more blocks in flight per SM, one per lane, tables in an L2-resident global arena.
// All warps take blocks from the same counter and never synchronise with each other.
// (Measured and dropped in round 2: more warp-per-block encoders with their tables in global memory -- 18 / 20 / 24 / 28
// warps per SM ran at 0.83 / 0.74 / 0.5 / 0.4 of the 14-warp kernel's throughput, profiles/encoder_r02_history.md.)
constexpr int ENC_SMEM_WARPS = 14;

template <int DUP, bool LPB>
__global__ void __launch_bounds__(32 * (ENC_SMEM_WARPS + (LPB ? 1 : 0)), 1)
lz4_encode_fast_kernel(BatchArgs a, unsigned long long* queue, EncTune tune, int warp_warps, uint8_t* arena, uint32_t reserve)
{
    // queue: (blocks taken from the front) << 32 | (blocks taken from the back) -- warps take from the front, the lanes
    // of the lane-per-block warp from the back (lz4_encode_lpb.cuh)
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (LPB && warp >= warp_warps) {
        const EncLpbBatch b{a.src, a.src_off, a.src_len, a.dst, a.dst_off, a.dst_cap, a.out_len, a.n_blocks};
        lpb_encode_warp(arena + (size_t)blockIdx.x * ENC_LPB_LANES * ENC_LPB_TABLE, b, queue, reserve, lane);
        return;
    }
    EncShared* sh = (EncShared*)smem + warp;
    for (;;) {
        uint32_t b = 0xFFFFFFFFu;
        if (lane == 0) {
            const unsigned long long old = atomicAdd(queue, 1ull << 32);
            const uint32_t f = (uint32_t)(old >> 32), taken_back = (uint32_t)old;
            if (f < (uint32_t)a.n_blocks && f + taken_back < (uint32_t)a.n_blocks) b = f;
        }
        b = simt::shfl(0xFFFFFFFFu, b, 0);
        if (b == 0xFFFFFFFFu) break;
        const int r = encode_block<DUP, 0>(sh, a.src + a.src_off[b], a.src_len[b], a.dst + a.dst_off[b], a.dst_cap[b], lane, tune);
        if (lane == 0) a.out_len[b] = r;
    }
}

template <int DUP, bool LPB>
static cudaError_t launch_fast_t(const BatchArgs& a, uint32_t* counter, int dyn, long long grid, int warps, uint8_t* arena,
                                 const EncTune& tune, cudaStream_t stream, bool forced = false)
{
    cudaError_t e = cudaFuncSetAttribute(lz4_encode_fast_kernel<DUP, LPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    // the lanes stop taking blocks when fewer than `reserve` are left: enough for every warp to stay busy for about one
    // lane-block time (a lane holds a block ~8 warp-block times)
    const uint32_t reserve = forced ? 0u : (uint32_t)(grid * warps * 8);       // (forced: tests with small batches)
    lz4_encode_fast_kernel<DUP, LPB><<<(unsigned)grid, 32 * (warps + (LPB ? 1 : 0)), dyn, stream>>>(a, (unsigned long long*)counter, tune, warps, arena, reserve);
    return cudaGetLastError();
}

size_t encode_arena_bytes(const DeviceInfo& dev) { return (size_t)dev.num_sms * ENC_LPB_LANES * ENC_LPB_TABLE; }     // one table per block-owning lane

cudaError_t launch_encode_fast(const BatchArgs& a, uint32_t* counter, int warps_per_sm, int lane_warp, const int* tune4, int variant,
                               void* arena, const DeviceInfo& dev, cudaStream_t stream, int64_t* launches)
{
    if (a.n_blocks <= 0) return cudaSuccess;
    int max_warps = dev.smem_optin / (int)sizeof(EncShared);
    if (max_warps > ENC_SMEM_WARPS) max_warps = ENC_SMEM_WARPS;
    if (max_warps < 1) max_warps = 1;
    if (warps_per_sm < 1 || warps_per_sm > max_warps) warps_per_sm = max_warps;
    // small batches: spread the blocks over the SMs first, and leave the lane-per-block warp out (its lanes take a block
    // each and keep it ~30x longer than a warp does: it pays only when every SM has many blocks to go through)
    long long grid = dev.num_sms;
    int warps = warps_per_sm;
    bool lpb = lane_warp != 0 && arena != nullptr;
    if (lane_warp == 1 && a.n_blocks < (long long)dev.num_sms * warps * 32) lpb = false;    // (lane_warp == 2 forces it: tests)
    if (a.n_blocks < (long long)dev.num_sms * warps) {
        warps = (int)((a.n_blocks + dev.num_sms - 1) / dev.num_sms);
        grid = (a.n_blocks + warps - 1) / warps;
    }
    const int dyn = warps * (int)sizeof(EncShared);
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);      // (the context hands out counters in 8-byte pairs)
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    EncTune tune; tune.pf_dist = tune4[0]; tune.lane_copy_max = tune4[1]; tune.probe_max = tune4[2]; tune.wide_min = tune4[3];
    uint8_t* ar = (uint8_t*)arena;
    // variant: how same-hash iterations of one round are found by the warp-per-block encoders (lz4_encode.cuh):
    // 1 = always exact (one vote per hash bit), 2 = through the table, pairs resolved in place
    if (lpb) return variant == 1 ? launch_fast_t<1, true>(a, counter, dyn, grid, warps, ar, tune, stream, lane_warp == 2)
                                 : launch_fast_t<2, true>(a, counter, dyn, grid, warps, ar, tune, stream, lane_warp == 2);
    return variant == 1 ? launch_fast_t<1, false>(a, counter, dyn, grid, warps, ar, tune, stream)
                        : launch_fast_t<2, false>(a, counter, dyn, grid, warps, ar, tune, stream);
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
