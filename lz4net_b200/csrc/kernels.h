// kernels.h -- internal launcher interface between the C-ABI translation unit and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lz4b200 {

struct BatchArgs {
    const uint8_t* src; const int64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const int64_t* dst_off; const int32_t* dst_cap;
    int32_t* out_len; int32_t n_blocks;
};

struct DeviceInfo { int num_sms; int smem_per_sm; int smem_optin; };

// All launchers are asynchronous on `stream`; `counter` is a device uint32 the launcher zeroes itself (dynamic
// block hand-out).  They return the launch error (cudaSuccess on success) and add to *launches.
// lanes_per_block: 32 | 16 | 8 | 4 (+100 = output-staged), 1 | 2 = the lane-per-block decoder, 0 = chosen on the device from
// the batch's compression ratio.  `counter` points at FOUR device words the launcher owns (block counter, pick, two more counters).
int         decode_lanes_for_ratio(double ratio);
cudaError_t launch_decode(const BatchArgs& a, bool known_len, int lanes_per_block, uint32_t* counter,
                          const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
cudaError_t launch_encode_fast(const BatchArgs& a, uint32_t* counter, int warps_per_sm, const int* tune4 /* prefetch, lane_copy_max, probe_max, wide_min */, int variant,
                               const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
size_t      hc_scratch_bytes(int concurrency);
cudaError_t launch_encode_hc(const BatchArgs& a, void* scratch, int concurrency, uint32_t* counter,
                             const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
// blocks <= 64 KiB: one warp per block on a static index, followed by the thread-per-block kernel over the blocks it handed back
// (variant 1: the block staged in shared memory, <= 3 warps per SM; 2: nothing in shared memory, <= 32 warps per SM)
size_t      hcw_scratch_bytes(int32_t n_blocks, int variant, int warps_per_sm, const DeviceInfo& dev);
cudaError_t launch_encode_hcw(const BatchArgs& a, void* scratch, int variant, int warps_per_sm, uint32_t* counter,
                              const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
// kernel chosen per batch: small batches -> the shared-memory warp kernel; large ones -> a sample of the batch decides on
// the device between the thread kernel and the warp kernel (`counter` = four words)
size_t      hc_auto_scratch_bytes(int32_t n_blocks, int concurrency, const DeviceInfo& dev);
cudaError_t launch_encode_hc_auto(const BatchArgs& a, void* scratch, int concurrency, uint32_t* counter,
                                  const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
cudaError_t launch_compact(const uint8_t* slots, const int64_t* slot_off, const int32_t* len, uint8_t* packed,
                           int64_t* out_off, int32_t n_blocks, void* scan_tmp, size_t scan_tmp_bytes,
                           const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);
size_t      compact_tmp_bytes(int32_t n_blocks);
cudaError_t launch_synth(uint8_t* dst, int64_t n_blocks, int32_t block_size, int cls, uint64_t seed, int64_t first_block,
                         const DeviceInfo& dev, cudaStream_t stream, int64_t* launches);

}  // namespace lz4b200
