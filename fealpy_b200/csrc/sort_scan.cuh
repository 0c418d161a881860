// Device-wide exclusive scan and stable LSD radix sort of (u64 key, u32 payload) pairs.
// Hand-written for sm_100a (no CUB/Thrust): the deterministic COO->CSR build (north-star
// item 2) and the mesh-topology build (unique edges/faces) sit on these two primitives.
#pragma once
#include "common.cuh"

namespace fb2 {

// exclusive scan: out[i] = sum_{j<i} in[j], out has n+1 entries when `with_total`
// (out[n] = total).  In: u8 / u32 / i32; out: i64.  ws >= scan_workspace_bytes(n).
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_u8(const uint8_t* in, int64_t* out, int64_t n, bool with_total, void* ws, cudaStream_t s);
int exclusive_scan_u32(const uint32_t* in, int64_t* out, int64_t n, bool with_total, void* ws, cudaStream_t s);
int exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, bool with_total, void* ws, cudaStream_t s);

// stable LSD radix sort on key bits [0, nbits).  keys/vals are overwritten (ping-pong with
// the workspace); on return *keys_out / *vals_out point at whichever buffer holds the result.
// vals == nullptr on entry means "payload = original position" (the first pass synthesises it).
size_t sort_workspace_bytes(int64_t n);
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int nbits, void* ws, cudaStream_t s,
                     uint64_t** keys_out, uint32_t** vals_out);

}  // namespace fb2
