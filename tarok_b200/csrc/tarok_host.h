// Internal interface between the host-side serialiser (tarok_host.cpp, g++) and the CUDA translation unit (tarok_abi.cu).
#pragma once
#include <cstdint>

struct tarok_pack_pool;
tarok_pack_pool* tarok_pack_pool_create(int threads);
void tarok_pack_pool_destroy(tarok_pack_pool* p);
int tarok_pack_pool_threads(const tarok_pack_pool* p);
// Upload chunks of the host pipelines (tapered: small first and last chunks), boundaries on multiples of `quantum` rows.
int tarok_chunk_bounds(uint64_t n, int want, uint64_t quantum, uint64_t* bounds /* [want + 1] */);
uint64_t tarok_pack_block_rows(void);
// One job = rows [0, n) -> records; bounds[0..nchunks] = the upload chunks _wait_chunk reports on (multiples of the pack block).
// _begin returns at once; _wait_chunk(c) packs along until chunk c is complete.
void tarok_pack_pool_begin(tarok_pack_pool* p, const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                           const uint8_t* king, uint64_t n, const uint64_t* bounds, int nchunks, void* records);
void tarok_pack_pool_wait_chunk(tarok_pack_pool* p, int c);
int64_t tarok_pack_pool_bad(const tarok_pack_pool* p);
