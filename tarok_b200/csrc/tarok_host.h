// Internal interface between the host-side serialiser (tarok_host.cpp, g++) and the CUDA translation unit (tarok_abi.cu).
#pragma once
#include <cstdint>

struct tarok_pack_pool;
tarok_pack_pool* tarok_pack_pool_create(int threads);
void tarok_pack_pool_destroy(tarok_pack_pool* p);
int tarok_pack_pool_threads(const tarok_pack_pool* p);
// One job = rows [0, n) -> records; `chunk` rows (rounded up to the pool's block size: read it back with _chunk_rows) is the
// granularity _wait_chunk reports on.  _begin returns at once; _wait_chunk(c) packs along until chunk c is complete.
void tarok_pack_pool_begin(tarok_pack_pool* p, const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                           const uint8_t* king, uint64_t n, uint64_t chunk, uint64_t* records);
uint64_t tarok_pack_pool_chunk_rows(const tarok_pack_pool* p);
void tarok_pack_pool_wait_chunk(tarok_pack_pool* p, int c);
int64_t tarok_pack_pool_bad(const tarok_pack_pool* p);
