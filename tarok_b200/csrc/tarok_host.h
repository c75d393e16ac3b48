// Internal interface between the host-side serialiser (tarok_host.cpp, g++) and the CUDA translation unit (tarok_abi.cu).
#pragma once
#include <cstdint>

struct tarok_pack_pool;
tarok_pack_pool* tarok_pack_pool_create(int threads);
void tarok_pack_pool_destroy(tarok_pack_pool* p);
int tarok_pack_pool_threads(const tarok_pack_pool* p);
int64_t tarok_pack_pool_run(tarok_pack_pool* p, const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                            const uint8_t* king, uint64_t a, uint64_t b, uint64_t* records);
