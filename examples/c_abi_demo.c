/* Plain-C consumer of the C ABI (include/tarok_b200.h): one million Bot_igralec deals, no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo -Ltarok_b200 -ltarok_b200 -Wl,-rpath,$PWD/tarok_b200
 *   /tmp/c_abi_demo [n_games]
 */
#include <stdio.h>
#include <stdlib.h>

#include "tarok_b200.h"

int main(int argc, char** argv) {
    uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000000ull;
    tarok_t* h = NULL;
    if (tarok_create(0, n, 0x5EED7A20C0001ull, 0, &h) != 0) {
        fprintf(stderr, "tarok_create: %s\n", tarok_last_error(NULL));
        return 1;
    }
    /* deal, Bot bidding, exchange, 48 random plays, scoring -- all on the default stream */
    if (tarok_rollout_stepwise(h, TAROK_MODE_AUCTION_BOT, 0, NULL) != 0) {
        fprintf(stderr, "rollout: %s\n", tarok_last_error(h));
        return 1;
    }
    int64_t st[TAROK_STATS_LEN];
    if (tarok_read_stats(h, st, NULL) != 0) {
        fprintf(stderr, "read_stats: %s\n", tarok_last_error(h));
        return 1;
    }
    printf("deals %lld env-steps %lld errors %lld\n", (long long)st[18], (long long)st[19], (long long)st[20]);
    printf("rezultati by player: %lld %lld %lld %lld\n", (long long)st[4], (long long)st[5], (long long)st[6], (long long)st[7]);
    printf("contracts Klop/Tri/Dve/Ena: %lld %lld %lld %lld\n", (long long)st[8], (long long)st[9], (long long)st[10], (long long)st[11]);
    printf("kernel launches: %llu\n", (unsigned long long)tarok_launch_count(h));
    return tarok_destroy(h) == 0 ? 0 : 1;
}
