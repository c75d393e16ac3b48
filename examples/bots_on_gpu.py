"""Four uniform-random Bot_igralec players, one million concurrent deals, entirely on the GPU.

Same call shape as the reference's driver (main.py:114-115: ``Tarok(igralci, num_games).paralel_start()``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import time

from tarok_b200 import Bot_igralec, Tarok

if __name__ == "__main__":
    igralci = [Bot_igralec() for _ in range(4)]
    for poskus in ("first call (includes CUDA start-up and the buffer allocation)",
                   "second call (the device environment is kept between calls)"):
        t = Tarok(igralci, 1_000_000, seed=2026)
        t0 = time.time()
        t.paralel_start()                   # prints t.rezultati like the reference
        print("Time need for 1000000 game: %.4f s -- %s" % (time.time() - t0, poskus))
    print("contract histogram (Klop, Tri, Dve, Ena, ...):", t.statistika[8:18].tolist())
