"""BASELINE config 4 in miniature: policy-net forward + GPU env step, replay targets, one training pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tarok_b200.samoigra import Samoigra

if __name__ == "__main__":
    s = Samoigra(16384, seed=1, random_card=0.1)
    for it in range(3):
        stats, ms = s.odigraj(first_game_id=it * 16384, meri=True)
        print("iteration", it, "env-steps", int(stats[19]), "contracts", stats[8:18].tolist(),
              {k: round(v, 1) for k, v in ms.items()}, "ms")
        print("  loss per net:", {k: round(v, 3) for k, v in s.nauci().items()})
    s.zapri()
