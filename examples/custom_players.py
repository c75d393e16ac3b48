"""Any ``Igralec`` subclass written for the reference runs on the CUDA engine through the callback protocol.

Put ``tarok_b200/compat`` first on sys.path and the reference's own import lines keep working."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tarok_b200", "compat"))

from Igralec import Bot_igralec          # noqa: E402  (tarok_b200.igralec)
from Tarok import Tarok                  # noqa: E402  (tarok_b200.igra)
from Tip_igre import Tip_igre            # noqa: E402


class Previdni(Bot_igralec):
    """Plays the lowest legal card, never bids above Tri."""
    device_policy = None                 # host callbacks, not the device-side bot

    def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
        from tarok_b200.igralec import Igralec
        return Igralec.licitiram(self, random.choice([Tip_igre.Naprej, Tip_igre.Tri]), min_igra, id_igre, obvezno, prednost)

    def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
        from tarok_b200.igralec import Igralec
        return Igralec.igraj_karto(self, min(mozne), id_igre)


if __name__ == "__main__":
    igralci = [Previdni(), Bot_igralec(), Previdni(), Bot_igralec()]
    for b in igralci:
        b.device_policy = None
    t = Tarok(igralci, 64)
    t.paralel_start()
