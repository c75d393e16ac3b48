"""tarok_b200 -- B200-native batched Tarok environment behind the anzeA/Tarok class API.

    from tarok_b200 import Tarok, Igra, Igralec, Bot_igralec, Tip_igre, Karta, Barva, Roka   # reference names
    from tarok_b200 import TarokEnv                                                          # batched device env

The rules run in hand-written sm_100a CUDA kernels (tarok_b200/csrc) behind a C ABI
(include/tarok_b200.h, libtarok_b200.so).  There is no CPU fallback.
"""
from .karte import Barva, Karta, Roka, Tip_igre  # noqa: F401
from .igralec import Bot_igralec, Igralec  # noqa: F401


def __getattr__(name):      # torch / the CUDA library are only loaded when the engine is used
    if name in ("Igra", "Tarok", "Partije", "Navadna_igra", "Klop", "Berac"):
        from . import igra
        return getattr(igra, name)
    if name == "TarokEnv":
        from .env import TarokEnv
        return TarokEnv
    raise AttributeError(name)
