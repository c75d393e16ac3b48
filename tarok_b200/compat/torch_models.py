"""Stand-in for the reference's ``torch_models.py``, which is missing upstream: ``Igralec.py:24`` does
``from torch_models import *`` and expects ``torch``, ``TensorDataset`` and the six network classes
(``Igralec.py:239-240,247,592``).  The classes are the PyTorch restatement of the Keras builders in ``train.py``
(``tarok_b200/mreze.py``; random init -- no weights exist upstream).  With this directory on ``sys.path`` the reference's own
``Nevronski_igralec`` can build its models and run its ``predict_*`` methods; Lightning training
(``load_from_checkpoint`` / ``Trainer.fit``) stays with whoever owns a real ``torch_models``."""
import torch  # noqa: F401
from torch import nn  # noqa: F401
from torch.utils.data import DataLoader, TensorDataset  # noqa: F401

from tarok_b200.mreze import (Net_Berac, Net_Klop, Net_Navadna_igra, Net_Solo,  # noqa: F401
                              Net_vrednotenje_roke, Net_zalaganje)
