from tarok_b200.igra import Klop  # noqa: F401
