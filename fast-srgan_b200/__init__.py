"""fast-srgan_b200: B200-native (sm_100a) engine for the Fast-SRGAN hot path.

Mirrors the reference's Python surface (model.py / trainer.py / inference.py) on top of the
C-ABI library libfsr_b200.so (include/fsr_b200.h).  Import as `fast_srgan_b200` (the directory name
has a hyphen; the sibling shim package sets __path__ here)."""
__version__ = "0.1.0"
