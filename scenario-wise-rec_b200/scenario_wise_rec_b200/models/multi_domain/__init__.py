"""Drop-in multi-domain models (reference: scenario_wise_rec/models/multi_domain/__init__.py)."""
from .sharebottom import SharedBottom
from .mmoe import MMOE
from .ple import PLE
from .star import Star
from .ppnet import PPNet
from .epnet import EPNet
from .hamur import HamurLarge, HamurSmall
from .m3oe import M3oE

__all__ = ["SharedBottom", "MMOE", "PLE", "Star", "PPNet", "EPNet", "HamurLarge", "HamurSmall", "M3oE"]
