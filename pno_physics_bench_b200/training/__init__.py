from .losses import ELBOLoss, PNOLoss  # noqa: F401
from .trainer import PNOTrainer  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
