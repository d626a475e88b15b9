"""camradepth_b200 -- B200-native hot path of TUMFTM/CamRaDepth (drop-in nn.Module, losses, optimizer)."""
from .args import args, set_model                                        # noqa: F401
from .model import CamRaDepth, load_checkpoint_with_shape_match           # noqa: F401
from .losses import MaskedSmoothL1Loss, MaskedFocalLoss, MaskedMSELoss    # noqa: F401
from .optim import diffGradNorm                                           # noqa: F401
from .metrics import depth_metrics, mean_iou                              # noqa: F401
from .train import TrainStep, evaluate, save_checkpoint                             # noqa: F401

__all__ = ["args", "set_model", "CamRaDepth", "load_checkpoint_with_shape_match", "MaskedSmoothL1Loss",
           "MaskedFocalLoss", "MaskedMSELoss", "diffGradNorm"]
