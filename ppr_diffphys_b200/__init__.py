"""B200-native replacement for the ppr-diffphys motion-imitation rollout hot path.

Public surface (mirrors /root/reference/diffphys/dp_model.py:1014-1400):
  ForwardKinematics, ForwardWarp, convert_ppr_warp  -- torch.autograd Functions / helper
  Se3Loss, FrameCompose, RefsFromFrames             -- fused se3 loss / batch-input producer kernels
  SimEnv                                            -- the ``env``/``self`` object those Functions read
  load_robot / compile_robot / RobotModel           -- static model arrays
Importing this package never touches ``oracle/``; every compute entry point raises if the CUDA
library (``libppr_b200.so``) is missing -- there is no CPU fallback.
"""
from .model import RobotModel, compile_robot, load_robot, ROBOT_PRESETS  # noqa: F401

__all__ = ["RobotModel", "compile_robot", "load_robot", "ROBOT_PRESETS"]
from .ops import (ForwardKinematics, ForwardWarp, ForwardWarpLoss, FrameCompose, RefsFromFrames, Se3Loss, SimEnv, convert_ppr_warp,  # noqa: E402,F401
                  LazyFrames)

__all__ += ["ForwardKinematics", "ForwardWarp", "ForwardWarpLoss", "FrameCompose", "RefsFromFrames", "Se3Loss", "SimEnv", "convert_ppr_warp",
            "LazyFrames"]
