"""Moved: the stand-in TTA loop lives in the package now (dg_tta_b200/tta/standin.py).  This shim keeps old imports working."""
from dg_tta_b200.tta.standin import *  # noqa: F401,F403
from dg_tta_b200.tta.standin import DropInTransforms, StandInUNet, build_model, calc_branch, tta_inner_step  # noqa: F401
